"""Does sampling the clocks perturb the timed kernel?  40 back-to-back C2 local-step calls (4.9 ms each) with no
sampler, with the nvidia-smi -lms loop and with the in-process NVML thread: per-step min / median / max.
python scripts/clock_sampler_probe.py"""
import json
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
from flowmc_b200 import random as frandom, targets as T  # noqa: E402
from flowmc_b200.resource.buffers import Buffer  # noqa: E402
from flowmc_b200.resource.kernel.MALA import MALA  # noqa: E402
from flowmc_b200.resource.logPDF import LogPDF  # noqa: E402
from flowmc_b200.resource.states import State  # noqa: E402
from flowmc_b200.strategy.take_steps import TakeSerialSteps  # noqa: E402

dev = torch.device("cuda", 0)
n, d, steps = 8192, 128, 1000
res = {"logpdf": LogPDF(T.ar1_gaussian(0.9), n_dims=d), "kernel": MALA(step_size=0.1),
       "positions": Buffer("positions", (n, steps, d), 1, device=dev), "log_prob": Buffer("log_prob", (n, steps), 1, device=dev),
       "acceptance": Buffer("acceptance", (n, steps), 1, device=dev),
       "state": State({"positions": "positions", "log_prob": "log_prob", "acceptance": "acceptance"}, "state")}
strat = TakeSerialSteps("logpdf", "kernel", "state", ["positions", "log_prob", "acceptance"], steps)
x0 = frandom.normal(frandom.PRNGKey(0), (n, d), device=dev)


def run(k):
    strat.set_current_position(0)
    k, _, _ = strat(k, res, x0, None)
    return k


k = frandom.PRNGKey(1)
for _ in range(3):
    k = run(k)
torch.cuda.synchronize()
for name in ("none", "smi", "nvml", "none", "smi", "nvml"):
    s = None if name == "none" else (bench.ClockSampler(0) if name == "smi" else bench.NvmlSampler(0))
    if s is not None:
        s.start()
        s.wait_first_row(5.0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(40):
        k = run(k)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(40)]
    out = {"sampler": name, "per_step_ms": bench.stats(ms), "steps_over_1.1x_median": sum(m > 1.1 * sorted(ms)[20] for m in ms)}
    if s is not None:
        out["clocks"] = s.stop()
    print(json.dumps(out), flush=True)
    time.sleep(0.5)
