"""Device timings of the flow kernels at the BASELINE.json shapes (C4 training, C5 global steps).
Usage: python scripts/bench_flow.py [--quick]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from flowmc_b200 import random as frandom, targets as T  # noqa: E402
from flowmc_b200._lib import lib  # noqa: E402
from flowmc_b200.resource.buffers import Buffer  # noqa: E402
from flowmc_b200.resource.kernel.NF_proposal import NFProposal  # noqa: E402
from flowmc_b200.resource.logPDF import LogPDF  # noqa: E402
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline  # noqa: E402
from flowmc_b200.resource.optimizer import Optimizer  # noqa: E402
from flowmc_b200.resource.states import State  # noqa: E402
from flowmc_b200.strategy.take_steps import TakeGroupSteps  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def flops_fwd(d, L, hidden, K):
    dims = [d] + hidden
    f = 0
    for i in range(len(hidden)):
        f += dims[i] * dims[i + 1] if i else (d // 2) * dims[1]
    f += hidden[-1] * (d // 2) * (3 * K + 1)
    return 2 * L * f


def main():
    quick = "--quick" in sys.argv
    out = []
    for name, d, L, hidden, K, n in (("C4 flow 32-D 10x[128,128]", 32, 10, [128, 128], 8, 16384 * (4 if not quick else 1)),
                                     ("C5 flow 64-D 8x[128,128]", 64, 8, [128, 128], 8, 16384 * (4 if not quick else 1)),
                                     ("C1 flow 5-D 4x[32,32]", 5, 4, [32, 32], 8, 65536)):
        m = MaskedCouplingRQSpline(d, L, hidden, K, frandom.PRNGKey(1))
        x = frandom.normal(frandom.PRNGKey(2), (n, d))
        fl = flops_fwd(d, L, hidden, K)
        ms = timed(lambda: m.log_prob(x))
        out.append(dict(case=name, op="log_prob", n=n, ms=ms, samples_per_s=n / ms * 1e3, useful_tflops=fl * n / ms / 1e9))
        ms = timed(lambda: m.sample(frandom.PRNGKey(3), n))
        out.append(dict(case=name, op="sample", n=n, ms=ms, samples_per_s=n / ms * 1e3, useful_tflops=fl * n / ms / 1e9))
        opt = Optimizer(m, 1e-3)
        bs = 16384
        idx = torch.arange(bs, dtype=torch.int32, device="cuda")
        ms = timed(lambda: m.train_step(x, opt.optim, opt.optim_state, idx))
        out.append(dict(case=name, op="train_step", n=bs, ms=ms, samples_per_s=bs / ms * 1e3,
                        useful_tflops=3 * fl * bs / ms / 1e9))
        print(json.dumps(out[-3]), flush=True)
        print(json.dumps(out[-2]), flush=True)
        print(json.dumps(out[-1]), flush=True)
    # C5 global steps: 65536 chains x 10 proposals
    d, n_chains, n_steps = 64, 65536 if not quick else 8192, 10
    m = MaskedCouplingRQSpline(d, 8, [128, 128], 8, frandom.PRNGKey(1))
    mu = np.zeros((8, d), np.float32)
    for i in range(8):
        mu[i, i] = 3.0 if i % 2 == 0 else -3.0
    res = {
        "p": Buffer("p", (n_chains, n_steps, d), 1), "l": Buffer("l", (n_chains, n_steps), 1),
        "a": Buffer("a", (n_chains, n_steps), 1), "s": State({"p": "p", "l": "l", "a": "a"}, "s"),
        "k": NFProposal(m), "logpdf": LogPDF(T.gaussian_mixture(mu, 1.0), n_dims=d),
    }
    x0 = frandom.normal(frandom.PRNGKey(5), (n_chains, d))
    strat = TakeGroupSteps("logpdf", "k", "s", ["p", "l", "a"], n_steps)

    def run():
        strat.set_current_position(0)
        strat(frandom.PRNGKey(9), res, x0, None)
    ms = timed(run)
    fl = flops_fwd(d, 8, [128, 128], 8)
    out.append(dict(case="C5 global steps 64-D", op="nf_global_steps", n=n_chains * n_steps, ms=ms,
                    proposals_per_s=n_chains * n_steps / ms * 1e3, useful_tflops=2 * fl * n_chains * n_steps / ms / 1e9))
    print(json.dumps(out[-1]), flush=True)
    with open("gpurun_out/bench_flow.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
