"""Time-slicing sweep of the persistent local-step kernel at the BASELINE shapes: device time of one TakeSerialSteps
call with slicing forced off (force_n_seg = -1), automatic (0) and forced to S segments.  Prints one JSON line per
(config, setting): median / min of 7 back-to-back calls."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import local_plan, timed_calls  # noqa: E402


def main():
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.kernel.HMC import HMC
    from flowmc_b200.resource.kernel.MALA import MALA
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.resource.states import State
    from flowmc_b200.strategy.take_steps import TakeSerialSteps
    dev = torch.device("cuda", 0)
    mu = np.zeros((8, 64), np.float32)
    for i in range(8):
        mu[i, i] = 3.0 if i % 2 == 0 else -3.0
    cases = {
        "C2": (MALA(0.1), T.ar1_gaussian(0.9), 8192, 128, 1000),
        "C3": (HMC(np.diag(np.linspace(0.5, 2.0, 64).astype(np.float32)), 0.01, 10), T.rosenbrock(), 32768, 64, 200),
        "C5-local": (MALA(0.1), T.gaussian_mixture(mu, 1.0), 65536, 64, 50),
    }
    only = sys.argv[1:] or list(cases)
    for name in only:
        kernel, target, n, d, steps = cases[name]
        res = {"p": Buffer("p", (n, steps, d), 1, device=dev), "l": Buffer("l", (n, steps), 1, device=dev),
               "a": Buffer("a", (n, steps), 1, device=dev), "s": State({"p": "p", "l": "l", "a": "a"}, "s"),
               "k": kernel, "logpdf": LogPDF(target, n_dims=d)}
        strat = TakeSerialSteps("logpdf", "k", "s", ["p", "l", "a"], steps)
        x0 = frandom.normal(frandom.split(frandom.PRNGKey(0))[1], (n, d), device=dev)

        def call():
            strat.set_current_position(0)
            strat(frandom.PRNGKey(1), res, x0, None)
        for seg in (-1, 0, 2, 3, 4, 5, 6, 8):
            if seg > 1 and steps // seg < 8:
                continue
            kernel.force_n_seg = seg
            st = timed_calls(call, 7, 2)
            plan = local_plan(kernel, res["logpdf"], n, d, steps, dev)
            print(json.dumps({"config": name, "force_n_seg": seg, "median_ms": st["median"], "min_ms": st["min"],
                              "chain_steps_per_s": n * steps / st["median"] * 1e3,
                              "plan": {k: plan[k] for k in ("n_seg", "seg_len", "n_launches", "ctas_per_launch",
                                                            "resident_cta_slots", "chain_groups")}}), flush=True)
        kernel.force_n_seg = 0
        del res
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
