"""Small time-sliced runs of the local-step kernel (MALA d=128, HMC d=64, GRW d=12; 3-5 segments, rounds of a few
CTAs) for compute-sanitizer:
  compute-sanitizer --tool memcheck  python scripts/sanitize_sliced.py
  compute-sanitizer --tool racecheck python scripts/sanitize_sliced.py
Checks bit-identity with the unsliced launch on the way."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.kernel.Gaussian_random_walk import GaussianRandomWalk
    from flowmc_b200.resource.kernel.HMC import HMC
    from flowmc_b200.resource.kernel.MALA import MALA
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.resource.states import State
    from flowmc_b200.strategy.take_steps import TakeSerialSteps
    dev = torch.device("cuda", 0)
    cases = [(MALA(0.1), T.ar1_gaussian(0.9), 128), (HMC(np.eye(64, dtype=np.float32), 0.01, 3), T.rosenbrock(), 64),
             (GaussianRandomWalk(0.1), T.rosenbrock(), 12)]
    for kernel, target, d in cases:
        n, steps = 21, 70
        outs = []
        for seg, slots in ((-1, 0), (4, 5)):
            kernel.force_n_seg, kernel.slots_override = seg, slots
            res = {"p": Buffer("p", (n, steps, d), 1, device=dev), "l": Buffer("l", (n, steps), 1, device=dev),
                   "a": Buffer("a", (n, steps), 1, device=dev), "s": State({"p": "p", "l": "l", "a": "a"}, "s"),
                   "k": kernel, "logpdf": LogPDF(target, n_dims=d)}
            strat = TakeSerialSteps("logpdf", "k", "s", ["p", "l", "a"], steps)
            x0 = frandom.normal(frandom.PRNGKey(0), (n, d), device=dev)
            _, res, last = strat(frandom.PRNGKey(1), res, x0, None)
            torch.cuda.synchronize()
            outs.append((res["p"].data.clone(), res["a"].data.clone(), last.clone()))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
        print(f"{kernel!r} d={d}: sliced == unsliced", flush=True)


if __name__ == "__main__":
    main()
