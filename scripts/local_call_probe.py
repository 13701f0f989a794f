"""Time one TakeSerialSteps call of the C5 local phase (65536 chains, 64-D mixture, 50 MALA steps) into a compact
buffer vs the sampler's long, strided buffer (n_total = 240)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from flowmc_b200 import random as frandom, targets as T
from flowmc_b200.resource.buffers import Buffer
from flowmc_b200.resource.kernel.MALA import MALA
from flowmc_b200.resource.logPDF import LogPDF
from flowmc_b200.resource.states import State
from flowmc_b200.strategy.take_steps import TakeSerialSteps

n, d, steps = 65536, 64, 50
mu = np.zeros((8, d), np.float32)
for i in range(8):
    mu[i, i] = 3.0 if i % 2 == 0 else -3.0
for n_total in (50, 240):
    res = {"p": Buffer("p", (n, n_total, d), 1), "l": Buffer("l", (n, n_total), 1), "a": Buffer("a", (n, n_total), 1),
           "s": State({"p": "p", "l": "l", "a": "a"}, "s"), "k": MALA(0.1), "logpdf": LogPDF(T.gaussian_mixture(mu, 1.0), n_dims=d)}
    strat = TakeSerialSteps("logpdf", "k", "s", ["p", "l", "a"], steps)
    x0 = frandom.normal(frandom.PRNGKey(5), (n, d))
    for cursor in (0, 0, 0, n_total - steps):
        strat.set_current_position(cursor)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        strat(frandom.PRNGKey(9), res, x0, None)
        e1.record()
        torch.cuda.synchronize()
        print(f"n_total={n_total} cursor={cursor}: {e0.elapsed_time(e1):.3f} ms")
