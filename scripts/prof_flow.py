"""Launch the flow kernels a few times (for ncu): python scripts/prof_flow.py [c4|c5] [log_prob|sample|nf|train] [n]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import torch
from flowmc_b200 import random as frandom
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
from flowmc_b200.resource.optimizer import Optimizer

case = sys.argv[1] if len(sys.argv) > 1 else "c4"
op = sys.argv[2] if len(sys.argv) > 2 else "log_prob"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 128
d, L = (32, 10) if case == "c4" else (64, 8)
m = MaskedCouplingRQSpline(d, L, [128, 128], 8, frandom.PRNGKey(1))
x = frandom.normal(frandom.PRNGKey(2), (n, d))
for _ in range(3):
    if op == "log_prob":
        m.log_prob(x)
    elif op == "sample":
        m.sample(frandom.PRNGKey(3), n)
    elif op == "train":
        opt = Optimizer(m, 1e-3)
        m.train_step(x, opt.optim, opt.optim_state)
torch.cuda.synchronize()
print("done", case, op, n)
