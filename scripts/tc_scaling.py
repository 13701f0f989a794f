"""Per-wave time of the tensor-core flow kernel vs number of CTAs (isolates L2 / weight-stream contention)."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from flowmc_b200 import random as frandom
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline

def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for d, L in ((32, 10), (64, 8)):
    m = MaskedCouplingRQSpline(d, L, [128, 128], 8, frandom.PRNGKey(1))
    for terms in (3, 1):
        m.tc_terms = terms
        for ctas in (1, 8, 37, 74, 148, 296):
            x = frandom.normal(frandom.PRNGKey(2), (ctas * 128, d))
            us = timed(lambda: m.log_prob(x))
            print(f"d={d} L={L} terms={terms} ctas={ctas:4d}: {us:8.1f} us  ({us / L:6.1f} us per tile-layer-wave)", flush=True)
