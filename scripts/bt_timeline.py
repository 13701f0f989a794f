"""Timeline (clock64) of epilogue thread 0, CTA 0 of the tensor-core backward kernel for one C4 training step."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from flowmc_b200 import random as frandom
from flowmc_b200._lib import lib
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline

d, L = (32, 10) if (len(sys.argv) < 2 or sys.argv[1] == "c4") else (64, 8)
m = MaskedCouplingRQSpline(d, L, [128, 128], 8, frandom.PRNGKey(1))
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 16384   # few rows (e.g. 2048) -> the feature-split kernels
x = frandom.normal(frandom.PRNGKey(2), (rows, d))
m.loss_and_grad(x)
buf = torch.zeros(4 * 256, dtype=torch.int64, device="cuda")
lib.flowmc_trace_tc_timeline(buf.data_ptr())
m.loss_and_grad(x)
torch.cuda.synchronize()
lib.flowmc_trace_tc_timeline(None)
t = buf.cpu().numpy().reshape(4, 256)[3]
if t[252] > 0 and t[253] > 0:
    first = t[:252][t[:252] > 0]
    print(f"CTA 0 (cycles): kernel entry -> first unit prepared {int(first[0] - t[252])}, first -> last stamp "
          f"{int(first[-1] - first[0])}, last stamp -> tiles done {int(t[253] - first[-1])}, tiles done -> reduction done "
          f"{int(t[254] - t[253]) if t[254] > 0 else -1}")
t = t[:252]
v = t[t > 0]
v = v - v[0]
print("backward epilogue stamps per chunk: [operand written, dgrad done, transposed written, wgrad done, reduced]")
print(" ".join(str(int(c)) for c in v[:120]))
import numpy as np
dv = np.diff(v[:100])
print("deltas:", " ".join(str(int(c)) for c in dv))
