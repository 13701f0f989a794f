"""Pipeline timeline of CTA 0 of the tensor-core flow kernel (clock64 stamps) for one log_prob launch
(python scripts/tc_timeline.py [c4|c5] [terms] [train] [rows]: the training forward of one loss_and_grad call)."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from flowmc_b200 import random as frandom
from flowmc_b200._lib import lib
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline

d, L = (32, 10) if (len(sys.argv) < 2 or sys.argv[1] == "c4") else (64, 8)
terms = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = MaskedCouplingRQSpline(d, L, [128, 128], 8, frandom.PRNGKey(1))
m.tc_terms = terms
train = len(sys.argv) > 3 and sys.argv[3] == "train"
rows = int(sys.argv[4]) if len(sys.argv) > 4 else (128 if train else 148) * 128
x = frandom.normal(frandom.PRNGKey(2), (rows, d))
run = (lambda: m.loss_and_grad(x)) if train else (lambda: m.log_prob(x))
run()
buf = torch.zeros(4 * 256, dtype=torch.int64, device="cuda")
lib.flowmc_trace_tc_timeline(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.flowmc_trace_tc_timeline(None)
t = buf.cpu().numpy().reshape(4, 256)[:3]
t0 = t[t > 0].min()
for role, name in enumerate(("producer (stage slot free -> copy issued)", "mma (ready | first stage | issued)",
                             "epilogue thread 0")):
    v = t[role][t[role] > 0] - t0
    print(name, len(v))
    print(" ", " ".join(str(int(c)) for c in v[:120]))
