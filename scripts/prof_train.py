"""Per-kernel device durations of flow training steps (torch.profiler / CUPTI; warm caches, launches back to back --
unlike an ncu launch list, whose times are serialised and cold).  Usage: python scripts/prof_train.py [c4|c5] [batch]"""
import sys
from collections import defaultdict

sys.path.insert(0, ".")
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from flowmc_b200 import random as frandom  # noqa: E402
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline  # noqa: E402
from flowmc_b200.resource.optimizer import Optimizer  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "c4"
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
d, L = (32, 10) if case == "c4" else (64, 8)
m = MaskedCouplingRQSpline(d, L, [128, 128], 8, frandom.PRNGKey(1))
x = frandom.normal(frandom.PRNGKey(2), (bs, d))
opt = Optimizer(m, 1e-3)
idx = torch.arange(bs, dtype=torch.int32, device="cuda")
for _ in range(3):
    m.train_step(x, opt.optim, opt.optim_state, idx)
torch.cuda.synchronize()
steps = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        m.train_step(x, opt.optim, opt.optim_state, idx)
    torch.cuda.synchronize()
tot = defaultdict(float)
cnt = defaultdict(int)
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        tot[ev.name[:90]] += ev.device_time
        cnt[ev.name[:90]] += 1
print(f"{case} batch {bs}: device time per train step by kernel (us), {steps} steps")
s = 0.0
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v / steps:9.1f}  x{cnt[k] / steps:4.1f}  {k}")
    s += v / steps
print(f"{s:9.1f}  sum")
