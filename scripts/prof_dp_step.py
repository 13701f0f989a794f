"""Per-kernel device time of the data-parallel training step on rank 0 (torch.profiler):
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/prof_dp_step.py [c4|c5]"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from flowmc_b200 import random as frandom  # noqa: E402
from flowmc_b200.parallel import ChainShard  # noqa: E402
from flowmc_b200.resource.model.nf_model.base import _TrainScratch  # noqa: E402
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline  # noqa: E402
from flowmc_b200.resource.optimizer import Optimizer  # noqa: E402

world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
case = sys.argv[1] if len(sys.argv) > 1 else "c4"
d, L = (32, 10) if case == "c4" else (64, 8)
m = MaskedCouplingRQSpline(d, L, [128, 128], 8, frandom.PRNGKey(1), device=dev)
sh = ChainShard(world, rank, world)
m.dp = (rank, world, sh.all_reduce, sh.broadcast, sh)
opt = Optimizer(m, 1e-3)
bs = 16384
x = frandom.normal(frandom.PRNGKey(2), (bs, d), device=dev)
idx = torch.arange(bs, dtype=torch.int32, device=dev)
sc = _TrainScratch(m, 0, bs)
step = m._bind_train_step(x, opt.optim, opt.optim_state, sc, bs)
for _ in range(5):
    step(idx.data_ptr())
torch.cuda.synchronize()
dist.barrier()
steps = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step(idx.data_ptr())
    torch.cuda.synchronize()
if rank == 0:
    tot, cnt = defaultdict(float), defaultdict(int)
    t_first, t_last = None, None
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            tot[ev.name[:80]] += ev.device_time
            cnt[ev.name[:80]] += 1
            t0, t1 = ev.time_range.start, ev.time_range.end
            t_first = t0 if t_first is None else min(t_first, t0)
            t_last = t1 if t_last is None else max(t_last, t1)
    print(f"{case} DP step on {world} GPUs, rank 0: device time per step by kernel (us), {steps} steps; "
          f"span per step {(t_last - t_first) / steps:.1f} us")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v / steps:9.1f}  x{cnt[k] / steps:4.1f}  {k}")
    print(f"{sum(tot.values()) / steps:9.1f}  sum")
dist.destroy_process_group()
