"""Host-side cost of one NFModel.train_step: wall time per step of a loop whose GPU work is tiny (128 rows), i.e. the
time Python + ctypes + the C launchers need to enqueue a step.  python scripts/host_overhead.py"""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from flowmc_b200 import random as frandom  # noqa: E402
from flowmc_b200.resource.model.nf_model.base import _TrainScratch  # noqa: E402
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline  # noqa: E402
from flowmc_b200.resource.optimizer import Optimizer  # noqa: E402

for d, L in ((32, 10), (64, 8)):
    m = MaskedCouplingRQSpline(d, L, [128, 128], 8, frandom.PRNGKey(1))
    opt = Optimizer(m, 1e-3)
    for bs in (128, 16384):
        x = frandom.normal(frandom.PRNGKey(2), (bs, d))
        idx = torch.arange(bs, dtype=torch.int32, device="cuda")
        sc = _TrainScratch(m, 0, bs)
        for _ in range(20):
            m.train_step(x, opt.optim, opt.optim_state, idx, sc)
        torch.cuda.synchronize()
        n = 60   # below the driver's launch-queue depth (~1000 launches): beyond it the host is throttled to the GPU rate
        t0 = time.perf_counter()
        for _ in range(n):
            m.train_step(x, opt.optim, opt.optim_state, idx, sc)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"d={d} L={L} batch={bs}: train_step enqueue {1e6 * (t1 - t0) / n:.1f} us/step, with drain {1e6 * (t2 - t0) / n:.1f} us/step")
        step = m._bind_train_step(x, opt.optim, opt.optim_state, sc, bs)   # what train_epoch runs per batch
        ip = idx.data_ptr()
        for _ in range(20):
            step(ip)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            step(ip)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"d={d} L={L} batch={bs}: bound step  enqueue {1e6 * (t1 - t0) / n:.1f} us/step, with drain {1e6 * (t2 - t0) / n:.1f} us/step")
