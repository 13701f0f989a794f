"""One small training step of the spline flow (tensor-core forward + backward with the in-kernel accumulator
reduction, then clip + AdamW) for compute-sanitizer:
  compute-sanitizer --tool memcheck  python scripts/sanitize_train.py [split]
  compute-sanitizer --tool racecheck python scripts/sanitize_train.py [split]
`split` (or FLOWMC_TC_SPLIT=2/4) runs the feature-split kernels (clusters of CTAs per tile, DSMEM exchange)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "split":
    os.environ.setdefault("FLOWMC_TC_SPLIT", "2")
else:
    os.environ.setdefault("FLOWMC_TC_SPLIT", "0")

import torch  # noqa: E402

from flowmc_b200 import random as frandom  # noqa: E402
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline  # noqa: E402
from flowmc_b200.resource.optimizer import Optimizer  # noqa: E402

d, L, rows = 16, 2, 300          # 3 tiles (the last one ragged), 2 spline chunks per layer
m = MaskedCouplingRQSpline(d, L, [128, 128], 8, frandom.PRNGKey(1))
x = frandom.normal(frandom.PRNGKey(2), (rows, d))
opt = Optimizer(m, 1e-3)
idx = torch.arange(rows, dtype=torch.int32, device="cuda")
for _ in range(2):
    loss = m.train_step(x, opt.optim, opt.optim_state, idx)
torch.cuda.synchronize()
print("mode", os.environ["FLOWMC_TC_SPLIT"], "loss", float(loss), "params finite", bool(torch.isfinite(m.params).all()))
