"""Small runs of the remaining kernel families for compute-sanitizer --tool memcheck: NFProposal global steps (NF-mode
tensor-core kernel, target evaluation, accept scan), flow sample / inverse, RealNVP forward / inverse / loss gradient,
tempered local steps + exchange (ParallelTempering), Adam pre-optimisation.
  compute-sanitizer --tool memcheck python scripts/sanitize_misc.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from flowmc_b200 import random as frandom, targets as T  # noqa: E402
from flowmc_b200.resource.buffers import Buffer  # noqa: E402
from flowmc_b200.resource.kernel.NF_proposal import NFProposal  # noqa: E402
from flowmc_b200.resource.logPDF import LogPDF  # noqa: E402
from flowmc_b200.resource.model.nf_model.realNVP import RealNVP  # noqa: E402
from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline  # noqa: E402
from flowmc_b200.resource.optimizer import Optimizer  # noqa: E402
from flowmc_b200.resource.states import State  # noqa: E402
from flowmc_b200.strategy.take_steps import TakeGroupSteps  # noqa: E402

dev = torch.device("cuda", 0)
d, n_chains, n_steps = 6, 37, 5
m = MaskedCouplingRQSpline(d, 3, [32, 32], 8, frandom.PRNGKey(1), device=dev)
x = frandom.normal(frandom.PRNGKey(2), (300, d), device=dev)
y, ld = m.forward(x)
xi, _ = m.inverse(y)
s = m.sample(frandom.PRNGKey(3), 200)
print("flow fwd/inv round trip", float((xi - x).abs().max()), "sample", tuple(s.shape))
res = {"p": Buffer("p", (n_chains, n_steps, d), 1, device=dev), "l": Buffer("l", (n_chains, n_steps), 1, device=dev),
       "a": Buffer("a", (n_chains, n_steps), 1, device=dev), "s": State({"p": "p", "l": "l", "a": "a"}, "s"),
       "k": NFProposal(m), "logpdf": LogPDF(T.rosenbrock(), n_dims=d)}
x0 = frandom.normal(frandom.PRNGKey(5), (n_chains, d), device=dev)
TakeGroupSteps("logpdf", "k", "s", ["p", "l", "a"], n_steps)(frandom.PRNGKey(9), res, x0, None)
print("nf global steps acc", float(res["a"].data.mean()))
nvp = RealNVP(d, 4, 16, frandom.PRNGKey(4), device=dev)
yy, _ = nvp.forward(x)
xx, _ = nvp.inverse(yy)
opt = Optimizer(nvp, 1e-3)
loss = nvp.train_step(x, opt.optim, opt.optim_state)
print("realnvp round trip", float((xx - x).abs().max()), "loss", float(loss))
res["k"] = NFProposal(nvp)
TakeGroupSteps("logpdf", "k", "s", ["p", "l", "a"], n_steps)(frandom.PRNGKey(9), res, x0, None)
print("nf global steps over RealNVP acc", float(res["a"].data.mean()))

import numpy as np  # noqa: E402
from flowmc_b200.resource.kernel.MALA import MALA  # noqa: E402
from flowmc_b200.resource.logPDF import TemperedPDF  # noqa: E402
from flowmc_b200.strategy.optimization import AdamOptimization  # noqa: E402
from flowmc_b200.strategy.parallel_tempering import ParallelTempering  # noqa: E402

n_t = 4
pt_res = {"logpdf": TemperedPDF(T.iso_gaussian(0.5, "data"), None, n_dims=5, n_temps=n_t), "MALA": MALA(0.5),
          "tempered_positions": Buffer("tempered_positions", (n_chains, n_t - 1, 5), 2),
          "temperatures": Buffer("temperatures", (n_t,), 0),
          "sampler_state": State({"target_positions": "tempered_positions", "target_log_prob": "logpdf",
                                  "target_temperatures": "temperatures", "training": True}, name="sampler_state")}
pt_res["tempered_positions"].update_buffer(frandom.normal(frandom.PRNGKey(7), (n_chains, n_t - 1, 5)))
pt_res["temperatures"].update_buffer(torch.arange(n_t) + 1.0)
pt = ParallelTempering(n_steps=7, tempered_logpdf_name="logpdf", kernel_name="MALA",
                       tempered_buffer_names=["tempered_positions", "temperatures"], state_name="sampler_state")
_, _, pos = pt(frandom.PRNGKey(11), pt_res, frandom.normal(frandom.PRNGKey(8), (n_chains, 5)),
               {"data": np.arange(5, dtype=np.float32)})
print("parallel tempering", tuple(pos.shape), bool(torch.isfinite(pos).all()))
ao = AdamOptimization(LogPDF(T.rosenbrock(), n_dims=d), n_steps=9, learning_rate=0.05, noise_level=0.1)
_, _, xo = ao(frandom.PRNGKey(12), {}, x0, {})
print("adam optimisation", bool(torch.isfinite(xo).all()))
torch.cuda.synchronize()
