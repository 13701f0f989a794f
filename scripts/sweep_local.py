"""GPU experiment: time every valid lane layout of the local-steps kernel on the BASELINE shapes."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from flowmc_b200 import random as frandom, targets as T
from flowmc_b200._lib import LocalParams, check, lib

LAYOUTS = [(1, 8, 1), (4, 8, 1), (8, 8, 1), (32, 4, 1), (32, 16, 1), (8, 4, 4), (16, 4, 4), (16, 8, 4),
           (32, 4, 4), (32, 8, 4), (32, 16, 4)]
u32p = C.POINTER(C.c_uint32)


def run(kind, tgt, d, n_chains, n_steps, hint, step_size, n_leapfrog=0, reps=3, use_ws=True):
    dev = torch.device("cuda", 0)
    key = frandom.PRNGKey(1)
    x0 = frandom.normal(frandom.split(frandom.PRNGKey(0))[1], (n_chains, d))
    pos = torch.empty((n_chains, n_steps, d), device=dev)
    lp = torch.empty((n_chains, n_steps), device=dev)
    acc = torch.empty((n_chains, n_steps), device=dev)
    last = torch.empty((n_chains, d), device=dev)
    p = LocalParams()
    p.step_size = step_size
    p.n_leapfrog = n_leapfrog
    p.layout_hint = hint
    keep = []
    if kind == 1:
        m = np.linspace(0.5, 2.0, d).astype(np.float32)
        L = torch.from_numpy(np.diag(1 / np.sqrt(m)).astype(np.float32)).to(dev)
        cs = torch.from_numpy(m).to(dev)
        p.hmc_chol, p.hmc_colsum, p.hmc_chol_diagonal = L.data_ptr(), cs.data_ptr(), 1
        keep = [L, cs]
    wsb = int(lib.flowmc_local_steps_workspace_bytes(n_chains, d, hint)) if use_ws else 0
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    p.workspace = ws.data_ptr() if use_ws else None
    p.workspace_bytes = ws.numel() if use_ws else 0
    pk = tgt.packed_on(None, d, dev)
    ko = np.zeros(2, np.uint32)
    times = []
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        check(lib.flowmc_local_steps(kind, tgt.target_id, pk.data_ptr(), key.ctypes.data_as(u32p), x0.data_ptr(),
                                     pos.data_ptr(), lp.data_ptr(), acc.data_ptr(), n_steps, 0, n_chains, d, n_steps, 1,
                                     0, n_chains, C.byref(p), ko.ctypes.data_as(u32p), last.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        if r:
            times.append(e0.elapsed_time(e1))
    ms = min(times)
    return ms, n_chains * n_steps / ms * 1e3, float(acc.mean())


if __name__ == "__main__":
    out = []
    cases = [("C2 MALA ar1 d128 n8192", 0, T.ar1_gaussian(0.9), 128, 8192, 1000, 0.1, 0),
             ("C5loc MALA mix d64 n65536", 0, T.gaussian_mixture(np.random.RandomState(0).randn(8, 64) * 3), 64, 65536, 96, 0.1, 0),
             ("C3 HMC rosen d64 n32768", 1, T.rosenbrock(), 64, 32768, 64, 0.01, 10),
             ("GRW iso d128 n8192", 2, T.iso_gaussian(0.5, None), 128, 8192, 320, 0.1, 0)]
    for name, kind, tgt, d, n, steps, ss, nl in cases:
        for h, (G, DPL, VEC) in enumerate(LAYOUTS, 1):
            if G * DPL < d or (VEC == 4 and d % 4) or G * DPL > 4 * d:
                continue
            for use_ws in (True,):
                try:
                    ms, rate, acc = run(kind, tgt, d, n, steps, h, ss, nl, use_ws=use_ws)
                    bytes_per = 4 * (d + 2)
                    rec = dict(case=name, layout=[G, DPL, VEC], sliced=use_ws, ms=round(ms, 3),
                               chain_steps_per_s=round(rate / 1e6, 1), GBps=round(rate * bytes_per / 1e9, 1), acc=acc)
                except Exception as ex:  # noqa
                    rec = dict(case=name, layout=[G, DPL, VEC], error=str(ex))
                print(json.dumps(rec), flush=True)
                out.append(rec)
    json.dump(out, open("/root/repo/gpurun_out/sweep_local.json", "w"), indent=1)
