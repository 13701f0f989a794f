#!/usr/bin/env python
"""bench.py -- headline benchmark of the flowMC sampling hot path on B200.

Headline workload (BASELINE.json configs[1], the config the chain-steps/s metric is quoted on and the largest
local-step config that is defined for one GPU): 128-D AR(1)-correlated Gaussian, 8192 chains per GPU, MALA
step_size=0.1, one "step" = one TakeSerialSteps call of 1000 MALA steps for every chain (8.192 M chain-steps, 4.26 GB
of samples written, >> the 126 MB L2, so no L2 flush is needed between steps).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU).  The headline shards chains across ranks with no communication on
the data path (weak scaling: 8192 chains per GPU, global chain index keys).  The paths that DO communicate are
measured in the same process on the same N ranks and reported under "scaled": the full C5 Sampler (65536 chains in
total, strong scaling: chain shards + all-gather of the training set + data-parallel flow training with a gradient
all-reduce per step) and the C4 data-parallel flow-training step.  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU stand-in for the reference (the C restatement under oracle/, all host threads, set
explicitly because torchrun exports OMP_NUM_THREADS=1) on the same config: every step is the full 8192 chains x 1000
MALA steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 128
CHAINS_PER_GPU = 8192
N_LOCAL_STEPS = 1000
STEP_SIZE = 0.1
RHO = 0.9
BYTES_PER_CHAIN_STEP = 4 * (D + 2)  # position (d fp32) + log-prob + accept flag, SURVEY.md 8(d)
WORKLOAD = "C2: 128-D AR(1) Gaussian (rho=0.9), 8192 chains/GPU, MALA step_size=0.1, 1000 local steps per call"
METRIC = "chain-steps/s (MALA)"


def make_config(world: int) -> dict:
    """The `config` object of BOTH arms (the driver compares them key by key)."""
    return {"workload": WORKLOAD, "n_dim": D, "n_chains_per_gpu": CHAINS_PER_GPU,
            "n_chains_global": CHAINS_PER_GPU * world, "local_steps_per_bench_step": N_LOCAL_STEPS,
            "parallelism": f"chains sharded x{world}, no collectives on the local-step path",
            "l2": "outputs (4.26 GB per step) >> 126 MB L2, no flush needed"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_constants() -> dict:
    """Numbers that only a profiler can give (DRAM bytes, executed instructions), read from the committed summary of
    the ncu capture -- NOT measured in this run; every use carries its source."""
    p = os.path.join(ROOT, "profiles", "ncu_constants.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_first_row(self, timeout_s: float):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.02)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)]
        return {"sm_mhz": float(np.median(busy or sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


class NvmlSampler:
    """The same clocks / event reasons read in-process through NVML (the library nvidia-smi wraps) by a thread: no
    process start-up and one cheap query per field instead of nvidia-smi's full device refresh per row."""

    def __init__(self, index: int, period_s: float = 0.05):
        self.index, self.period, self.rows, self.ok = index, period_s, [], False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    @staticmethod
    def _physical_index(i: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def start(self):
        if self.ok:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def _run(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), int(get_reasons(self.h))))
            except Exception:
                pass
            self._stop.wait(self.period)

    def wait_first_row(self, timeout_s: float):
        t0 = time.perf_counter()
        while self.ok and not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def stop(self):
        if not self.ok:
            return None
        self._stop.set()
        self.thread.join(timeout=2)
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        reasons = sorted(nm for nm, b in bits.items() if any(r & b for _, r in self.rows))
        sm = [float(c) for c, _ in self.rows]
        busy = [c for c in sm if c > 0.5 * self.max_sm]
        return {"sm_mhz": float(np.median(busy or sm)) if sm else None, "sm_max_mhz": float(self.max_sm),
                "reasons": reasons, "samples": len(sm),
                "source": "NVML in-process (the counters nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.* prints), "
                          "every 50 ms during the timed region"}


def make_clock_sampler(index: int):
    """NVML in-process when pynvml is importable, else the nvidia-smi -lms loop.  FLOWMC_BENCH_CLOCKS=smi forces the
    latter (measured: its queries stall kernel launches for up to 40 ms, which lands in a 50 ms timed region)."""
    if os.environ.get("FLOWMC_BENCH_CLOCKS", "nvml") != "smi":
        s = NvmlSampler(index)
        if s.ok:
            return s
    return ClockSampler(index)


def gate(ms: float = 8.0):
    """Keeps the GPU busy for ~ms milliseconds (a spin kernel) so that the first timed launches are already queued when
    the start event fires: without it the first interval of a back-to-back loop contains the host's latency to enqueue
    the first call -- 0.6 to 2.4 ms here, 32 ms once -- while the GPU idles (measured: it was always step 1 that was
    slow, profiles/r02_bench_first_step_artifact.txt)."""
    import torch
    torch.cuda._sleep(int(ms * 1.9e6))


def bind_near_gpu(index: int):
    """Pins this process to the CPUs of the GPU's NUMA node (sysfs local_cpulist of its PCI function) so that the
    pinned host buffers of the e2e leg are first-touched on the memory the GPU's root port is attached to -- with 8
    ranks on one box the D2H copies otherwise all land on one socket.  Returns (original mask, description)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        old = os.sched_getaffinity(0)
        cpus &= old
        if not cpus or cpus == old:
            return old, f"GPU {bdf}: local CPUs {txt or '?'} = the whole mask, not bound"
        os.sched_setaffinity(0, cpus)
        return old, f"GPU {bdf}: bound to its NUMA node's CPUs {txt}"
    except Exception as ex:
        return None, f"not bound ({type(ex).__name__}: {ex})"


def stats(ms: list) -> dict:
    a = np.asarray(ms, dtype=np.float64)
    return {"min": float(a.min()), "median": float(np.median(a)), "mean": float(a.mean()), "max": float(a.max()),
            "n": int(a.size)}


def timed_calls(fn, iters: int, warm: int) -> dict:
    """Per-call CUDA-event times of `fn` launched back to back (one event pair per call, ONE statistic for every
    extra: the median; min / mean / max beside it)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    gate(3.0)
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return stats([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)])


def cpu_threads() -> int:
    return int(os.cpu_count() or 1)


def cpu_reference_rate(target_seconds: float = 12.0):
    """C restatement (oracle/c, OpenMP over chains) on a bounded sample of the workload."""
    from oracle import cref, rng, targets as otargets
    threads = cref.set_num_threads(cpu_threads())
    key = rng.PRNGKey(1)
    n = CHAINS_PER_GPU
    x0 = rng.normal(rng.split(rng.PRNGKey(0))[1], (n, D))
    data = otargets.AR1Gaussian.pack(D, RHO)
    cref.take_serial_steps(key, x0, "ar1_gaussian", data, "MALA", 1, step_size=STEP_SIZE, store=False)  # warm
    t0 = time.perf_counter()
    cref.take_serial_steps(key, x0, "ar1_gaussian", data, "MALA", 2, step_size=STEP_SIZE, store=False)
    per_step = (time.perf_counter() - t0) / 2
    steps = int(max(2, min(N_LOCAL_STEPS, target_seconds / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    cref.take_serial_steps(key, x0, "ar1_gaussian", data, "MALA", steps, step_size=STEP_SIZE, store=True)
    dt = time.perf_counter() - t0
    return n * steps / dt, threads, steps, dt


def tf32_peak(dev, n=8192, burst_iters=10, sustained_s=2.0):
    """Measured TF32 tensor-core peak of this GPU: cuBLAS fp32 GEMM with TF32 allowed (torch.matmul, n^3), best single
    launch (burst) and a seconds-long back-to-back loop (sustained), as MEASURED_PEAKS.json does for bf16.  The
    denominator for the flow kernels' TF32 fractions (BASELINE.md section 2 asks for a measured figure)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn((n, n), device=dev, dtype=torch.float32)
        b = torch.randn((n, n), device=dev, dtype=torch.float32)
        c = torch.empty((n, n), device=dev, dtype=torch.float32)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(burst_iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        iters = max(4, int(sustained_s * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / iters
        fl = 2.0 * n ** 3
        return {"burst_tflops": fl / best / 1e9, "sustained_tflops": fl / sus / 1e9,
                "how": f"torch.matmul fp32 {n}^3 with allow_tf32 (cuBLAS TF32), best of {burst_iters} / {iters} back to back"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def flow_extras(dev):
    """Device-timed measurements of the flow kernels at the BASELINE.json shapes (C4 training batch, C5 global
    steps); reported under "extra" next to the headline local-step metric."""
    import torch
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.kernel.NF_proposal import NFProposal
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    from flowmc_b200.resource.optimizer import Optimizer
    from flowmc_b200.resource.states import State
    from flowmc_b200.strategy.take_steps import TakeGroupSteps

    def useful_flops(d, L, h, K):   # transformed half of W3, conditioning half of W1
        return 2 * L * ((d // 2) * h + h * h + h * (d // 2) * (3 * K + 1))

    nc = ncu_constants()
    out = {}
    tf32_burst = None   # the TF32 GEMM probe runs LAST (two seconds at the power cap leave the clocks low for the
                        # kernels that follow it: an earlier build measured the flow extras 1.5-1.8x slow that way);
                        # the fractions of the measured peak are filled in at the end
    # C4 flow: 32-D, 10 layers, [128,128], 8 bins
    m = MaskedCouplingRQSpline(32, 10, [128, 128], 8, frandom.PRNGKey(1), device=dev)
    n = 148 * 128 * 4
    x = frandom.normal(frandom.PRNGKey(2), (n, 32), device=dev)
    st = timed_calls(lambda: m.log_prob(x), 8, 3)
    ms = st["median"]
    fl = useful_flops(32, 10, 128, 8)
    # MMAs actually issued per sample (SURVEY 8d "dense-as-written" differs: only the transformed half of W3 is
    # multiplied, but its 4-feature chunks are padded 100 -> 112 columns and W1 sees the masked inputs as zeros)
    issued = 10 * 2 * (32 * 128 + 128 * 128 + 4 * 128 * 112) * (3 if m.desc.tc_terms == 3 else 1)
    pipe_peak = 4096.0 * 148 * 1.965e9 / 1e12   # kind::tf32 128x128x8 per 64 cycles per SM, at the maximum SM clock
    out["flow_log_prob_c4"] = {
        "samples_per_s": n / ms * 1e3, "ms_per_call": st, "useful_tflops": fl * n / ms / 1e9,
        "path": f"tcgen05 {m.desc.tc_terms}xTF32" if m.desc.tc_terms else "fp32 CUDA cores",
        "issued_tflops": issued * n / ms / 1e9,
        "tensor_pipe_frac_from_rate": issued * n / ms / 1e9 / pipe_peak,
        "issued_frac_of_measured_tf32_burst": (issued * n / ms / 1e9 / tf32_burst) if tf32_burst else None,
        "useful_frac_of_measured_tf32_burst": (fl * n / ms / 1e9 / tf32_burst) if tf32_burst else None,
        "tensor_pipe_active_ncu": nc.get("flow_tc_log_prob_c4", {"note": "no committed capture"}),
        "note": "issued = 3 TF32 terms x the padded GEMM shapes; pipe peak = 4096 flop/clk/SM x 148 SMs x 1.965 GHz = "
                "1191 TFLOP/s (nominal; tensor_pipe_frac_from_rate is computed from THIS run's rate, "
                "tensor_pipe_active_ncu is copied from the committed ncu summary it names, not measured in this run)"}
    opt = Optimizer(m, 1e-3)
    bs = 16384
    idx = torch.arange(bs, dtype=torch.int32, device=dev)
    st = timed_calls(lambda: m.train_step(x, opt.optim, opt.optim_state, idx), 8, 3)
    ms = st["median"]
    out["flow_train_c4"] = {"samples_per_s": bs / ms * 1e3, "batch": bs, "ms_per_step": st,
                            "tensor_pipe_active_ncu": {k: nc[k] for k in ("flow_tc_train_forward_c4", "flow_backward_tc_c4")
                                                       if k in nc},
                            "useful_tflops": 3 * fl * bs / ms / 1e9,
                            "useful_frac_of_measured_tf32_burst": (3 * fl * bs / ms / 1e9 / tf32_burst) if tf32_burst else None,
                            "note": "training forward (tcgen05, leaves spline parameters + packed activation images) + "
                                    "hand-written backward (tcgen05 dgrad/wgrad, in-kernel deterministic reduction) + "
                                    "fused clip/AdamW; per-kernel split: scripts/prof_train.py"}
    # C5 global steps: 64-D, 8 layers, 65536 chains x 10 proposals
    d, n_chains, n_steps = 64, 65536, 10
    m5 = MaskedCouplingRQSpline(d, 8, [128, 128], 8, frandom.PRNGKey(1), device=dev)
    mu = np.zeros((8, d), np.float32)
    for i in range(8):
        mu[i, i] = 3.0 if i % 2 == 0 else -3.0
    res = {"p": Buffer("p", (n_chains, n_steps, d), 1, device=dev), "l": Buffer("l", (n_chains, n_steps), 1, device=dev),
           "a": Buffer("a", (n_chains, n_steps), 1, device=dev), "s": State({"p": "p", "l": "l", "a": "a"}, "s"),
           "k": NFProposal(m5), "logpdf": LogPDF(T.gaussian_mixture(mu, 1.0), n_dims=d)}
    x0 = frandom.normal(frandom.PRNGKey(5), (n_chains, d), device=dev)
    strat = TakeGroupSteps("logpdf", "k", "s", ["p", "l", "a"], n_steps)

    def run():
        strat.set_current_position(0)
        strat(frandom.PRNGKey(9), res, x0, None)
    st = timed_calls(run, 5, 2)
    ms = st["median"]
    out["nf_global_steps_c5"] = {"chain_steps_per_s": n_chains * n_steps / ms * 1e3, "ms_per_call": st,
                                 "useful_tflops": 2 * useful_flops(d, 8, 128, 8) * n_chains * n_steps / ms / 1e9,
                                 "note": "65536 chains x 10 NFProposal steps: flow inverse + forward, target, accept scan"}
    return out


def tf32_fractions(out: dict, dev):
    """Runs the TF32 GEMM probe (after every other extra) and fills the fractions of the measured peak in."""
    try:
        out["tf32_peak_measured"] = tf32_peak(dev)
    except Exception as ex:
        out["tf32_peak_measured"] = {"error": repr(ex)}
    tf32_burst = out["tf32_peak_measured"].get("burst_tflops")
    if tf32_burst and "flow_log_prob_c4" in out and "flow_train_c4" in out:
        lp, tr = out["flow_log_prob_c4"], out["flow_train_c4"]
        lp["issued_frac_of_measured_tf32_burst"] = lp["issued_tflops"] / tf32_burst
        lp["useful_frac_of_measured_tf32_burst"] = lp["useful_tflops"] / tf32_burst
        tr["useful_frac_of_measured_tf32_burst"] = tr["useful_tflops"] / tf32_burst


def local_extras(dev):
    """Device-timed local-step calls of the other BASELINE.json local configs (C3 HMC, the local phase of C5); reported
    under "extra" next to the headline C2 metric, same statistic (median of back-to-back calls)."""
    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.kernel.HMC import HMC
    from flowmc_b200.resource.kernel.MALA import MALA
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.resource.states import State
    from flowmc_b200.strategy.take_steps import TakeSerialSteps

    def run(kernel, target, n, d, steps):
        res = {"p": Buffer("p", (n, steps, d), 1, device=dev), "l": Buffer("l", (n, steps), 1, device=dev),
               "a": Buffer("a", (n, steps), 1, device=dev), "s": State({"p": "p", "l": "l", "a": "a"}, "s"),
               "k": kernel, "logpdf": LogPDF(target, n_dims=d)}
        strat = TakeSerialSteps("logpdf", "k", "s", ["p", "l", "a"], steps)
        x0 = frandom.normal(frandom.split(frandom.PRNGKey(0))[1], (n, d), device=dev)

        def call():
            strat.set_current_position(0)
            strat(frandom.PRNGKey(1), res, x0, None)
        st = timed_calls(call, 5, 2)
        acc = float(res["a"].data.mean())
        plan = local_plan(kernel, res["logpdf"], n, d, steps, dev)
        del res
        return n * steps / st["median"] * 1e3, st, acc, plan

    out = {}
    m = np.linspace(0.5, 2.0, 64).astype(np.float32)
    rate, st, acc, plan = run(HMC(np.diag(m), 0.01, 10), T.rosenbrock(), 32768, 64, 200)
    out["hmc_c3"] = {"chain_steps_per_s": rate, "ms_per_call": st, "acceptance_rate": acc, "launch_plan": plan,
                     "workload": "C3: 64-D Rosenbrock, 32768 chains, HMC step 0.01, 10 leapfrog steps, diagonal mass, "
                                 "200 steps per call",
                     "hbm_frac": rate * 4 * (64 + 2) / 1e9 / measured_peak()[0],
                     "note": "12 gradient evaluations + 69 threefry blocks per chain-step: fp32 / issue bound"}
    mu = np.zeros((8, 64), np.float32)
    for i in range(8):
        mu[i, i] = 3.0 if i % 2 == 0 else -3.0
    rate, st, acc, plan = run(MALA(0.1), T.gaussian_mixture(mu, 1.0), 65536, 64, 50)
    out["mala_c5_local"] = {"chain_steps_per_s": rate, "ms_per_call": st, "acceptance_rate": acc, "launch_plan": plan,
                            "workload": "C5 local phase: 64-D 8-component mixture, 65536 chains, MALA 0.1, 50 steps per call",
                            "hbm_frac": rate * 4 * (64 + 2) / 1e9 / measured_peak()[0]}
    # C2 with the SAME covariance as a dense precision matrix (reported separately; 2 d^2 flop per gradient)
    try:
        idx = np.arange(128)
        cov = (0.9 ** np.abs(idx[:, None] - idx[None, :])).astype(np.float64)
        rate, st, acc, plan = run(MALA(0.1), T.dense_gaussian(np.linalg.inv(cov).astype(np.float32)), 8192, 128, 200)
        out["mala_c2_dense_precision"] = {
            "chain_steps_per_s": rate, "ms_per_call": st, "acceptance_rate": acc, "launch_plan": plan,
            "workload": "C2 variant: the AR(1) covariance as a dense 128 x 128 precision matrix (no structure used), 8192 "
                        "chains, MALA 0.1, 200 steps per call",
            "gradient_tflops": rate * 2.0 * 128 * 128 / 1e12, "hbm_frac": rate * 4 * (128 + 2) / 1e9 / measured_peak()[0]}
    except Exception as ex:
        out["mala_c2_dense_precision"] = {"error": repr(ex)}
    return out


def local_plan(kernel, logpdf, n, d, n_steps, dev) -> dict:
    """What flowmc_local_steps decides for this call: lane layout, resident slots, time slicing, launches."""
    import ctypes as C
    import torch
    from flowmc_b200._lib import check, lib
    p, keep = kernel._local_params(d, dev)
    ws = torch.empty(max(256, int(lib.flowmc_local_steps_workspace_bytes(n, d, p.layout_hint))), dtype=torch.uint8,
                     device=dev)
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    out = (C.c_int * 12)()
    with torch.cuda.device(dev):
        check(lib.flowmc_local_steps_plan(kernel.KIND, logpdf.target.target_id, n, d, n_steps, C.byref(p), out))
    names = ["layout", "lanes_per_chain", "dims_per_lane", "store_vec", "chain_groups", "resident_cta_slots",
             "ctas_per_sm", "static_smem_per_cta", "n_seg", "seg_len", "n_launches", "ctas_per_launch"]
    return dict(zip(names, [int(v) for v in out]))


def flow_train_dp(dev, rank, world, iters=20, warm=5, d=32, n_layers=10, tag="C4", bs=16384):
    """C4 (BASELINE.json configs[3]): one NFModel.train_step on a GLOBAL batch of 16384 rows, data-parallel over the
    `world` ranks: each rank takes 16384 / world rows (feature-split tensor-core kernels when that leaves it only a
    few tiles), then gradient all-reduce + global-norm clip + AdamW in ONE kernel over NVLink peer memory
    (flowmc_dp_reduce_adamw; FLOWMC_DP_PEER=0: NCCL all-reduce + flowmc_clip_adamw).  Collective: every rank calls it.
    Device time, max over ranks."""
    import torch
    import torch.distributed as dist
    from flowmc_b200 import random as frandom
    from flowmc_b200.parallel import ChainShard
    from flowmc_b200.resource.model.nf_model.rqSpline import MaskedCouplingRQSpline
    from flowmc_b200.resource.optimizer import Optimizer
    m = MaskedCouplingRQSpline(d, n_layers, [128, 128], 8, frandom.PRNGKey(1), device=dev)
    if world > 1:
        sh = ChainShard(world, rank, world)
        m.dp = (rank, world, sh.all_reduce, sh.broadcast, sh)
    opt = Optimizer(m, 1e-3)
    x = frandom.normal(frandom.PRNGKey(2), (bs * 2, d), device=dev)
    idx = torch.arange(bs, dtype=torch.int32, device=dev)
    from flowmc_b200.resource.model.nf_model.base import _TrainScratch
    sc = _TrainScratch(m, 0, bs)
    # the call sequence NFModel.train_epoch runs per batch (train_step with its per-step lookups bound once)
    step = m._bind_train_step(x, opt.optim, opt.optim_state, sc, bs)
    idx_ptr = idx.data_ptr()
    for _ in range(warm):
        step(idx_ptr)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gate(3.0)
    e0.record()
    for _ in range(iters):
        step(idx_ptr)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    collective = "none (single GPU)"
    if world > 1:
        collective = ("fused: reduce-scatter + all-gather over NVLink peer memory inside the optimiser kernel "
                      "(flowmc_dp_reduce_adamw)") if sc.peer is not None else "NCCL all-reduce + flowmc_clip_adamw"
    return {"workload": f"{tag}: flow {d}-D, {n_layers} layers, [128,128], 8 bins; train_step on a global batch of "
                        f"{bs} rows split over {world} GPU(s)", "n_gpus": world, "global_batch": bs, "ms_per_step": ms,
            "samples_per_s": bs / ms * 1e3, "grad_allreduce_bytes": int(m.params.numel()) * 4 if world > 1 else 0,
            "collective": collective, "timing": f"{iters} back-to-back steps, CUDA events, max over ranks"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from oracle import cref, rng, targets as otargets
    threads = cref.set_num_threads(cpu_threads())     # torchrun exports OMP_NUM_THREADS=1
    key = rng.PRNGKey(1)
    n = CHAINS_PER_GPU
    x0 = rng.normal(rng.split(rng.PRNGKey(0))[1], (n, D))
    data = otargets.AR1Gaussian.pack(D, RHO)
    cref.take_serial_steps(key, x0, "ar1_gaussian", data, "MALA", 2, step_size=STEP_SIZE, store=False)
    for _ in range(args.warmup):
        cref.take_serial_steps(key, x0, "ar1_gaussian", data, "MALA", N_LOCAL_STEPS, step_size=STEP_SIZE)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cref.take_serial_steps(key, x0, "ar1_gaussian", data, "MALA", N_LOCAL_STEPS, step_size=STEP_SIZE)
    dt = time.perf_counter() - t0
    rate = n * N_LOCAL_STEPS * args.steps / dt
    sample = (f"{n} chains x {N_LOCAL_STEPS} MALA steps per bench step (the full per-GPU workload"
              + (f"; 1/{world} of the {n * world} global chains -- chain-steps/s does not depend on the chain count"
                 if world > 1 else "") + "), samples stored to host memory")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "chain-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(world),
        "cpu_baseline": {"value": rate, "unit": "chain-steps/s", "cores": threads, "kind": "port",
                         "sample": sample,
                         "note": "C restatement of flowMC's MALA path (oracle/c, OpenMP over chains, pinned to the numpy "
                                 "oracle by tests/test_oracle_c.py); the reference's own JAX CPU path cannot run here "
                                 "(jax not installable)"},
        "e2e": {"value": rate, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from flowmc_b200 import random as frandom, targets as T
    from flowmc_b200._lib import lib
    from flowmc_b200.resource.buffers import Buffer
    from flowmc_b200.resource.kernel.MALA import MALA
    from flowmc_b200.resource.logPDF import LogPDF
    from flowmc_b200.resource.states import State
    from flowmc_b200.strategy.take_steps import TakeSerialSteps

    n = CHAINS_PER_GPU
    n_global = n * world
    resources = {
        "positions": Buffer("positions", (n, N_LOCAL_STEPS, D), 1, device=dev),
        "log_prob": Buffer("log_prob", (n, N_LOCAL_STEPS), 1, device=dev),
        "acceptance": Buffer("acceptance", (n, N_LOCAL_STEPS), 1, device=dev),
        "state": State({"p": "positions", "l": "log_prob", "a": "acceptance"}, name="state"),
        "kernel": MALA(step_size=STEP_SIZE),
        "logpdf": LogPDF(T.ar1_gaussian(RHO), n_dims=D),
    }
    strat = TakeSerialSteps("logpdf", "kernel", "state", ["p", "l", "a"], N_LOCAL_STEPS)
    strat.set_chain_shard(rank * n, n_global)
    # initial positions: normal(split(PRNGKey(0))[1], (n_global, d)) -- this rank's rows
    x0_all_key = frandom.split(frandom.PRNGKey(0))[1]
    x0 = frandom.normal(x0_all_key, (n_global, D), device=dev)[rank * n:(rank + 1) * n].contiguous()
    old_mask, affinity_note = bind_near_gpu(local_rank)   # before the pinned buffers are first touched
    x0_host = x0.cpu().pin_memory()
    key = frandom.PRNGKey(1)
    plan = local_plan(resources["kernel"], resources["logpdf"], n, D, N_LOCAL_STEPS, dev)

    def step_device(k):
        strat.set_current_position(0)
        k, _, last = strat(k, resources, x0, None)
        return k, last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`): K calls back to back ------------------------------------------
    # the clock sampler (nvidia-smi -lms) is started BEFORE the warm-up and must have delivered its first row before
    # the timed region begins: its start-up (NVML initialisation on an 8-GPU box takes up to a second) stalls launches
    # for milliseconds and used to land inside the timed steps (per_step_ms.max 8.1 ms against a 4.88 ms median)
    sampler = make_clock_sampler(local_rank)
    if rank == 0:
        sampler.start()
    k = key
    for _ in range(args.warmup):
        k, _ = step_device(k)
    barrier()
    if rank == 0:
        sampler.wait_first_row(5.0)
    launches0 = lib.flowmc_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    gate()
    ev[0].record()
    for i in range(args.steps):
        k, last = step_device(k)
        ev[i + 1].record()
    barrier()
    launches = lib.flowmc_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    # ... and the same calls with a device synchronisation after each (launch-to-launch effects show as a difference)
    sync_ms = []
    for i in range(min(args.steps, 10)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        k, last = step_device(k)
        e1.record()
        torch.cuda.synchronize()
        sync_ms.append(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    acc_rate = float(resources["acceptance"].data.mean())
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    chain_steps_per_step = n_global * N_LOCAL_STEPS
    value = chain_steps_per_step * args.steps / (total_ms_max * 1e-3)

    # ---- end-to-end through the strategy API with host buffers (`e2e`) ------------------------
    out_pos = torch.empty((n, N_LOCAL_STEPS, D), dtype=torch.float32).pin_memory()
    out_lp = torch.empty((n, N_LOCAL_STEPS), dtype=torch.float32).pin_memory()
    out_acc = torch.empty((n, N_LOCAL_STEPS), dtype=torch.float32).pin_memory()
    out_last = torch.empty((n, D), dtype=torch.float32).pin_memory()

    def step_e2e(k, full):
        strat.set_current_position(0)
        xin = x0_host.to(dev, non_blocking=True)                 # H2D of this step's input
        k, _, last = strat(k, resources, xin, None)
        out_last.copy_(last, non_blocking=True)                   # D2H of the strategy's return value
        if full:                                                  # D2H of the three sample buffers
            out_pos.copy_(resources["positions"].data, non_blocking=True)
            out_lp.copy_(resources["log_prob"].data, non_blocking=True)
            out_acc.copy_(resources["acceptance"].data, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return k

    e2e_steps = max(1, min(args.steps, 5))
    res_e2e = {}
    for full in (True, False):
        kk = key
        kk = step_e2e(kk, full)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            kk = step_e2e(kk, full)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res_e2e[full] = chain_steps_per_step * e2e_steps / float(dt.item())
    h2d = x0_host.numel() * 4
    d2h_full = (out_pos.numel() + out_lp.numel() + out_acc.numel() + out_last.numel()) * 4
    if old_mask is not None:
        os.sched_setaffinity(0, old_mask)   # the CPU legs below use every host core
    del out_pos, out_lp, out_acc
    del resources["positions"], resources["log_prob"], resources["acceptance"]
    torch.cuda.empty_cache()

    # ---- the communicating paths on the same ranks (collective: every rank takes part) -------------------------
    # FLOWMC_BENCH_EXTRAS=0 skips everything below the headline (used for the ncu launch list of the timed step itself)
    extras_on = os.environ.get("FLOWMC_BENCH_EXTRAS", "1") != "0"
    scaled = {"skipped": "FLOWMC_BENCH_EXTRAS=0"}
    if extras_on:
        scaled = {}
        try:
            scaled["flow_train_c4_dp"] = flow_train_dp(dev, rank, world)
            if world > 1:
                # the same step with the per-GPU work held at 16384 rows (a user scaling batch_size with the GPUs)
                scaled["flow_train_c4_dp_weak"] = flow_train_dp(dev, rank, world, bs=16384 * world)
        except Exception as ex:  # the headline line must still be printed
            scaled["flow_train_c4_dp"] = {"error": repr(ex)}
        try:
            from scripts.bench_sampler import run_sampler
            scaled["sampler_c5"] = run_sampler(dev, rank, world)
        except Exception as ex:
            scaled["sampler_c5"] = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    extras = {"skipped": "FLOWMC_BENCH_EXTRAS=0" if not extras_on else "single-GPU kernel extras are reported at N=1"}
    if extras_on and world == 1:
        try:
            extras = flow_extras(dev)
            extras.update(local_extras(dev))
            try:   # C5 with the bundle's default flow (4 x [32, 32] x 8, RQSpline_MALA.py defaults), SURVEY 8d
                from scripts.bench_sampler import run_sampler
                r = run_sampler(dev, 0, 1, flow=(4, [32, 32]))
                extras["sampler_c5_bundle_default_flow"] = {k: r[k] for k in (
                    "workload", "wall_s", "phase_ms_rank0", "flow_train_ms_per_step", "sampler_chain_steps_per_s",
                    "ess_per_s", "global_acceptance", "local_acceptance")}
            except Exception as ex:
                extras["sampler_c5_bundle_default_flow"] = {"error": repr(ex)}
            tf32_fractions(extras, dev)
        except Exception as ex:
            extras = {"error": repr(ex)}
    peak, peak_src = measured_peak()
    kst = stats(kernel_ms)
    avg_kernel_ms = kst["mean"]
    achieved = n * N_LOCAL_STEPS * BYTES_PER_CHAIN_STEP / (avg_kernel_ms * 1e-3) / 1e9
    nc = ncu_constants().get("local_steps_c2", {})
    traffic = nc.get("dram_bytes_per_step")
    instr = nc.get("warp_instr_per_chain_step")
    sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "chain-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(world),
        "acceptance_rate": acc_rate,
        "e2e": {"value": res_e2e[True], "unit": "chain-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h_full, "host_affinity_rank0": affinity_note,
                "note": "TakeSerialSteps call with pinned-host initial positions in and ALL sample buffers "
                        "(positions, log-probs, accept flags, last position) copied to pinned host memory"},
        "e2e_device_resident_buffers": {
            "value": res_e2e[False], "unit": "chain-steps/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": out_last.numel() * 4,
            "note": "same call, buffers stay on the device as in the reference (jax arrays); only the "
                    "strategy's return value (positions[:, -1]) is read back"},
        "gpu_launches": int(launches),
        "launch_plan": plan,
        "per_step_ms": {"back_to_back": kst, "sync_per_step": stats(sync_ms),
                        "back_to_back_each": [round(v, 4) for v in kernel_ms],
                        "note": "CUDA-event time of each of the K timed steps (one TakeSerialSteps call = "
                                f"{plan['n_launches']} launches of local_steps_kernel); `value` uses the back-to-back total"},
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "traffic_source": nc.get("source", "no committed ncu capture") + " (ncu --set full, dram__bytes_read.sum "
                                       "+ dram__bytes_write.sum summed over the launches of one step; NOT measured in this run)",
                     "kernel": "flowmc::local_steps_kernel<AR1Gaussian, MALA, Layout<16,8,4>>",
                     "algorithmic_bytes_per_launch": n * N_LOCAL_STEPS * BYTES_PER_CHAIN_STEP,
                     "launches_per_step": plan["n_launches"],
                     "avg_launch_ms": avg_kernel_ms,
                     "avg_launch_ms_note": "mean over the K timed steps of the device time of one step (all of its "
                                           "local_steps_kernel launches); algorithmic bytes are per step likewise",
                     "issue_slots": {
                         "warp_instructions_per_chain_step": instr,
                         "achieved_frac": ((n * N_LOCAL_STEPS / (avg_kernel_ms * 1e-3)) * instr / (148 * 4 * sm_hz))
                         if instr else None,
                         "source": nc.get("source", "no committed ncu capture") + " (smsp__inst_executed.sum / chain-steps; "
                                   "NOT measured in this run) x this run's chain-steps/s over 148 SMs x 4 schedulers x the "
                                   "SM clock sampled during the run"},
                     "note": "bit-exact threefry2x32 (d+5 blocks per chain-step) + XLA's erf_inv make this kernel "
                             "instruction-issue bound, not HBM bound: see issue_slots and DESIGN.md 4.1"},
        "scaled": scaled,
        "extra": extras,
    }
    if world == 1:
        cpu_rate, cpu_thr, cpu_steps, cpu_dt = cpu_reference_rate()
        line["cpu_baseline"] = {"value": cpu_rate, "unit": "chain-steps/s", "cores": cpu_thr, "kind": "port",
                                "sample": f"{n} chains x {cpu_steps} MALA steps ({cpu_dt:.1f} s), same target and seeds"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
