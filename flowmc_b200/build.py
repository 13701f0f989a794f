"""In-tree build of libflowmc_b200.so (sm_100a only) with nvcc.

``python -m flowmc_b200.build`` compiles every ``csrc/**/*.cu`` to an object under ``build/`` (in
parallel, incremental by mtime) and links ``flowmc_b200/lib/libflowmc_b200.so``.  The shared
object stays in-tree so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
CSRC = ROOT / "csrc"
BUILD = REPO / "build" / "obj"
LIBDIR = ROOT / "lib"
LIB = LIBDIR / "libflowmc_b200.so"
TEST_LIB = LIBDIR / "libflowmc_b200_test.so"   # csrc/testlib/*.cu: probe kernels for tests, not product code

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _headers_mtime() -> float:
    hs = list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.h")) + list((REPO / "include").glob("*"))
    return max(p.stat().st_mtime for p in hs)


def sources() -> list[Path]:
    """Product sources: every .cu under csrc/ except the test scaffolding in csrc/testlib/."""
    return sorted(p for p in CSRC.rglob("*.cu") if "testlib" not in p.relative_to(CSRC).parts)


def test_sources() -> list[Path]:
    return sorted((CSRC / "testlib").glob("*.cu"))


def _compile_host(src: Path, verbose: bool) -> Path:
    """Host-only C++ sources of the product library (the XLA-FFI shim: an empty translation unit unless jaxlib's
    xla/ffi/api/ffi.h is on the include path -- add it with FLOWMC_XLA_FFI_INCLUDE)."""
    obj = BUILD / src.relative_to(CSRC).with_suffix(".o")
    obj.parent.mkdir(parents=True, exist_ok=True)
    if (not obj.exists()) or obj.stat().st_mtime < max(src.stat().st_mtime, _headers_mtime()):
        inc = ["-I", str(REPO / "include"), "-I", "/usr/local/cuda/include"]
        if os.environ.get("FLOWMC_XLA_FFI_INCLUDE"):
            inc += ["-I", os.environ["FLOWMC_XLA_FFI_INCLUDE"]]
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", *inc, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {src}:\n{r.stdout}{r.stderr}")
        if verbose:
            print(f"[flowmc_b200.build] compiled {src.relative_to(CSRC)}")
    return obj


def _compile(src: Path, hdr_m: float, verbose: bool) -> tuple[Path, str]:
    rel = src.relative_to(CSRC)
    obj = BUILD / rel.with_suffix(".o")
    obj.parent.mkdir(parents=True, exist_ok=True)
    log = ""
    if (not obj.exists()) or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m):
        # flow kernels: no implicit FMA contraction, so the spline arithmetic rounds like the reference's
        # op-by-op float32 evaluation (explicit fmaf() in the GEMM loops is unaffected)
        # (the tensor-core backward only produces gradients, compared at 2e-4: it keeps FMA contraction -- 8 % fewer
        # instructions in its epilogue-bound spline adjoints)
        extra = ["-fmad=false"] if src.name.startswith(("flow", "nf_")) and src.name != "flow_train_tc.cu" else []
        cmd = [NVCC, *ARCH, *CFLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        (obj.with_suffix(".ptxas.log")).write_text(log)
        if verbose:
            print(f"[flowmc_b200.build] compiled {rel}")
    return obj, log


def build(verbose: bool = True, jobs: int | None = None) -> Path:
    BUILD.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)
    hdr_m = _headers_mtime()
    srcs = sources()
    jobs = jobs or min(len(srcs), os.cpu_count() or 4)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        objs = [o for o, _ in ex.map(lambda s: _compile(s, hdr_m, verbose), srcs)]
        test_objs = [o for o, _ in ex.map(lambda s: _compile(s, hdr_m, verbose), test_sources())]
    objs += [_compile_host(p, verbose) for p in sorted(CSRC.glob("*.cc"))]
    newest = max(o.stat().st_mtime for o in objs)
    if (not LIB.exists()) or LIB.stat().st_mtime < newest:
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
        if verbose:
            print(f"[flowmc_b200.build] linked {LIB}")
    if test_objs and ((not TEST_LIB.exists()) or TEST_LIB.stat().st_mtime < max(
            [LIB.stat().st_mtime] + [o.stat().st_mtime for o in test_objs])):
        cmd = [NVCC, *ARCH, "-shared", "-o", str(TEST_LIB), *map(str, test_objs), "-L", str(LIBDIR), "-lflowmc_b200",
               "-Xlinker", "-rpath=$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
        if verbose:
            print(f"[flowmc_b200.build] linked {TEST_LIB}")
    return LIB


def build_plugin(src: str | os.PathLike, out: str | os.PathLike, verbose: bool = False) -> Path:
    """Compile a user target plugin (.cu including include/flowmc_target.cuh) into its own .so."""
    out = Path(out)
    out.parent.mkdir(parents=True, exist_ok=True)
    cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
           "-I", str(REPO / "include"), "-shared", str(src), "-o", str(out),
           "-L", str(LIBDIR), "-lflowmc_b200", "-Xlinker", f"-rpath={LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"plugin build failed:\n{r.stdout}{r.stderr}")
    if verbose:
        print(r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    build(verbose=True)
    print(LIB)
