"""tcgen05 building blocks (csrc/tc_common.cuh): a 128 x N x K GEMM with A in tensor memory, B as packed
SWIZZLE_128B K-major stages, kind::tf32 with 1 and 3 terms, against float64 matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 32), (128, 128), (16, 32), (112, 64), (208, 96), (256, 64)])
@pytest.mark.parametrize("terms", [1, 3])
def test_tc_gemm(cuda, N, K, terms):
    from flowmc_b200._lib import check, load_test_lib
    lib = load_test_lib()
    r = np.random.default_rng(N * 1000 + K)
    A = r.standard_normal((128, K)).astype(np.float32)
    W = (r.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    Ad, Wd = torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    scratch = torch.zeros((K // 32) * 2 * N * 32, device="cuda")
    check(lib.flowmc_test_tc_gemm(Ad.data_ptr(), Wd.data_ptr(), N, K, terms, out.data_ptr(), scratch.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    err = np.abs(out.cpu().numpy() - ref).max()
    scale = np.abs(ref).max()
    tol = (2e-6 if terms == 3 else 3e-3) * scale
    assert err <= tol, f"N={N} K={K} terms={terms}: max err {err:.3e} vs tol {tol:.3e}"
    if terms == 1:
        assert err > 1e-6 * scale      # really TF32 (not a silent fp32 path)


@pytest.mark.parametrize("N,K", [(128, 32), (128, 128), (112, 128), (32, 64), (256, 64)])
@pytest.mark.parametrize("terms", [1, 3])
def test_tc_gemm_cta_pair(cuda, N, K, terms):
    """cta_group::2: a 256 x N x K product on two SMs, each streaming half of W."""
    from flowmc_b200._lib import check, load_test_lib
    lib = load_test_lib()
    r = np.random.default_rng(N * 1000 + K + 7)
    A = r.standard_normal((256, K)).astype(np.float32)
    W = (r.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    Ad, Wd = torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda()
    out = torch.full((256, N), float("nan"), device="cuda")
    scratch = torch.zeros((K // 32) * 2 * N * 32, device="cuda")
    check(lib.flowmc_test_tc_gemm_pair(Ad.data_ptr(), Wd.data_ptr(), N, K, terms, out.data_ptr(), scratch.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    err = np.abs(out.cpu().numpy() - ref).max()
    scale = np.abs(ref).max()
    tol = (2e-6 if terms == 3 else 3e-3) * scale
    assert err <= tol, f"N={N} K={K} terms={terms}: max err {err:.3e} vs tol {tol:.3e}"
