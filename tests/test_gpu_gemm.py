"""fp32 GEMM primitives of the library against torch (float64 reference)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _lib():
    from cliora_b200 import _lib
    return _lib


@pytest.mark.parametrize('M,N,K,act', [(1, 4, 4, 0), (37, 400, 400, 1), (640, 400, 400, 2), (3200, 1200, 400, 0),
                                       (100, 101, 48, 0), (65, 36, 24, 0), (40000, 400, 400, 1)])
def test_linear(M, N, K, act):
    L = _lib()
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).cuda()
    if act == 2:
        A = A * 0.05   # keep tanh out of saturation so the check is about the GEMM, not tanh'(x) ~ 0
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    C = torch.empty(M, N, device='cuda')
    L.check(L.lib().cliora_linear(M, N, K, L.ptr(A), L.ptr(W), L.ptr(b), act, L.ptr(C), L.stream()), 'linear')
    ref = A.double() @ W.double().t() + b.double()
    ref = {0: ref, 1: torch.relu(ref), 2: torch.tanh(ref)}[act]
    assert rel_err(C, ref) < (1e-5 if act == 2 else 2e-6)   # tanhf is a few ulp


@pytest.mark.parametrize('M,N,K', [(5, 8, 12), (3200, 400, 400), (640, 400, 1200), (77, 48, 101)])
def test_matmul_nn(M, N, K):
    L = _lib()
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, K, generator=g).cuda()
    Bm = torch.randn(K, N, generator=g).cuda()
    C0 = torch.randn(M, N, generator=g).cuda()
    C = C0.clone()
    L.check(L.lib().cliora_matmul_nn(M, N, K, L.ptr(A), L.ptr(Bm), L.ptr(C), 1, L.stream()), 'nn')
    assert rel_err(C, C0.double() + A.double() @ Bm.double()) < 2e-6


@pytest.mark.parametrize('M,Ka,Kb', [(7, 8, 12), (6720, 400, 400), (42560, 400, 400), (300, 36, 48), (0, 8, 8)])
def test_matmul_tn(M, Ka, Kb):
    L = _lib()
    g = torch.Generator().manual_seed(M + 1)
    A = torch.randn(max(M, 1), Ka, generator=g).cuda()[:M]
    Bm = torch.randn(max(M, 1), Kb, generator=g).cuda()[:M]
    C = torch.full((Ka, Kb), 7.0, device='cuda')
    scratch = torch.empty(int(L.lib().cliora_matmul_tn_scratch_floats(M, Ka, Kb)) + 8, device='cuda')
    A2 = A.contiguous() if M else torch.zeros(1, Ka, device='cuda')
    B2 = Bm.contiguous() if M else torch.zeros(1, Kb, device='cuda')
    L.check(L.lib().cliora_matmul_tn(M, Ka, Kb, L.ptr(A2), L.ptr(B2), L.ptr(C), 0, L.ptr(scratch), L.stream()), 'tn')
    ref = A.double().t() @ Bm.double()
    if M == 0:
        assert C.abs().max().item() == 0
    else:
        assert rel_err(C, ref) < 2e-6


@pytest.mark.parametrize('M,N,K,act', [(128, 80, 32, 0), (128, 80, 400, 0), (37, 400, 400, 1), (3200, 400, 400, 1),
                                       (12160, 400, 400, 1), (640, 1200, 400, 0), (500, 400, 1200, 0)])
def test_tc_linear_3xtf32(M, N, K, act):
    """tcgen05 3xTF32 GEMM: fp32-grade accuracy (single-pass TF32 would be ~1e-3)."""
    L = _lib()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    Ap = torch.empty(2, M, K, device='cuda')
    Wp = torch.empty(2, N, K, device='cuda')
    L.check(L.lib().cliora_split_tf32(L.ptr(A), A.numel(), L.ptr(Ap), L.stream()), 'split')
    L.check(L.lib().cliora_split_tf32(L.ptr(W), W.numel(), L.ptr(Wp), L.stream()), 'split')
    assert torch.equal(Ap[0] + Ap[1], A)
    C = torch.full((M, N), float('nan'), device='cuda')
    L.check(L.lib().cliora_tc_linear(M, N, K, L.ptr(Ap), L.ptr(Wp), L.ptr(b), act, L.ptr(C), L.stream()), 'tc_linear')
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + b.double()
    if act == 1:
        ref = torch.relu(ref)
    err = rel_err(C, ref)
    assert err < 5e-6, err   # K=1200: 3.2e-6 (truncating TMEM accumulation), cuBLAS fp32 is 1.3e-6


@pytest.mark.parametrize('M,Ka,Kb', [(32, 128, 224), (64, 400, 400), (6720, 400, 400), (42560, 400, 400),
                                     (1000, 64, 48), (5, 400, 400)])
def test_tc_matmul_tn_3xtf32(M, Ka, Kb):
    """Weight-gradient GEMM on tcgen05 with MN-major operands and split-K."""
    L = _lib()
    g = torch.Generator().manual_seed(M + Ka)
    A = torch.randn(M, Ka, generator=g).cuda()
    Bm = torch.randn(M, Kb, generator=g).cuda()
    Ap = torch.empty(2, M, Ka, device='cuda')
    Bp = torch.empty(2, M, Kb, device='cuda')
    L.check(L.lib().cliora_split_tf32(L.ptr(A), A.numel(), L.ptr(Ap), L.stream()), 'split')
    L.check(L.lib().cliora_split_tf32(L.ptr(Bm), Bm.numel(), L.ptr(Bp), L.stream()), 'split')
    C0 = torch.randn(Ka, Kb, generator=g).cuda()
    C = C0.clone()
    scratch = torch.empty(int(L.lib().cliora_tc_matmul_tn_scratch_floats(M, Ka, Kb)) + 8, device='cuda')
    L.check(L.lib().cliora_tc_matmul_tn(M, Ka, Kb, L.ptr(Ap), L.ptr(Bp), L.ptr(C), 1, L.ptr(scratch), L.stream()), 'tn')
    torch.cuda.synchronize()
    ref = C0.double() + A.double().t() @ Bm.double()
    err = rel_err(C, ref)
    assert err < 1e-5, err   # 42k rows: 5.7e-6 (truncating TMEM accumulation over ~300 k-steps per split)


@pytest.mark.parametrize('cfg', [1, 2, 3])
@pytest.mark.parametrize('M,N,K', [(300, 400, 400), (1000, 1200, 400), (129, 144, 64)])
def test_tc_linear_both_tile_configs(cfg, M, N, K):
    """Narrow (80-column) and wide (256-column, runtime UMMA N on the ragged tile) configurations agree with fp64."""
    L = _lib()
    L.lib().cliora_debug_set(2, cfg)
    try:
        test_tc_linear_3xtf32(M, N, K, 1)
    finally:
        L.lib().cliora_debug_set(2, 0)
