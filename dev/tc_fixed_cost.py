"""Development probe (GPU): fixed per-launch cost of the tcgen05 GEMM - one CTA / one wave, K from 32 up."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200 import _lib as L
lib = L.lib()
def bench(M, N, K, mode=2, reps=40):
    A = torch.randn(M, K).cuda(); W = torch.randn(N, K).cuda()
    Ap = torch.empty(2, M, K, device='cuda'); Wp = torch.empty(2, N, K, device='cuda')
    L.check(lib.cliora_split_tf32(L.ptr(A), A.numel(), L.ptr(Ap), L.stream()), 's')
    L.check(lib.cliora_split_tf32(L.ptr(W), W.numel(), L.ptr(Wp), L.stream()), 's')
    C = torch.empty(M, N, device='cuda')
    lib.cliora_debug_set(0, mode)
    fn = lambda: lib.cliora_tc_linear(M, N, K, L.ptr(Ap), L.ptr(Wp), None, 0, L.ptr(C), L.stream())
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): L.check(fn(), 'tc')
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): L.check(fn(), 'tc')
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): g.replay()
        e1.record(s); torch.cuda.synchronize()
    lib.cliora_debug_set(0, 2)
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)
def empty_kernel(reps=40):
    x = torch.zeros(1024, device='cuda')
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        x.add_(1); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): x.add_(1)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): g.replay()
        e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)
print('tiny elementwise kernel in a graph chain: %.2f us/launch' % empty_kernel())
for M, N in ((128, 80), (128 * 25, 400), (128 * 148, 80)):
    for K in (32, 64, 128, 256, 400, 800):
        print('M=%5d N=%3d K=%4d  mode1 %6.2f us   mode2 %6.2f us' % (M, N, K, bench(M, N, K, 1), bench(M, N, K, 2)))
