"""Development probe (GPU): do independent small launches on separate streams overlap inside a CUDA graph?"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200 import _lib as L
lib = L.lib()
N = K = 400
W = torch.randn(N, K).cuda(); b = torch.randn(N).cuda()
Wp = torch.empty(2, N, K, device='cuda'); L.check(lib.cliora_split_tf32(L.ptr(W), W.numel(), L.ptr(Wp), L.stream()), 's')
def make(M):
    A = torch.randn(M, K).cuda(); Ap = torch.empty(2, M, K, device='cuda')
    L.check(lib.cliora_split_tf32(L.ptr(A), A.numel(), L.ptr(Ap), L.stream()), 's')
    return A, Ap, torch.empty(M, N, device='cuda')
def run(kind, M, nbranch, per_branch):
    bufs = [make(M) for _ in range(nbranch)]
    def launch(i, st):
        A, Ap, C = bufs[i]
        if kind == 'tc':
            return lib.cliora_tc_linear(M, N, K, L.ptr(Ap), L.ptr(Wp), L.ptr(b), 1, L.ptr(C), st)
        return lib.cliora_linear(M, N, K, L.ptr(A), L.ptr(W), L.ptr(b), 1, L.ptr(C), st)
    main = torch.cuda.Stream(); sides = [torch.cuda.Stream() for _ in range(nbranch)]
    with torch.cuda.stream(main):
        for i in range(nbranch): L.check(launch(i, main.cuda_stream), kind)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            for i, s in enumerate(sides):
                s.wait_stream(main)
                for _ in range(per_branch): L.check(launch(i, s.cuda_stream), kind)
            for s in sides: main.wait_stream(s)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for _ in range(5): g.replay()
        e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 5
for kind in ('tc', 'simt'):
    for M in (608, 3200):
        t1 = run(kind, M, 1, 20); t2 = run(kind, M, 2, 20); t4 = run(kind, M, 4, 20)
        print('%-4s M=%5d: 1 branch x20 = %7.1f us | 2 branches x20 = %7.1f us | 4 branches x20 = %7.1f us  (perfect overlap would stay at %.0f)' % (kind, M, t1, t2, t4, t1))
