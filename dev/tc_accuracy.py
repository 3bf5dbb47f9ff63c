"""Development probe (GPU): accuracy of the tcgen05 3xTF32 GEMM under different accumulation schemes."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200 import _lib as L
lib = L.lib()
for scale_name, mk in (('randn', lambda g, *s: torch.randn(*s, generator=g)),
                       ('relu-like', lambda g, *s: torch.relu(torch.randn(*s, generator=g)))):
    for K in (32, 128, 400, 1200):
        M, N = 1024, 400
        g = torch.Generator().manual_seed(K)
        A = mk(g, M, K).cuda(); W = torch.randn(N, K, generator=g).cuda()
        Ap = torch.empty(2, M, K, device='cuda'); Wp = torch.empty(2, N, K, device='cuda')
        L.check(lib.cliora_split_tf32(L.ptr(A), A.numel(), L.ptr(Ap), L.stream()), 's')
        L.check(lib.cliora_split_tf32(L.ptr(W), W.numel(), L.ptr(Wp), L.stream()), 's')
        ref = A.double() @ W.double().t()
        f32 = (A @ W.t()).double()
        big = ref.abs() > 0.5 * ref.abs().mean()
        line = '%-9s K=%4d  fp32-cublas %.2e |' % (scale_name, K, ((f32 - ref).abs().max() / ref.abs().max()).item())
        for mode in (0, 2, 3):
            lib.cliora_debug_set(0, mode)
            C = torch.empty(M, N, device='cuda')
            L.check(lib.cliora_tc_linear(M, N, K, L.ptr(Ap), L.ptr(Wp), None, 0, L.ptr(C), L.stream()), 'tc')
            torch.cuda.synchronize()
            d = C.double() - ref
            line += ' mode%d max %.2e bias %.2e |' % (mode, (d.abs().max() / ref.abs().max()).item(),
                                                     ((d / ref)[big]).mean().item())
        print(line)
lib.cliora_debug_set(0, 2)
