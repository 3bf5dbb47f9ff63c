"""Development probe (GPU): CUDA-graph-replayed time of sub-pieces of the step (B=32, n=20, D=400, R=36)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cliora_b200.net.cliora import DioraMLP
from cliora_b200.net.diora import DioraMLP as TextDiora
B, n, D, R = int(os.environ.get('B', 32)), 20, 400, 36
C = n * (n + 1) // 2
def timeit(fn, reps=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for chains in (1, 2):
    for vl in (False, True):
        m = (DioraMLP(D) if vl else TextDiora(D)).cuda(); m.chains = chains; m.train()
        x = torch.randn(B, n, D, device='cuda', requires_grad=True)
        obj = (0.05 * torch.randn(B, R, D, device='cuda')).requires_grad_() if vl else None
        gi, go = torch.randn(B, C, D, device='cuda'), torch.randn(B, C, D, device='cuda')
        gs = torch.randn(B, C, 1, device='cuda')
        def fwd():
            with torch.no_grad():
                m(x, x, obj, obj) if vl else m(x, x)
        def fwd_in_only():
            m.outside = False
            with torch.no_grad():
                m(x, x, obj, obj) if vl else m(x, x)
            m.outside = True
        def fwdbwd():
            for p in m.parameters(): p.grad = None
            x.grad = None
            m(x, x, obj, obj) if vl else m(x, x)
            ((m.inside_h * gi).sum() + (m.outside_h * go).sum() + (m.inside_s * gs).sum() + (m.outside_s * gs).sum()).backward()
        print('chains=%d %-6s inside fwd %.2f ms | inside+outside fwd %.2f ms | fwd+bwd %.2f ms' % (
            chains, 'CLIORA' if vl else 'DIORA', timeit(fwd_in_only), timeit(fwd), timeit(fwdbwd)))
