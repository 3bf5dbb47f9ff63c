"""Dev tool: start / end (globaltimer) of every fused level kernel of one pass, to see the time BETWEEN the level kernels
(= the per-cell GEMM in between plus two launch gaps).  python dev/level_gaps.py [--outside] [--bwd] [--batch 32] [--graph]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument('--outside', action='store_true'); ap.add_argument('--bwd', action='store_true')
ap.add_argument('--batch', type=int, default=32); ap.add_argument('--text', action='store_true')
ap.add_argument('--graph', action='store_true', help='replay the pass from a CUDA graph (what the trainer does)')
args = ap.parse_args()
B, n, D, R = args.batch, 20, 400, 0 if args.text else 36
if R:
    from cliora_b200.net.cliora import DioraMLP
else:
    from cliora_b200.net.diora import DioraMLP
torch.manual_seed(0)
m = DioraMLP(D).cuda(); m.chains = 1
x = torch.randn(B, n, D).cuda(); obj = 0.05 * torch.randn(B, max(R, 1), D).cuda()
L = _lib.lib()
dbg = torch.zeros(32 * 1024, 128, dtype=torch.int64, device='cuda')

def run():
    if args.bwd:
        xg = x.clone().requires_grad_()
        m(xg, xg, obj, obj) if R else m(xg, xg)
        (m.inside_h.sum() + m.outside_h.sum() + m.inside_s.sum() + m.outside_s.sum()).backward()
    else:
        with torch.no_grad():
            m(x, x, obj, obj) if R else m(x, x)

for _ in range(3):
    run()
torch.cuda.synchronize()
L.cliora_debug_ptr(0, dbg.data_ptr()); L.cliora_debug_set(8, 99)
L.cliora_debug_set(9, (2 if args.bwd else 0) + (1 if args.outside else 0))
if args.graph:
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    dbg.zero_(); g.replay(); g.replay()
else:
    torch.cuda._sleep(int(0.05 * 1.9e9))
    run()
torch.cuda.synchronize()
L.cliora_debug_set(8, 0); L.cliora_debug_ptr(0, None)
t = dbg.cpu().view(32, 1024, 128)
prev_end = None
order = range(1, n) if not (args.outside ^ False) else range(n - 2, -1, -1)
if args.bwd:
    order = range(0, n - 1) if args.outside else range(n - 1, 0, -1)
for lvl in order:
    rows = t[lvl]; used = rows[:, 30] != 0
    if not used.any(): continue
    st, en = rows[used, 30].min().item(), rows[used, 31].max().item()
    print('level %2d  ctas %3d  kernel %6.1f us   since previous level kernel ended %6.1f us' % (
        lvl, int(used.sum()), (en - st) / 1e3, (st - prev_end) / 1e3 if prev_end else 0.0))
    prev_end = en
