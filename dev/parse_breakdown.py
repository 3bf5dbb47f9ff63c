"""Development probe (GPU): where config[2] (parse B=256, n=30) spends its time."""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200.net.diora import DioraMLP
from cliora_b200.analysis.cky import ParsePredictor, backpointers, spans, tree_from_backpointers
m = DioraMLP(400).cuda().eval(); m.outside = False
x = torch.randn(256, 30, 400, device='cuda'); pp = ParsePredictor(m)
def ev(fn, reps=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
def fwd():
    with torch.no_grad(): m(x, x)
res = {'inside_fwd_ms': ev(fwd)}
res['fwd+cky_ms'] = ev(lambda: (fwd(), backpointers(m)))
res['fwd+cky+device_spans_ms'] = ev(lambda: (fwd(), spans(m)))
res['fwd+parse_batch_host_trees_ms'] = ev(lambda: (fwd(), pp.parse_batch({'sentences': torch.zeros(256, 30, dtype=torch.int64)})))
bp, _ = backpointers(m); rows = bp.cpu().tolist()
t0 = time.perf_counter()
for _ in range(5): [tree_from_backpointers(r, 30) for r in rows]
res['host_tree_build_only_ms'] = (time.perf_counter() - t0) / 5 * 1e3
for ch in (1, 2, 4, 8):
    m.chains = ch
    res['inside_fwd_chains%d_ms' % ch] = ev(fwd)
print(json.dumps(res))
