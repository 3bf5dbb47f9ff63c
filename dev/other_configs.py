"""Development probe (GPU): throughput of the BASELINE configs other than config[1] (graph-free, CUDA events)."""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cliora_b200.net.diora import DioraMLP
from cliora_b200.analysis.cky import ParsePredictor
res = {}
def ev(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# c1: DIORA-MLP chart fwd+bwd, B=32, n=20
m = DioraMLP(400).cuda(); x = torch.randn(32, 20, 400, device='cuda', requires_grad=True)
def c1():
    for p in m.parameters(): p.grad = None
    m(x, x); (m.outside_h[:, :20].sum() + m.inside_s.sum() + m.outside_s.sum()).backward()
t = ev(c1, 10); res['c1_diora_fwd_bwd_B32_n20'] = dict(ms=t, sentences_per_s=32e3 / t)
# c3: parse B=256 n=30 (inside pass + CKY + tree extraction on host)
m.eval(); m.outside = False; x3 = torch.randn(256, 30, 400, device='cuda')
pp = ParsePredictor(m)
def c3():
    with torch.no_grad(): m(x3, x3)
    return pp.parse_batch({'sentences': torch.zeros(256, 30, dtype=torch.int64)})
t0 = time.perf_counter(); c3(); torch.cuda.synchronize()
t = ev(c3, 5); res['c3_parse_B256_n30'] = dict(ms=t, sentences_per_s=256e3 / t)
# c4: n=64, B=16 fwd+bwd
m.train(); m.outside = True; x4 = torch.randn(16, 64, 400, device='cuda', requires_grad=True)
def c4():
    for p in m.parameters(): p.grad = None
    m(x4, x4); (m.outside_h[:, :64].sum() + m.inside_s.sum() + m.outside_s.sum()).backward()
t = ev(c4, 3); res['c4_diora_fwd_bwd_B16_n64'] = dict(ms=t, sentences_per_s=16e3 / t)
print(json.dumps(res))
