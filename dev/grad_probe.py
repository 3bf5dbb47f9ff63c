"""Development probe (GPU): gradient errors vs the float64 oracle for several shapes, tc and SIMT GEMM paths."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import rel_err
from test_gpu_chart import _oracle_run, _fill, _grads
from cliora_b200 import _lib
def run(B, n, D, R, share, simt):
    if R:
        from cliora_b200.net.cliora import DioraMLP
    else:
        from cliora_b200.net.diora import DioraMLP
    P0, x, obj, keep, ct, ref64 = _oracle_run(torch.float64, B, n, D, R, share)
    _, _, _, _, _, ref32 = _oracle_run(torch.float32, B, n, D, R, share)
    pre_in, pre_out = ref64.pop('pre'); ref32.pop('pre')
    _lib.lib().cliora_debug_set(1, 1 if simt else 0)
    m = DioraMLP(D, share=share).cuda(); _fill(m, P0)
    xc = x.cuda().requires_grad_(); oc = obj.cuda().requires_grad_() if R else None
    m.train()
    if R:
        m.set_dropout_mask(keep.cuda()); m(xc, xc, oc, oc)
    else:
        m(xc, xc)
    fw = max(rel_err(getattr(m, k), ref64[k]) for k in ct)
    flips = 0
    for outside, pre in ((False, pre_in), (True, pre_out)):
        for level, (u1, u2) in pre.items():
            flips += ((m._run.split_z(level, outside).cpu() > 0) != (u1 > 0)).sum().item()
            flips += ((m._run.split_h(level, outside).cpu() > 0) != (u2 > 0)).sum().item()
    print('  ReLU mask flips vs float64 oracle:', flips)
    sum((getattr(m, k) * ct[k].cuda()).sum() for k in ct).backward()
    mine = {'grad_x': xc.grad}
    if R: mine['grad_obj'] = oc.grad
    for k, v in _grads(m).items():
        if not (share and k.startswith('outside_')): mine['grad:' + k] = v
    worst = max(mine, key=lambda k: rel_err(mine[k], ref64[k]) / max(1e-4, 2 * rel_err(ref32[k], ref64[k])))
    print('B=%d n=%2d D=%d R=%2d share=%d %s fwd %.1e | grad_x %.1e (floor %.1e) | worst %s %.1e (floor %.1e)' % (
        B, n, D, R, share, 'SIMT' if simt else 'TC  ', fw, rel_err(mine['grad_x'], ref64['grad_x']),
        rel_err(ref32['grad_x'], ref64['grad_x']), worst.replace('grad:', '')[-28:], rel_err(mine[worst], ref64[worst]),
        rel_err(ref32[worst], ref64[worst])))
    _lib.lib().cliora_debug_set(1, 0)
for cfg in [(3, 9, 400, 36, True), (4, 10, 400, 36, True), (3, 9, 400, 0, True), (2, 9, 64, 8, True), (3, 5, 400, 36, True)]:
    for simt in (False, True):
        run(*cfg, simt)
