"""In-kernel timeline of one fused level launch (dev tool): clock64 stamps recorded by designated threads of every
CTA, printed as phase durations in microseconds (SM clock from nvidia-smi max, 1.965 GHz on the pool's B200s).

    python dev/level_timeline.py --level 10 [--outside] [--batch 32] [--text]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200 import _lib  # noqa: E402

NAMES = {0: 'start', 1: 'init+cluster sync', 2: 'epi start', 3: 'score loads done', 4: 'xchg0 done', 5: 'softmax done',
         6: 'tmem_full seen', 7: 'tmem->Y/stage done', 8: 'cell sums + ss put', 9: 'xchg1 done', 10: 'normalise stored',
         11: 'obj staged (bar2)', 12: 'logits put', 13: 'xchg2 done', 14: 'att softmax', 15: 'a2 + ss2 put',
         16: 'xchg3 done', 17: 'epi end', 18: 'exit', 20: 'mma first full', 21: 'mma last commit', 22: 'prod start',
         23: 'prod loop end', 24: 'prod obj staged'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--level', type=int, default=10)
    ap.add_argument('--outside', action='store_true')
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--length', type=int, default=20)
    ap.add_argument('--text', action='store_true')
    ap.add_argument('--bwd', action='store_true')
    ap.add_argument('--ghz', type=float, default=1.965)
    ap.add_argument('--precision', default='fp32')
    ap.add_argument('--debug-set', default='')
    args = ap.parse_args()
    B, n, D, R = args.batch, args.length, 400, 0 if args.text else 36
    if R:
        from cliora_b200.net.cliora import DioraMLP
    else:
        from cliora_b200.net.diora import DioraMLP
    torch.manual_seed(0)
    m = DioraMLP(D).cuda()
    m.chains = 1
    m.precision = args.precision
    for kv in filter(None, args.debug_set.split(',')):
        k, v = kv.split('=')
        _lib.lib().cliora_debug_set(int(k), int(v))
    x = torch.randn(B, n, D).cuda()
    obj = 0.05 * torch.randn(B, max(R, 1), D).cuda()
    L = _lib.lib()
    dbg = torch.zeros(4096, 128, dtype=torch.int64, device='cuda')
    for it in range(3):
        if it == 2:
            L.cliora_debug_ptr(0, dbg.data_ptr())
            L.cliora_debug_set(8, args.level + 1)
            L.cliora_debug_set(9, (2 if args.bwd else 0) + (1 if args.outside else 0))
        if args.bwd:
            xg = x.clone().requires_grad_()
            m(xg, xg, obj, obj) if R else m(xg, xg)
            (m.inside_h.sum() + m.outside_h.sum() + m.inside_s.sum() + m.outside_s.sum()).backward()
        else:
            with torch.no_grad():
                m(x, x, obj, obj) if R else m(x, x)
        torch.cuda.synchronize()
    L.cliora_debug_set(8, 0)
    L.cliora_debug_ptr(0, None)
    t = dbg.cpu()
    used = (t[:, 0] != 0).nonzero().flatten()
    print('ctas recorded:', len(used))
    g0 = t[used, 30].min().item()
    span = (t[used, 31].max().item() - g0) / 1e3
    print('launch span (globaltimer): %.1f us; cta start offsets us: min %.1f max %.1f' % (
        span, (t[used, 30].min().item() - g0) / 1e3, (t[used, 30].max().item() - g0) / 1e3))
    life = (t[used, 31] - t[used, 30]).double() / 1e3
    print('cta lifetime us: mean %.1f min %.1f max %.1f' % (life.mean(), life.min(), life.max()))
    for cta in [int(used[0])]:
        row = t[cta]
        base = row[0].item()
        print('--- cta', cta)
        for k in sorted(NAMES):
            if row[k].item() and not args.bwd:
                print('  %-22s %8.2f us' % (NAMES[k], (row[k].item() - base) / (args.ghz * 1e3)))
        us = lambda k: (row[k].item() - base) / (args.ghz * 1e3)
        if args.bwd:
            for k, nm in [(6, 'vl: staging issued'), (7, 'vl: ga2 done'), (8, 'vl: regions staged'), (9, 'vl: region dots'), (10, 'vl: cell done'), (1, 'prologue done'), (20, 'mma first full'), (2, 'db2 sums done'), (21, 'mma last full'), (3, 'tmem_full + sync'),
                          (4, 'GZ staged in smem'), (5, 'scatter done'), (18, 'exit')]:
                print('  %-22s %8.2f us' % (nm, us(k)))
            continue
        for i in range(4):
            print('  xform kb=%d: top %.2f raw_full %.2f math %.2f emptyA %.2f st_done %.2f end %.2f' % tuple([4 + i] + [us(32 + 6 * i + q) for q in range(6)]))
        for i in range(4):
            print('  copy kb=%d: raw_empty %.2f waited %.2f flushed %.2f copies %.2f issued %.2f' % (
                4 + i, us(56 + 2 * i), us(64 + 4 * i), us(65 + 4 * i), us(66 + 4 * i), us(57 + 2 * i)))
        for i in range(4):
            print('  mma kb=%d: full %.2f issued %.2f' % (3 + i, us(48 + 2 * i), us(49 + 2 * i)))


if __name__ == '__main__':
    main()
