"""CKY decoding: drop-in for ``cliora.analysis.cky.ParsePredictor`` (cliora/analysis/cky.py:1-110).

The Viterbi recursion over split scores runs in one kernel launch per batch (csrc/cky_kernels.cuh);
only the final backpointer table crosses to the host (the reference syncs once per chart cell).
"""
import torch

from .. import _lib
from .._lib import check, ptr


def backpointers(diora):
    """int32 [B, cells] best split per cell (first max wins, -1 at leaves) + Viterbi scores."""
    run = diora._run
    if run is None:
        raise RuntimeError('run the model forward before parsing')
    B, n = run.B, run.n
    C = n * (n + 1) // 2
    dev = run.ws.device
    bp = torch.empty(B, C, device=dev, dtype=torch.int32)
    best = torch.empty(B, C, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        for b0, b1, ws, lay in run.parts:      # one launch per forward chain (each owns its split-score region)
            scores = ws[lay.Ein: lay.Ein + max(int(lay.rows_in), 1)]
            check(_lib.lib().cliora_cky(b1 - b0, n, ptr(scores), ptr(bp) + b0 * C * 4, ptr(best) + b0 * C * 4,
                                        _lib.stream()), 'cliora_cky')
    return bp, best


def tree_from_backpointers(row, n):
    """cky.py:101-109: nested tuples of word positions."""
    off = [l * n - l * (l - 1) // 2 for l in range(n)]

    def rec(level, pos):
        if level == 0:
            return pos
        k = row[off[level] + pos]
        return (rec(k, pos), rec(level - 1 - k, pos + k + 1))

    return rec(n - 1, 0)


class ParsePredictor(object):
    def __init__(self, net):
        self.net = net

    def parse_batch(self, batch_map):
        n = batch_map['sentences'].shape[1]
        bp, _ = backpointers(self.net)
        rows = bp.cpu().tolist()           # the single device->host transfer
        return [tree_from_backpointers(r, n) for r in rows]
