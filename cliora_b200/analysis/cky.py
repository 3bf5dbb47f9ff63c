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


def spans(diora):
    """int32 [B, n-1, 2]: (start, end) inclusive spans of every constituent of the decoded trees, computed on
    the device from the backpointer table (the reference goes tree -> str -> actions -> spans on the host)."""
    bp, _ = backpointers(diora)
    B, n = diora._run.B, diora._run.n
    if n < 2:
        return torch.empty(B, 0, 2, device=bp.device, dtype=torch.int32)
    out = torch.empty(B, n - 1, 2, device=bp.device, dtype=torch.int32)
    scratch = torch.empty(3 * B * n, device=bp.device, dtype=torch.int32)
    with torch.cuda.device(bp.device):
        check(_lib.lib().cliora_tree_spans(B, n, ptr(bp), ptr(out), ptr(scratch), _lib.stream()), 'cliora_tree_spans')
    return out


def span_f1(diora, gold_spans):
    """Per-sentence (tp, fp, fn, F1) of the decoded trees against gold spans, on the device.

    ``gold_spans``: list (one per sentence) of lists of (start, end) like ``batch_map['GT']`` (scripts/parse.py:216;
    the last entry of each list is dropped, as there).  Returns a float tensor [B, 4]; corpus F1 follows from the
    column sums exactly as parse.py:283-287."""
    sp = spans(diora)
    B, n = diora._run.B, diora._run.n
    G = max(1, max(len(g) for g in gold_spans))
    gold = torch.zeros(B, G, 2, dtype=torch.int32)
    glen = torch.zeros(B, dtype=torch.int32)
    for b, g in enumerate(gold_spans):
        glen[b] = len(g)
        if len(g):
            gold[b, :len(g)] = torch.tensor([list(x) for x in g], dtype=torch.int32)
    gold, glen = gold.to(sp.device), glen.to(sp.device)
    out = torch.empty(B, 4, device=sp.device, dtype=torch.float32)
    with torch.cuda.device(sp.device):
        check(_lib.lib().cliora_span_f1(B, n, G, ptr(sp), ptr(gold), ptr(glen), ptr(out), _lib.stream()),
              'cliora_span_f1')
    return out


def trees_from_table(bp, n):
    """List of nested-tuple trees from an int32 [B, cells] backpointer tensor: one device->host copy, then the
    CPython helper (csrc/pytrees.c); falls back to the Python recursion if the helper was not built."""
    host = bp.to('cpu', torch.int32).contiguous()
    try:
        from .. import _pytrees
    except ImportError:
        return [tree_from_backpointers(r, n) for r in host.tolist()]
    return _pytrees.build(host.numpy(), host.shape[0], n)


def pack_scalars(scalars, B, n):
    """The reference's per-cell score dict (``net.saved_scalars`` filled by the inside hook,
    analysis/utils.py:78-95: ``scalars[level][pos]`` is ``[B, level]`` or ``[B, level, 1]``) as one flat tensor in
    the CKY kernel's order: level blocks ``[B, n-level, level]`` one after the other, levels 1 .. n-1."""
    parts = []
    for level in range(1, n):
        cells = [scalars[level][pos].reshape(B, level) for pos in range(n - level)]
        parts.append(torch.stack(cells, 1).reshape(-1))
    if not parts:
        return torch.zeros(4, dtype=torch.float32)
    return torch.cat(parts).to(torch.float32).contiguous()


class ParsePredictor(object):
    def __init__(self, net):
        self.net = net

    def batched_cky(self, batch_map, scalars):
        """cky.py:31-99 for callers that bring their own score dict (e.g. ``net.saved_scalars`` after editing it):
        the scores are packed and decoded by the same kernel ``parse_batch`` uses."""
        B, n = batch_map['sentences'].shape[0], batch_map['sentences'].shape[1]
        dev = self.net.device
        scores = pack_scalars(scalars, B, n).to(dev)
        # The kernel subtracts each cell's maximum itself, which is a no-op on hook-made scores (the hook already
        # did it) but would change the result for scores that are not normalised that way: refuse those.
        for level in range(1, n):
            for pos in range(n - level):
                if float(scalars[level][pos].reshape(B, level).max(1)[0].abs().max()) > 1e-6:
                    raise NotImplementedError('batched_cky expects per-cell max-normalised scores (as saved by '
                                              'override_inside_hook); got a cell whose maximum is not 0')
        C = n * (n + 1) // 2
        bp = torch.empty(B, C, device=dev, dtype=torch.int32)
        best = torch.empty(B, C, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_lib.lib().cliora_cky(B, n, ptr(scores), ptr(bp), ptr(best), _lib.stream()), 'cliora_cky')
        return trees_from_table(bp, n)

    def follow_backpointers(self, bp, pair):
        """cky.py:101-109 on the reference's own backpointer structure (``bp[level][pos]`` = pair of (level, pos)
        children, ints at the leaves)."""
        if isinstance(pair, int):
            return pair
        left, right = pair
        return (self.follow_backpointers(bp, bp[left[0]][left[1]]),
                self.follow_backpointers(bp, bp[right[0]][right[1]]))

    def parse_batch(self, batch_map):
        n = batch_map['sentences'].shape[1]
        bp, _ = backpointers(self.net)
        return trees_from_table(bp, n)      # the single device->host transfer

    def parse_spans(self, batch_map=None):
        """Predicted constituent spans as a device tensor [B, n-1, 2] (no host tree building)."""
        return spans(self.net)
