"""Hook installers with the reference's names (cliora/analysis/utils.py:67-95) plus the span helpers."""
import types


def override_init_with_batch(var):
    init_with_batch = var.init_with_batch

    def func(self, *args, **kwargs):
        init_with_batch(*args, **kwargs)
        self.saved_scalars = {i: {} for i in range(self.length)}
        self.saved_scalars_out = {i: {} for i in range(self.length)}

    var.init_with_batch = types.MethodType(func, var)


def override_inside_hook(var):
    """Saves s - max_k s per (level, pos) like the reference.  ParsePredictor here does not need it
    (the CKY kernel reads the raw scores and subtracts the max itself) but callers may."""

    def func(self, level, h, c, s):
        s = s - s.max(2, keepdim=True)[0]
        for pos in range(self.length - level):
            self.saved_scalars[level][pos] = s[:, pos, :]

    var.inside_hook = types.MethodType(func, var)


def get_spans_from_tree(tree):
    """Spans (start, end) inclusive of every internal node of a nested-tuple tree (analysis/utils.py:27-48)."""
    spans = []

    def rec(t):
        if isinstance(t, int):
            return t, t
        l0, _ = rec(t[0])
        _, r1 = rec(t[1])
        spans.append((l0, r1))
        return l0, r1

    rec(tree)
    return spans
