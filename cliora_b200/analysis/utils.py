"""Hook installers with the reference's names (cliora/analysis/utils.py:67-95) plus the span helpers
(``get_actions`` / ``get_spans`` / ``get_stats``, analysis/utils.py:3-64, for callers that still score on the host;
the device versions are ``analysis.cky.spans`` / ``span_f1``)."""
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


SHIFT, REDUCE = 0, 1


def get_actions(tree, SHIFT=SHIFT, REDUCE=REDUCE, OPEN='(', CLOSE=')'):
    """Shift/reduce sequence of a bracketed tree string such as ``str(nested_tuple)`` or ``((A B) (C D))``
    (analysis/utils.py:3-25): one SHIFT per terminal token, one REDUCE per closing bracket."""
    actions, prev = [], OPEN
    for ch in tree.strip():
        if ch == CLOSE:
            actions.append(REDUCE)
        elif ch != OPEN and ch != ' ' and (prev == OPEN or prev == ' '):
            actions.append(SHIFT)    # first character of a terminal; a tuple's commas follow a digit or ')' and do not count
        prev = ch
    if actions.count(SHIFT) != actions.count(REDUCE) + 1:
        raise AssertionError('not a binary bracketing: %r' % (tree,))
    return actions


def get_spans(actions, SHIFT=SHIFT, REDUCE=REDUCE):
    """(start, end) inclusive word spans of the constituents in reduce order (analysis/utils.py:27-48)."""
    spans, stack, word = [], [], 0
    for a in actions:
        if a == SHIFT:
            stack.append((word, word))
            word += 1
        elif a == REDUCE:
            right = stack.pop()
            left = stack.pop()
            node = (left[0], right[1])
            spans.append(node)
            stack.append(node)
    return spans


def get_stats(span1, span2):
    """(tp, fp, fn) of predicted spans ``span1`` against gold spans ``span2`` (analysis/utils.py:51-64)."""
    tp = sum(1 for s in span1 if s in span2)
    fn = sum(1 for s in span2 if s not in span1)
    return tp, len(span1) - tp, fn
