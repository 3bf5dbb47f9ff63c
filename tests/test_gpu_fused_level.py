"""The fused per-level forward kernel (csrc/level_kernels.cuh: gather + compose GEMM + softmax-weighted sums + cell
finalize in one launch per level) against the unfused kernel chain it replaces (split_build -> tcgen05 GEMM ->
cell_aggregate), buffer by buffer: chart tensors, split scores / probabilities, the compose MLP's hidden and output
rows, ReLU bit masks and the per-cell tensors saved for backward.  Parity with the reference itself is what the other
GPU tests check with whichever path is the default; this one localises a disagreement to a buffer."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _run(B, n, D, R, share, fused, train, seed=3):
    from cliora_b200 import _lib
    from oracle import cliora_oracle as O
    from test_gpu_chart import _fill
    if R:
        from cliora_b200.net.cliora import DioraMLP
    else:
        from cliora_b200.net.diora import DioraMLP
    _lib.lib().cliora_debug_set(6, 0 if fused else 1)
    _lib.lib().cliora_debug_set(10, 1)      # have the fused kernel emit the (optional) ReLU bit masks too
    try:
        P0 = O.init_params(D, share=share, seed=seed)
        g = torch.Generator().manual_seed(seed + 1)
        x = torch.randn(B, n, D, generator=g).cuda()
        obj = (0.05 * torch.randn(B, R, D, generator=g)).cuda() if R else None
        keep = (torch.rand(B, O.num_cells(n), R, generator=g) >= 0.1).cuda() if R else None
        m = DioraMLP(D, share=share).cuda()
        m.chains = 1
        _fill(m, P0)
        m.train() if train else m.eval()
        if R and train:
            m.set_dropout_mask(keep)
        with torch.no_grad():
            if R:
                m(x, x, obj, obj)
            else:
                m(x, x)
        torch.cuda.synchronize()
        lay = m._run.layout
        ws = m._run.ws.clone()
        out = {k: getattr(m, k).clone() for k in ('inside_h', 'inside_s', 'outside_h', 'outside_s')}
        return out, ws, lay
    finally:
        _lib.lib().cliora_debug_set(6, 0)
        _lib.lib().cliora_debug_set(10, 0)


SHAPES = [(4, 10, 400, 36, True, True), (3, 8, 400, 0, False, False), (2, 20, 400, 36, True, True),
          (5, 11, 132, 7, False, True), (7, 13, 260, 33, True, False), (2, 33, 64, 5, True, True),
          (1, 3, 36, 1, True, False), (2, 4, 520, 64, True, True), (1, 2, 32, 0, True, False),
          (16, 20, 400, 36, True, True)]


@pytest.mark.parametrize('B,n,D,R,share,train', SHAPES)
def test_fused_level_forward_matches_unfused_chain(B, n, D, R, share, train):
    ref, ws0, lay = _run(B, n, D, R, share, False, train)
    got, ws1, _ = _run(B, n, D, R, share, True, train)
    C = n * (n + 1) // 2
    rows_in, rows_out = int(lay.rows_in), int(lay.rows_out)

    def seg(ws, off, count):
        return ws[off: off + count]
    checks = [('Ein', lay.Ein, rows_in, 2e-5), ('Prin', lay.Prin, rows_in, 2e-5), ('Eout', lay.Eout, rows_out, 2e-5),
              ('Prout', lay.Prout, rows_out, 2e-5), ('Yin', lay.Yin, rows_in * D, 2e-5),
              ('Yout', lay.Yout, rows_out * D, 2e-5), ('nrm_in', lay.nrm_in, B * C, 2e-5),
              ('nrm_out', lay.nrm_out, B * C, 2e-5)]
    if R:
        checks += [('q_in', lay.q_in, B * C * D, 2e-5), ('nrm2_in', lay.nrm2_in, B * C, 2e-5),
                   ('att_in', lay.att_in, B * C * R, 2e-5)]
    for name, off, count, tol in checks:
        if count:
            assert rel_err(seg(ws1, off, count), seg(ws0, off, count)) < tol, name
    # hidden activations: same fp32 adds in both paths, but their inputs (the lower levels' cell vectors) already
    # differ in the last bits between the two paths (summation order of the per-cell reductions), so the pairs agree to
    # rounding and the ReLU masks everywhere except at pre-activations within rounding of zero
    for name, off, rows in (('Zin', lay.Zin, rows_in), ('Zout', lay.Zout, rows_out)):
        if rows:
            z1 = seg(ws1, off, rows * D) + seg(ws1, off + rows * D, rows * D)
            z0 = seg(ws0, off, rows * D) + seg(ws0, off + rows * D, rows * D)
            assert rel_err(z1, z0) < 2e-5, name
            lo1 = seg(ws1, off + rows * D, rows * D)
            assert float(lo1.abs().max()) <= 2.0 ** -10 * float(z1.abs().max()) + 1e-30, name + ' lo part'
    if lay.Mbin >= 0:
        for name, off, count in (('Mbin', lay.Mbin, rows_in * 16), ('Mbout', lay.Mbout, rows_out * 16)):
            if count:
                x = seg(ws1, off, count).view(torch.int32) ^ seg(ws0, off, count).view(torch.int32)
                x = x.view(-1, 16)[:, :4 * ((D + 127) // 128)]      # words of 128-column groups beyond D are never written
                flipped = sum(int(((x >> i) & 1).sum()) for i in range(32))
                assert flipped <= max(2, int(1e-5 * count * 32)), (name, flipped)
    for k in ref:
        assert rel_err(got[k], ref[k]) < 2e-5, k


@pytest.mark.parametrize('B,n,D,R,share', [(4, 10, 400, 36, True), (3, 8, 400, 0, False), (2, 6, 132, 4, True),
                                           (2, 20, 400, 36, True), (2, 6, 512, 4, True), (2, 5, 768, 0, False), (2, 4, 800, 4, True),
                                           (2, 4, 896, 0, True), (3, 5, 36, 0, True), (2, 4, 1024, 0, True), (8, 10, 512, 0, True),
                                           (6, 9, 768, 4, True)])
def test_wide_level_tiles_stay_parity_green(B, n, D, R, share):
    """Levels that do not fit one wave run with wide column slices (two CTAs per tile at D=400, single tensor-memory
    accumulator, three raw stages, two-pass backward epilogue).  Forced on for every level here: forward and every
    gradient against the float64 oracle, exactly like the default geometry."""
    from cliora_b200 import _lib
    from test_gpu_chart import test_chart_vs_oracle_live
    _lib.lib().cliora_debug_set(15, 2)
    try:
        test_chart_vs_oracle_live(B, n, D, R, share)
    finally:
        _lib.lib().cliora_debug_set(15, 0)


@pytest.mark.parametrize('R,wide', [(36, False), (0, False), (36, True), (0, True)])
def test_fused_forward_is_stable_across_runs(R, wide):
    """Every shared-memory stage of the level kernels is handed over through mbarriers.  The same input must give the
    same chart on every run up to the reordering noise of the projection GEMMs' split-K atomics (1e-7 per GEMM): a
    race in one of those hand-overs would show as a wrong element, not as last-bit noise."""
    from cliora_b200 import _lib
    if R:
        from cliora_b200.net.cliora import DioraMLP
    else:
        from cliora_b200.net.diora import DioraMLP
    B, n, D = 8, 12, 400
    torch.manual_seed(5)
    m = DioraMLP(D).cuda().eval()
    m.chains = 1
    x = torch.randn(B, n, D, device='cuda')
    obj = 0.05 * torch.randn(B, max(R, 1), D, device='cuda')
    _lib.lib().cliora_debug_set(15, 2 if wide else 1)
    try:
        outs = []
        for _ in range(6):
            with torch.no_grad():
                m(x, x, obj, obj) if R else m(x, x)
            outs.append([t.clone() for t in (m.inside_h, m.inside_s, m.outside_h, m.outside_s)])
        for o in outs[1:]:
            for a, b in zip(outs[0], o):
                assert rel_err(a, b) < 1e-5
    finally:
        _lib.lib().cliora_debug_set(15, 0)


@pytest.mark.parametrize('B,n,D,R', [(8, 10, 512, 0), (8, 10, 512, 36), (16, 12, 400, 36), (6, 9, 768, 0)])
def test_wide_and_narrow_tiles_agree_on_full_tiles(B, n, D, R):
    """Full 128-row tiles (the small oracle cases leave most of a tile empty): forward tensors and every gradient of the
    wide geometry against the narrow one on the same input -- exercises the column-pass staging of the backward
    epilogue at its largest (a slice of 113..128 columns once overflowed the operand rings there)."""
    from cliora_b200 import _lib
    if R:
        from cliora_b200.net.cliora import DioraMLP
    else:
        from cliora_b200.net.diora import DioraMLP
    torch.manual_seed(11)
    m = DioraMLP(D).cuda().eval()
    m.chains = 1
    x0 = torch.randn(B, n, D, device='cuda')
    obj0 = 0.05 * torch.randn(B, max(R, 1), D, device='cuda')
    C = n * (n + 1) // 2
    ct = [torch.randn(B, C, D, device='cuda'), torch.randn(B, C, 1, device='cuda'),
          torch.randn(B, C, D, device='cuda'), torch.randn(B, C, 1, device='cuda')]
    res = {}
    for name, knob in (('narrow', 1), ('wide', 2)):
        _lib.lib().cliora_debug_set(15, knob)
        try:
            for p in m.parameters():
                p.grad = None
            x = x0.clone().requires_grad_()
            obj = obj0.clone().requires_grad_()
            m(x, x, obj, obj) if R else m(x, x)
            outs = [m.inside_h, m.inside_s, m.outside_h, m.outside_s]
            sum((o * c).sum() for o, c in zip(outs, ct)).backward()
            res[name] = [o.detach().clone() for o in outs] + [x.grad.clone()] + ([obj.grad.clone()] if R else []) + \
                        [p.grad.clone() for p in m.parameters() if p.grad is not None]
        finally:
            _lib.lib().cliora_debug_set(15, 0)
    assert len(res['narrow']) == len(res['wide'])
    # The two geometries differ in accumulation order (single vs. split tensor-memory accumulator, ~2e-6), which flips the
    # ReLU mask of a few pre-activations that sit within rounding of zero: gradients then move by ~5e-4 of max (the same
    # effect the oracle tests detect and re-draw for).  A staging overflow or a wrong column pass is an O(1) error.
    for a, b in zip(res['narrow'][:4], res['wide'][:4]):
        assert rel_err(b, a) < 1e-4
    for a, b in zip(res['narrow'][4:], res['wide'][4:]):
        assert rel_err(b, a) < 3e-3
