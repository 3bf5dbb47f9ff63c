"""Tiling rules of the fused level kernels, checked without a GPU through cliora_level_plan_query (host arithmetic only).
The two sizing bugs of round 2 -- CLIORA tiles capped by a fixed exchange buffer (two waves of clusters on the low inside
levels) and multi-wave narrow tiles on the low outside levels -- would have failed these."""
import pytest

from cliora_b200 import _lib

SM_SLOTS = 148            # CTAs of these kernels that can be resident (1 per SM)
CLUSTER_SLOTS = 132       # ... when launched as clusters of 4 (33 clusters; the no-device fallback of the query)
MAX_SMEM = 227 * 1024


def _levels(n):
    yield from (('fwd', False, False, l) for l in range(1, n))
    yield from (('fwd', True, False, l) for l in range(n - 2, -1, -1))
    yield from (('bwd', True, True, l) for l in range(0, n - 1))
    yield from (('bwd', False, True, l) for l in range(n - 1, 0, -1))


# (B per chain, n, D, R, chains): c2 as benchmarked (2 chains of 16), c2 in one chain, c5 (4 chains of 32), c3, c4, odd D
CONFIGS = [(16, 20, 400, 36, 2), (32, 20, 400, 36, 1), (32, 20, 400, 36, 4), (256, 30, 400, 0, 1), (8, 64, 400, 0, 2),
           (4, 12, 132, 4, 1), (8, 20, 512, 36, 1), (8, 20, 768, 0, 1), (2, 7, 64, 4, 1)]


@pytest.mark.parametrize('B,n,D,R,chains', CONFIGS)
def test_every_level_plan_is_consistent(B, n, D, R, chains):
    for _, outside, backward, level in _levels(n):
        p = _lib.level_plan(B, n, D, R, level, outside, backward, flags=chains << 8)
        N = n - level - 1 if outside else level
        assert p.splits == N and p.cells == B * (n - level)
        assert p.fused == 1, 'every level of these shapes is covered by the fused kernels'
        assert 1 <= p.cells_per_tile and p.cells_per_tile * N <= 128          # whole cells, one UMMA M of split rows
        assert (p.tiles - 1) * p.cells_per_tile < p.cells <= p.tiles * p.cells_per_tile
        assert p.column_slices in (1, 2, 4, 8) and p.slice_cols % 4 == 0
        assert p.column_slices * p.slice_cols >= D > (p.column_slices - 1) * p.slice_cols
        assert p.umma_n % 16 == 0 and p.slice_cols <= p.umma_n <= 208
        assert p.smem_bytes <= MAX_SMEM and p.ring_bytes < p.smem_bytes
        # the epilogue's staging must fit the operand rings it reuses
        stage = 128 * (p.slice_cols + 4) * 4
        if backward:
            # GZ + h + V of one column pass (level_bwd_kernel: a slice wider than 112 columns is scattered in passes of a
            # multiple of 16 columns, shrunk until three [128][pitch] stages fit)
            pitch = lambda c: c if (c // 4) % 2 else c + 4
            cols = p.slice_cols
            if cols > 112:
                cols = 112
                while cols > 16 and 3 * 128 * pitch(cols) * 4 > p.ring_bytes:
                    cols -= 16
            assert 3 * 128 * pitch(cols) * 4 <= p.ring_bytes, (level, p.slice_cols, cols)
        else:
            assert stage + p.cells_per_tile * p.slice_cols * 4 <= p.ring_bytes


@pytest.mark.parametrize('B,n,D,R,chains', CONFIGS[:5])
def test_wide_tiles_only_where_the_level_overflows_a_wave(B, n, D, R, chains):
    for _, outside, backward, level in _levels(n):
        p = _lib.level_plan(B, n, D, R, level, outside, backward, flags=chains << 8)
        N = n - level - 1 if outside else level
        min_tiles = -(-p.cells // (128 // N))
        slots = (SM_SLOTS if backward else CLUSTER_SLOTS) // chains
        if p.umma_n > 112:          # wide
            assert min_tiles * 4 > slots, (level, outside, backward)
        else:
            assert min_tiles * p.column_slices <= slots or p.column_slices < 4, (level, outside, backward)


def test_c2_low_inside_levels_fit_one_wave_of_clusters():
    """Round-2 regression: with R = 36 the five lowest inside levels were cut into 44 tiles of 14 cells (176 CTAs, two
    waves).  One chain of 32 sentences must fit the 33 co-resident clusters."""
    for level in range(1, 20):
        p = _lib.level_plan(32, 20, 400, 36, level, False, False, flags=1 << 8)
        assert p.tiles * p.column_slices <= CLUSTER_SLOTS, (level, p.tiles, p.cells_per_tile)


def test_unfusable_shapes_report_the_unfused_chain():
    assert _lib.level_plan(2, 6, 28, 0, 3).fused == 0                   # D < 32
    assert _lib.level_plan(2, 6, 1024, 0, 3).fused == 0                 # D too wide for eight column slices
    assert _lib.level_plan(1, 200, 64, 0, 150).fused == 0               # more than 128 splits per cell
    with pytest.raises(_lib.ClioraError):
        _lib.level_plan(2, 6, 64, 0, 6)                                 # no such inside level
