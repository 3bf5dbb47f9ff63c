"""The factored formulation + hand-derived backward (oracle/factored.py, which the CUDA
kernels implement phase by phase) against autograd over the dense oracle.  float64, CPU."""
import pytest
import torch

from oracle import cliora_oracle as O
from oracle import factored as F
from conftest import rel_err


def _run(B, n, D, R, share, train, seed):
    g = torch.Generator().manual_seed(seed)
    dt = torch.float64
    P = {k: v.clone().requires_grad_() for k, v in O.init_params(D, share=share, seed=seed, dtype=dt).items()
         if share is False or not k.startswith('outside_')}
    if share:
        for k in list(P):
            if k.startswith('inside_'):
                P['outside_' + k[len('inside_'):]] = P[k]
    x = torch.randn(B, n, D, generator=g, dtype=dt).requires_grad_()
    obj = (0.3 * torch.randn(B, R, D, generator=g, dtype=dt)).requires_grad_() if R else None
    C = O.num_cells(n)
    keep = (torch.rand(B, C, R, generator=g) >= 0.1) if (R and train) else None
    out = O.chart_forward(P, x, obj, keep)
    ct = {k: torch.randn(getattr(out, k).shape, generator=g, dtype=dt)
          for k in ('inside_h', 'inside_s', 'outside_h', 'outside_s')}
    sum((getattr(out, k) * ct[k]).sum() for k in ct).backward()

    with torch.no_grad():
        sv = F.forward({k: v.detach() for k, v in P.items()}, x.detach(),
                       None if obj is None else obj.detach(), keep, True, share)
        for k in ct:
            assert rel_err(getattr(sv, k).reshape(getattr(out, k).shape), getattr(out, k)) < 1e-12, k
        G, gx, gobj = F.backward(sv, ct['inside_h'], ct['inside_s'], ct['outside_h'], ct['outside_s'])
    assert rel_err(gx, x.grad) < 1e-10
    if R:
        assert rel_err(gobj, obj.grad) < 1e-10
    for k, v in G.items():
        ref = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        assert (ref.abs().max() == 0 and v.abs().max() == 0) or rel_err(v, ref) < 1e-10, k


@pytest.mark.parametrize('B,n,D,R,share,train', [
    (2, 5, 8, 0, True, False), (2, 6, 8, 0, False, False), (1, 1, 8, 0, True, False), (2, 2, 8, 0, True, False),
    (3, 5, 8, 4, True, False), (2, 6, 8, 5, True, True), (2, 4, 8, 3, False, True), (2, 1, 8, 3, True, True),
])
def test_factored_matches_dense_autograd(B, n, D, R, share, train):
    _run(B, n, D, R, share, train, seed=100 + n)


def test_dead_cell_clamped_norm():
    """All-ReLU-dead cell: sum p*h == 0 -> unit() clamps the norm at 1e-8 (SURVEY section 7 hard part 6)."""
    dt = torch.float64
    D, B, n = 6, 2, 3
    P = O.init_params(D, seed=5, dtype=dt)
    P = {k: v.clone() for k, v in P.items()}
    P['inside_compose_func.h_fcs.2.bias'] = torch.full((D,), -1e3, dtype=dt)   # kills every compose output
    for k in list(P):
        if k.startswith('inside_'):
            P['outside_' + k[len('inside_'):]] = P[k]
    Pg = {k: v.clone().requires_grad_() for k, v in P.items() if k.startswith('inside_') or k.startswith('root')}
    for k in list(Pg):
        if k.startswith('inside_'):
            Pg['outside_' + k[len('inside_'):]] = Pg[k]
    x = torch.randn(B, n, D, dtype=dt, generator=torch.Generator().manual_seed(1)).requires_grad_()
    out = O.chart_forward(Pg, x)
    assert out.inside_h[:, n:].abs().max() == 0
    g = torch.Generator().manual_seed(2)
    ct = {k: torch.randn(getattr(out, k).shape, generator=g, dtype=dt)
          for k in ('inside_h', 'inside_s', 'outside_h', 'outside_s')}
    sum((getattr(out, k) * ct[k]).sum() for k in ct).backward()
    with torch.no_grad():
        sv = F.forward(P, x.detach())
        G, gx, _ = F.backward(sv, ct['inside_h'], ct['inside_s'], ct['outside_h'], ct['outside_s'])
    assert torch.isfinite(gx).all()
    assert rel_err(gx, x.grad) < 1e-10
    for k, v in G.items():
        ref = Pg[k].grad if Pg[k].grad is not None else torch.zeros_like(Pg[k])
        assert (ref.abs().max() == 0 and v.abs().max() == 0) or rel_err(v, ref) < 1e-10, k
