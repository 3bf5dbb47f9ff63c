"""Autograd bridges for the fused loss kernels (csrc/align_kernels.cuh)."""
import torch

from .. import _lib
from .._lib import check, ptr


class ContrastiveFn(torch.autograd.Function):
    """ContrastiveLoss.forward of the reference (cliora/net/trainer.py:91-128) on the max-over-regions
    scores smax [B,B,ncell]; the kernel produces the loss and its gradients in one pass."""

    @staticmethod
    def forward(ctx, smax, inside_s, outside_s, margin, alpha):
        smax = smax.contiguous().float()
        ins = inside_s.contiguous().float()
        outs = outside_s.contiguous().float()
        B, _, ncell = smax.shape
        cells = ins.shape[1]
        dev = smax.device
        loss = torch.empty((), device=dev, dtype=torch.float32)
        need = any(ctx.needs_input_grad[:3])
        g_s = torch.empty_like(smax) if need else None
        g_in = torch.zeros_like(ins) if need else None
        g_out = torch.zeros_like(outs) if need else None
        scratch = torch.empty(ncell * (B + 1) + 8, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_lib.lib().cliora_contrastive_loss(B, cells, ncell, ptr(smax), ptr(ins), ptr(outs), float(margin),
                                                     float(alpha), ptr(loss), ptr(g_s), ptr(g_in), ptr(g_out),
                                                     ptr(scratch), _lib.stream()), 'cliora_contrastive_loss')
        if need:
            ctx.save_for_backward(g_s, g_in, g_out)
        return loss

    @staticmethod
    def backward(ctx, g):
        g_s, g_in, g_out = ctx.saved_tensors
        return g * g_s, g * g_in, g * g_out, None, None


class VGLossFn(torch.autograd.Function):
    """VGLoss.forward of the reference (trainer.py:139-171) on wmax [B,B,n] = max over regions."""

    @staticmethod
    def forward(ctx, wmax, alpha):
        wmax = wmax.contiguous().float()
        B, _, n = wmax.shape
        dev = wmax.device
        loss = torch.empty((), device=dev, dtype=torch.float32)
        need = ctx.needs_input_grad[0]
        g_w = torch.empty_like(wmax) if need else None
        scratch = torch.empty(B + 8, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_lib.lib().cliora_vg_loss(B, n, ptr(wmax), float(alpha), ptr(loss), ptr(g_w), ptr(scratch),
                                            _lib.stream()), 'cliora_vg_loss')
        if need:
            ctx.save_for_backward(g_w)
        return loss

    @staticmethod
    def backward(ctx, g):
        (g_w,) = ctx.saved_tensors
        return g * g_w, None


class ReconCEFn(torch.autograd.Function):
    """mean over words of CE([pos.cell, neg_1.cell, ..., neg_K.cell], target 0)  (trainer.py:46-78), fused."""

    @staticmethod
    def forward(ctx, cell, pos, neg):
        cell = cell.reshape(-1, cell.shape[-1]).contiguous().float()
        pos = pos.reshape(-1, pos.shape[-1]).contiguous().float()
        neg = neg.reshape(-1, neg.shape[-1]).contiguous().float()
        rows, D = cell.shape
        K = neg.shape[0]
        rowloss = torch.empty(rows, device=cell.device, dtype=torch.float32)
        probs = torch.empty(rows, K + 1, device=cell.device, dtype=torch.float32)
        with torch.cuda.device(cell.device):
            check(_lib.lib().cliora_recon_ce_fwd(rows, D, K, ptr(cell), ptr(pos), ptr(neg), ptr(rowloss), ptr(probs),
                                                 _lib.stream()), 'cliora_recon_ce_fwd')
        ctx.save_for_backward(cell, pos, neg, probs)
        return rowloss.mean()

    @staticmethod
    def backward(ctx, g):
        cell, pos, neg, probs = ctx.saved_tensors
        L = _lib.lib()
        rows, D = cell.shape
        K = neg.shape[0]
        g = g.reshape(1).contiguous().float()
        gs = torch.empty(rows, K + 1, device=cell.device, dtype=torch.float32)
        g_cell, g_pos = torch.empty_like(cell), torch.empty_like(pos)
        g_all = torch.empty(K + 1, D, device=cell.device, dtype=torch.float32)
        scratch = torch.empty(int(L.cliora_matmul_tn_scratch_floats(rows, K + 1, D)) + 8, device=cell.device,
                              dtype=torch.float32)
        with torch.cuda.device(cell.device):
            st = _lib.stream()
            check(L.cliora_recon_ce_bwd(rows, D, K, ptr(cell), ptr(pos), ptr(neg), ptr(probs), ptr(g), ptr(gs),
                                        ptr(g_cell), ptr(g_pos), st), 'cliora_recon_ce_bwd')
            # gradient wrt the negatives: rows 1..K of g_scores^T cell (row 0 belongs to the per-word positives)
            check(L.cliora_matmul_tn(rows, K + 1, D, ptr(gs), ptr(cell), ptr(g_all), 0, ptr(scratch), st),
                  'cliora_matmul_tn')
        return g_cell, g_pos, g_all[1:]


def _pair(t):
    """Split pair [2, rows, cols] (tf32-rounded part, exact remainder) of a contiguous fp32 matrix."""
    out = torch.empty((2,) + tuple(t.shape), device=t.device, dtype=torch.float32)
    check(_lib.lib().cliora_split_tf32(ptr(t), t.numel(), ptr(out), _lib.stream()), 'cliora_split_tf32')
    return out


_TC_MIN_FLOP = 2e8     # below this the fp32 SIMT kernel is as fast (launch-bound)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (Embed / ImageEncoder / reconstruction projections, trainer.py:219-224, utils.py:52-55).
    Large problems run on the tcgen05 3xTF32 kernels (fp32-grade accuracy), small ones on the fp32 SIMT kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1]).contiguous().float()
        w = weight.contiguous().float()
        b = None if bias is None else bias.contiguous().float()
        M, K = x2.shape
        N = w.shape[0]
        L = _lib.lib()
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
        use_tc = (2.0 * M * N * K >= _TC_MIN_FLOP) and K % 4 == 0 and N % 4 == 0 and K >= 32
        xp = None
        with torch.cuda.device(x.device):
            if use_tc:
                xp, wp = _pair(x2), _pair(w)
                check(L.cliora_tc_linear(M, N, K, ptr(xp), ptr(wp), ptr(b), 0, ptr(out), _lib.stream()),
                      'cliora_tc_linear')
            else:
                check(L.cliora_linear(M, N, K, ptr(x2), ptr(w), ptr(b), 0, ptr(out), _lib.stream()), 'cliora_linear')
        ctx.use_tc = use_tc
        if use_tc:
            ctx.save_for_backward(xp, w)     # the input pair is reused by the weight-gradient GEMM
        else:
            ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        ctx.in_shape = x.shape
        return out.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, g):
        xs, w = ctx.saved_tensors
        L = _lib.lib()
        N, K = w.shape
        M = xs.shape[-2]
        g2 = g.reshape(M, N).contiguous().float()
        gx = gw = gb = None
        with torch.cuda.device(g.device):
            st = _lib.stream()
            if ctx.needs_input_grad[0]:
                gx = torch.empty(M, K, device=g.device, dtype=torch.float32)
                check(L.cliora_matmul_nn(M, K, N, ptr(g2), ptr(w), ptr(gx), 0, st), 'cliora_matmul_nn')
                gx = gx.view(ctx.in_shape)
            if ctx.needs_input_grad[1]:
                gw = torch.empty(N, K, device=g.device, dtype=torch.float32)
                if ctx.use_tc:
                    gp = _pair(g2)
                    scratch = torch.empty(int(L.cliora_tc_matmul_tn_scratch_floats(M, N, K)) + 8, device=g.device,
                                          dtype=torch.float32)
                    check(L.cliora_tc_matmul_tn(M, N, K, ptr(gp), ptr(xs), ptr(gw), 0, ptr(scratch), st),
                          'cliora_tc_matmul_tn')
                else:
                    scratch = torch.empty(int(L.cliora_matmul_tn_scratch_floats(M, N, K)) + 8, device=g.device,
                                          dtype=torch.float32)
                    check(L.cliora_matmul_tn(M, N, K, ptr(g2), ptr(xs), ptr(gw), 0, ptr(scratch), st),
                          'cliora_matmul_tn')
            if ctx.has_bias and ctx.needs_input_grad[2]:
                gb = g2.sum(0)
        return gx, gw, gb


def linear(x, weight, bias=None):
    return LinearFn.apply(x, weight, bias)
