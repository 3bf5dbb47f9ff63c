"""Fused clip-by-global-norm + Adam (the reference's Trainer.gradient_update, cliora/net/trainer.py:450-455:
``clip_grad_norm_(params, 5.0)`` then ``optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8).step()``) as three kernel
launches over all parameter tensors, instead of torch's ~40 small foreach / elementwise launches.

Graph-safe: gradients must live at fixed addresses (``zero_grad(set_to_none=False)`` or graph-captured
backward) and the step counter lives on the device.
"""
import ctypes

import torch

from . import _lib
from ._lib import check


class FusedClipAdam(object):
    def __init__(self, params, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=5.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params):
            raise _lib.ClioraError('FusedClipAdam needs contiguous fp32 CUDA parameters')
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.state = torch.zeros(3, device=dev, dtype=torch.float32)   # ||g||, clip coefficient, step
        self._grads = None
        n = len(self.params)
        self._table = torch.empty(int(_lib.lib().cliora_adam_table_bytes(n)), device=dev, dtype=torch.uint8)
        self._host_tables = []     # pinned staging copies, kept alive (a captured graph re-reads them on replay)
        self.param_groups = [dict(params=self.params, lr=lr)]

    def zero_grad(self, set_to_none=False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def _bind(self):
        """(Re)build the device table when gradient tensors appear or move, or the set of trainable tensors changes
        (a parameter frozen after construction -- Trainer.freeze_diora / freeze_except_vis -- is skipped, like
        torch's Adam skips tensors without a gradient: no moment update, no drift on old momentum)."""
        active = [i for i, p in enumerate(self.params) if p.requires_grad]
        for i in active:
            p = self.params[i]
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        key = (tuple(active), tuple(self.params[i].grad.data_ptr() for i in active))
        if self._grads == key:
            return
        n = len(active)
        self._n_active = n
        if n == 0:
            self._grads = key
            return
        L = _lib.lib()
        arr = lambda vals: (ctypes.c_void_p * n)(*vals)
        numel = (ctypes.c_int64 * n)(*[self.params[i].numel() for i in active])
        nbytes = int(L.cliora_adam_table_bytes(n))
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing or getattr(self, '_host_eager', None) is None:
            host = torch.empty(int(L.cliora_adam_table_bytes(len(self.params))), dtype=torch.uint8, pin_memory=True)
            if capturing:
                self._host_tables.append(host)   # a captured graph re-reads its staging copy on every replay
            else:
                self._host_eager = host          # eager rebinds reuse one pinned table (copied synchronously)
        else:
            host = self._host_eager
        total = ctypes.c_int64(0)
        check(L.cliora_adam_table_fill(n, arr([self.params[i].data_ptr() for i in active]), arr(list(key[1])),
                                       arr([self.exp_avg[i].data_ptr() for i in active]),
                                       arr([self.exp_avg_sq[i].data_ptr() for i in active]), numel, host.data_ptr(),
                                       ctypes.byref(total)), 'cliora_adam_table_fill')
        dev = self.params[0].device
        self._table[:nbytes].copy_(host[:nbytes], non_blocking=capturing)
        self._blocks = int(total.value)
        if getattr(self, '_scratch', None) is None or self._scratch.numel() < self._blocks:
            self._scratch = torch.empty(self._blocks, device=dev, dtype=torch.float32)
        self._grads = key

    @torch.no_grad()
    def step(self):
        self._bind()
        if self._n_active == 0:
            return
        self.lr = float(self.param_groups[0]['lr'])     # schedulers / manual changes write param_groups
        with torch.cuda.device(self.params[0].device):
            check(_lib.lib().cliora_adam_step(self._table.data_ptr(), self._n_active, self._blocks, self.lr,
                                              self.betas[0], self.betas[1], self.eps, self.max_norm,
                                              self.state.data_ptr(), self._scratch.data_ptr(), _lib.stream()),
                  'cliora_adam_step')

    def reset_state(self):
        """Back to step 0 with zero moments (in place: tensors keep their addresses, so captured graphs stay valid)."""
        for t in self.exp_avg + self.exp_avg_sq + [self.state]:
            t.zero_()

    @property
    def grad_norm(self):
        return self.state[0]
