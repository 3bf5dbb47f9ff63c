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
        """(Re)build the device table when gradient tensors appear or move."""
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        ptrs = [p.grad.data_ptr() for p in self.params]
        if self._grads == ptrs:
            return
        n = len(self.params)
        L = _lib.lib()
        arr = lambda vals: (ctypes.c_void_p * n)(*vals)
        numel = (ctypes.c_int64 * n)(*[p.numel() for p in self.params])
        nbytes = int(L.cliora_adam_table_bytes(n))
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)   # pinned: the copy below is graph-capturable
        total = ctypes.c_int64(0)
        check(L.cliora_adam_table_fill(n, arr([p.data_ptr() for p in self.params]), arr(ptrs),
                                       arr([t.data_ptr() for t in self.exp_avg]),
                                       arr([t.data_ptr() for t in self.exp_avg_sq]), numel, host.data_ptr(),
                                       ctypes.byref(total)), 'cliora_adam_table_fill')
        dev = self.params[0].device
        self._host_tables.append(host)
        self._table.copy_(host, non_blocking=True)
        self._blocks = int(total.value)
        if getattr(self, '_scratch', None) is None:
            self._scratch = torch.empty(self._blocks, device=dev, dtype=torch.float32)
        self._grads = ptrs

    @torch.no_grad()
    def step(self):
        self._bind()
        with torch.cuda.device(self.params[0].device):
            check(_lib.lib().cliora_adam_step(self._table.data_ptr(), len(self.params), self._blocks, self.lr,
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
