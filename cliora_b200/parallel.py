"""Data parallelism over the sentence batch: one process per GPU, one flat fp32 gradient buffer, one
NCCL all-reduce (sum, then 1/N) per step -- the equivalent of the reference's DDP wrap
(cliora/net/trainer.py:528-532,572-574).  In-batch negatives stay per rank, like the reference
(cliora/data/batch_iterator.py:134-136): there is no feature all-gather on the data path."""
import torch
import torch.distributed as dist


class GradSync(object):
    def __init__(self, params, world_size, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.world = world_size
        self.group = group
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def __call__(self):
        if self.world <= 1:
            return
        for p in self.params:
            if p.grad is None:       # unused in this step (the reference runs DDP with find_unused_parameters)
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.params]
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world)
        torch._foreach_copy_(grads, self.views)    # in place: gradient tensors keep their addresses (graph-safe)
