"""Data parallelism over the sentence batch: one process per GPU, one grouped NCCL all-reduce of the gradient
tensors (average) per step -- the equivalent of the reference's DDP wrap (cliora/net/trainer.py:528-532,572-574).
In-batch negatives stay per rank, like the reference (cliora/data/batch_iterator.py:134-136): there is no feature
all-gather on the data path.

What DDP does and this mirrors:
  * at construction, rank 0's parameters (and buffers) are broadcast so every replica starts from the same
    weights -- the reference never seeds torch, so without this each rank would train its own model;
  * every step, gradients are averaged over ranks before clip + Adam (trainer.py:450-455 run on identical
    gradients on every rank, so no further collective is needed).

The all-reduce is one coalesced (ncclGroupStart/End) call over the gradient tensors IN PLACE: no flat staging
buffer, no copy passes, the 1/N folded into the collective (ReduceOp.AVG).  Gradient tensors keep their addresses,
so the call can sit between, or inside, CUDA graphs.
"""
import torch
import torch.distributed as dist


class GradSync(object):
    def __init__(self, params, world_size, group=None, buffers=(), broadcast=True):
        self.params = [p for p in params if p.requires_grad]
        self.world = world_size
        self.group = group
        self.device = self.params[0].device
        backend = dist.get_backend(group) if (world_size > 1 and dist.is_initialized()) else None
        self._avg = backend == 'nccl'        # gloo has no AVG: sum, then scale
        if broadcast and world_size > 1 and dist.is_initialized():
            self.broadcast_parameters(list(buffers))

    @classmethod
    def for_module(cls, module, world_size, group=None):
        """Wrap a whole module the way DDP does: trainable parameters are synchronised every step, and EVERYTHING
        else it owns (frozen parameters such as the word-embedding table, trainer.py:538-541, and buffers) is
        broadcast from rank 0 once, so replicas that were initialised from different random streams agree."""
        params = [p for p in module.parameters() if p.requires_grad]
        rest = [p for p in module.parameters() if not p.requires_grad] + list(module.buffers())
        return cls(params, world_size, group=group, buffers=rest)

    def broadcast_parameters(self, extra=()):
        """Rank 0's values into every replica (what wrapping a module in DDP does, trainer.py:573-574)."""
        with torch.no_grad():
            for t in list(self.params) + list(extra):
                dist.broadcast(t.data, src=0, group=self.group)

    def _grads(self):
        for p in self.params:
            if p.grad is None:       # unused in this step (the reference runs DDP with find_unused_parameters)
                p.grad = torch.zeros_like(p)
        return [p.grad for p in self.params]

    def __call__(self):
        if self.world <= 1:
            return
        grads = self._grads()
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        with dist._coalescing_manager(group=self.group, async_ops=False):   # fast path: one allreduce_coalesced
            for g in grads:
                dist.all_reduce(g, op=op, group=self.group)
        if not self._avg:
            torch._foreach_mul_(grads, 1.0 / self.world)

    def checksum(self):
        """[sum, sum of squares] of all parameters in float64 (order-fixed, so equal weights give equal sums)."""
        with torch.no_grad():
            s = torch.stack([p.detach().double().sum() for p in self.params]).sum()
            q = torch.stack([(p.detach().double() ** 2).sum() for p in self.params]).sum()
        return torch.stack([s, q])

    def in_sync(self):
        """True when every rank holds bit-identical parameter checksums (identical weights after identical
        averaged gradients and a deterministic optimizer step)."""
        if self.world <= 1 or not dist.is_initialized():
            return True
        mine = self.checksum()
        allc = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine, group=self.group)
        return all(bool(torch.equal(c, allc[0])) for c in allc)
