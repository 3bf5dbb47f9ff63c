"""Losses, Embed, ImageEncoder, Net and a minimal Trainer: drop-ins for the hot-path pieces of the
reference's cliora/net/trainer.py and cliora/net/utils.py:37-55, built on the fused kernels.

Class names, constructor arguments, ``forward`` signatures, returned ``(loss, dict)`` pairs and
``state_dict`` keys follow the reference so ``build_net`` / ``Trainer.step`` callers keep working.
"""
import os

import torch
import torch.nn as nn
import torch.optim as optim

from .losses import ContrastiveFn, ReconCEFn, VGLossFn, linear


class ImageEncoder(nn.Module):
    """cliora/net/utils.py:37-55: two Linear(2048 -> D) heads, zero-initialised like the reference."""

    def __init__(self, input_size, size):
        super().__init__()
        self.fc = nn.Linear(input_size, size)
        self.fc_vis = nn.Linear(input_size, size)
        self.reset_parameters()

    def reset_parameters(self):
        for p in self.parameters():
            if p.requires_grad:
                p.data.zero_()

    def forward(self, obj_feats):
        # one N = 2D GEMM for both heads (the input pair and its weight-gradient GEMM are shared)
        f = obj_feats.float()
        D = self.fc.weight.shape[0]
        both = linear(f, torch.cat([self.fc.weight, self.fc_vis.weight], 0), torch.cat([self.fc.bias, self.fc_vis.bias], 0))
        return both[..., :D].contiguous(), both[..., D:].contiguous()


class Embed(nn.Module):
    """cliora/net/trainer.py:204-224: embedding gather + two E->D projections."""

    def __init__(self, embeddings, input_size, size):
        super().__init__()
        self.input_size, self.size = input_size, size
        self.embeddings = embeddings
        self.mat = nn.Parameter(torch.FloatTensor(size, input_size))
        self.mat1 = nn.Parameter(torch.FloatTensor(size, input_size))
        self.reset_parameters()

    def reset_parameters(self):
        for p in self.parameters():
            if p.requires_grad:
                p.data.normal_()

    def forward(self, x):
        B, n = x.shape
        emb = self.embeddings(x.view(-1))
        D = self.mat.shape[0]
        both = linear(emb, torch.cat([self.mat, self.mat1], 0))      # one N = 2D GEMM for span and word projections
        return both[:, :D].contiguous().view(B, n, -1), both[:, D:].contiguous().view(B, n, -1)


class ReconstructionSoftmaxLoss(nn.Module):
    """cliora/net/trainer.py:25-78: CE over [positive, k_neg negatives] scored against outside_h leaves."""
    name = 'reconstruct_softmax_loss'

    def __init__(self, embeddings, input_size, size, margin=1, k_neg=3, cuda=False):
        super().__init__()
        self.k_neg, self.margin, self.input_size = k_neg, margin, input_size
        self.embeddings = embeddings
        self.mat = nn.Parameter(torch.FloatTensor(size, input_size))
        self._cuda = cuda
        self.reset_parameters()

    def reset_parameters(self):
        for p in self.parameters():
            if p.requires_grad:
                p.data.normal_()

    def forward(self, sentences, neg_samples, diora, info=None):
        B, n = sentences.shape
        cell = diora.outside_h[:, :n]                                   # [B,n,D]
        pos = linear(self.embeddings(sentences), self.mat)             # [B,n,D]
        neg = linear(self.embeddings(neg_samples), self.mat)           # [K,D]
        D, K = cell.shape[-1], neg.shape[0]
        if D % 4 == 0 and D <= 512 and K <= 127:
            loss = ReconCEFn.apply(cell.reshape(B * n, D), pos.reshape(B * n, D), neg.reshape(K, D))   # fused kernel
        else:   # shapes outside the fused kernel's range: same maths on the library GEMM + torch CE
            xp = (pos * cell).sum(-1, keepdim=True)
            xn = linear(cell, neg.reshape(K, D))
            score = torch.cat([xp, xn], 2).view(B * n, -1)
            loss = nn.functional.cross_entropy(score, torch.zeros(B * n, dtype=torch.int64, device=score.device))
        return loss, dict(reconstruction_softmax_loss=loss)


class ContrastiveLoss(nn.Module):
    """cliora/net/trainer.py:81-128.  Consumes max-over-regions scores of the first cells//2 cells
    straight from the alignment kernel; the [B,B,cells,R] tensor is never built."""
    name = 'contrastive_loss'

    def __init__(self, margin=1.0, alpha_contr=0.01, use_contr_ce=False):
        super().__init__()
        self.min_val = 1e-8
        self.margin, self.alpha_contr, self.use_contr_ce = margin, alpha_contr, use_contr_ce

    def forward(self, batch, diora):
        cells = diora.inside_s.shape[1]
        smax, _ = diora.span_region_max(cells // 2)
        loss = ContrastiveFn.apply(smax, diora.inside_s.squeeze(-1), diora.outside_s.squeeze(-1), self.margin,
                                   self.alpha_contr)
        return loss, dict(contrastive_loss=loss)


class VGLoss(nn.Module):
    """cliora/net/trainer.py:131-171.  ``vg`` is either the reference's 4-D vg_atten_score
    [B,B,n,R] or the already max-reduced [B,B,n] from ``diora.word_region_max()``."""
    name = 'vg_loss'

    def __init__(self, alpha_vg=0.1):
        super().__init__()
        self.min_val = 1e-8
        self.alpha_vg = alpha_vg

    def forward(self, batch, vg):
        wmax = vg.max(-1).values if vg.dim() == 4 else vg
        loss = VGLossFn.apply(wmax, self.alpha_vg)
        return loss, dict(vg_loss=loss)


def get_loss_funcs(options, embedding_layer=None):
    """cliora/net/trainer.py:174-201."""
    input_dim = embedding_layer.weight.shape[1]
    funcs = [ReconstructionSoftmaxLoss(embedding_layer, margin=options.margin, k_neg=options.k_neg,
                                       input_size=input_dim, size=options.hidden_dim, cuda=options.cuda)]
    if getattr(options, 'vg_loss', False):
        funcs.append(VGLoss(options.alpha_vg))
    if options.obj_feats and getattr(options, 'use_contr', False):
        funcs.append(ContrastiveLoss(options.vl_margin, options.alpha_contr, getattr(options, 'use_contr_ce', False)))
    return funcs


class Net(nn.Module):
    """cliora/net/trainer.py:227-304 (forward signature kept; visualisation is out of scope)."""

    def __init__(self, embed, image_encoder, diora, obj_feats, visualize=False, loss_funcs=()):
        super().__init__()
        self.obj_feats = obj_feats
        if self.obj_feats:
            self.img_encoder = image_encoder
        self.embed = embed
        self.diora = diora
        self.visualize = visualize
        self.loss_func_names = [m.name for m in loss_funcs]
        for m in loss_funcs:
            setattr(self, m.name, m)

    def compute_loss(self, batch, neg_samples, info=None, batch_parse=None):
        ret, loss = {}, []
        diora = self.diora
        for name in self.loss_func_names:
            func = getattr(self, name)
            if 'reconstruct' in name:
                sub, desc = func(batch, neg_samples, diora, info)
            elif 'contrastive' in name:
                sub, desc = func(batch, diora)
            elif 'vg_loss' in name:
                # training: max-reduced word scores straight from the kernel; eval: the reference's
                # 4-D tensor (it mixes in all_atten_score, cliora.py:463-464)
                vg = diora.word_region_max()[0] if diora.training else diora.vg_atten_score
                sub, desc = func(batch, vg)
            else:
                continue
            loss.append(sub.view(1, 1))
            ret.update(desc)
        return ret, torch.cat(loss, 1)

    def forward(self, img_ids, idx2word, batch, image_feats, obj_feats, boxes, obj_cates, neg_samples=None,
                compute_loss=True, info=None, batch_parse=None):
        embed_span, embed_word = self.embed(batch)
        obj_span = obj_word = None
        if self.obj_feats:
            obj_span, obj_word = self.img_encoder(obj_feats=obj_feats)
        self.diora(embed_span, embed_word, obj_span, obj_word)
        if compute_loss:
            ret, loss = self.compute_loss(batch, neg_samples, info=info, batch_parse=batch_parse)
        else:
            ret, loss = {}, torch.full((1, 1), 1, dtype=torch.float32, device=embed_span.device)
        ret['total_loss'] = loss
        return ret


class Trainer(object):
    """The training-step part of cliora/net/trainer.py:337-501 (run_net, gradient_update, step)."""

    def __init__(self, net, k_neg=None, ngpus=1, cuda=True):
        self.net = net
        self.optimizer = None
        self.cuda = cuda
        self.ngpus = ngpus
        self.grad_sync = None     # set by cliora_b200.parallel for data-parallel runs

    # ---- bookkeeping helpers of the reference Trainer (trainer.py:350-435), same names and file format ----
    def freeze_diora(self):
        for p in self.net.diora.parameters():
            p.requires_grad = False

    def freeze_except_vis(self):
        for name, p in self.net.named_parameters():
            if '_vis' not in name:
                p.requires_grad = False

    def parameter_norm(self, requires_grad=True, diora=False):
        net = self.net.diora if diora else self.net
        return sum(p.norm().item() for p in net.parameters() if p.requires_grad or not requires_grad)

    @staticmethod
    def get_single_net(net):
        return getattr(net, 'module', net) if isinstance(net, torch.nn.parallel.DistributedDataParallel) else net

    def save_model(self, save_emb, model_file):
        """``{'state_dict': ...}`` like the reference (trainer.py:382-397); ``save_emb=False`` leaves the (frozen,
        large) embedding tables out.  Files are interchangeable with the reference's in both directions."""
        state = {k: v for k, v in self.net.state_dict().items() if save_emb or 'embeddings' not in k}
        torch.save({'state_dict': state}, model_file)

    @staticmethod
    def load_model(origin_emb, net, model_file):
        """Load a checkpoint written by either implementation (trainer.py:399-435): a DDP ``module.`` prefix is
        stripped, unknown keys are dropped, embedding tables are kept from ``net`` unless ``origin_emb``, missing
        ``*_vis`` tensors outside the image encoder start from their non-visual twin, anything else missing keeps
        its current value."""
        target = Trainer.get_single_net(net)
        own = target.state_dict()
        loaded = torch.load(model_file, map_location='cpu')['state_dict']
        loaded = {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in loaded.items()}
        loaded = {k: v for k, v in loaded.items() if k in own}
        present = set(loaded)
        for k in own:
            if not origin_emb and 'embeddings' in k:
                loaded[k] = own[k]
            elif k not in present:
                twin = k.replace('_vis', '')
                if '_vis' in k and 'img_encoder' not in k and twin in loaded:
                    loaded[k] = loaded[twin]
                else:
                    loaded[k] = own[k]
        target.load_state_dict(loaded)

    def init_optimizer(self, optimizer_cls=optim.Adam, optimizer_kwargs=None, fused=None):
        """Adam(lr, betas, eps) like the reference (trainer.py:580).  On CUDA the default is the library's fused
        clip(5.0)+Adam (three launches for all tensors, cliora_b200/optim.py); ``fused=False`` keeps torch's."""
        kw = dict(optimizer_kwargs or dict(lr=2e-3, betas=(0.9, 0.999), eps=1e-8))
        params = [p for p in self.net.parameters() if p.requires_grad]
        if fused is None:
            fused = bool(self.cuda) and optimizer_cls is optim.Adam
        self.fused_optimizer = fused
        if fused:
            from ..optim import FusedClipAdam
            self.optimizer = FusedClipAdam(params, lr=kw.get('lr', 2e-3), betas=kw.get('betas', (0.9, 0.999)),
                                           eps=kw.get('eps', 1e-8), max_norm=5.0)
            return
        if self.cuda and optimizer_cls is optim.Adam:
            kw.setdefault('capturable', True)   # lets Trainer.capture() put the Adam step inside a CUDA graph
        self.optimizer = optimizer_cls(params, **kw)

    def run_net(self, batch_map, idx2word=None, compute_loss=True):
        batch = batch_map['sentences']
        return self.net(batch_map.get('example_ids'), idx2word, batch, batch_map.get('image_feats'),
                        batch_map.get('obj_feats'), batch_map.get('boxes'), batch_map.get('obj_cates'),
                        neg_samples=batch_map.get('neg_samples'), compute_loss=compute_loss, info={},
                        batch_parse=batch_map.get('GT'))

    def gradient_update(self, loss):
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.grad_sync is not None:
            if hasattr(self.grad_sync, 'unbind_flat') and not torch.cuda.is_current_stream_capturing():
                self.grad_sync.unbind_flat()      # eager step: fresh gradient tensors, reduced in place
            self.grad_sync()
        self._optimizer_update()

    # ---- CUDA-graph replay of the whole training step (forward, losses, backward, clip, Adam) ----
    def _optimizer_update(self):
        if getattr(self, 'fused_optimizer', False):
            self.optimizer.step()          # clip_grad_norm_(5.0) and Adam in one pass
            return
        params = [p for p in self.net.parameters() if p.requires_grad]
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        self.optimizer.step()

    def capture(self, batch_map, warmup=3, pool=None):
        """Capture one training step for the shapes of ``batch_map`` into CUDA graph(s).

        The step is ~500 dependent small kernels at batch 32 / length 20; replaying a graph removes the
        per-launch host cost and most inter-kernel gaps.  Batches of the same shape are then run with
        ``step_graphed`` (inputs are copied into static device buffers).  Needs an optimizer created with
        ``capturable=True`` (``init_optimizer`` does that when ``cuda`` is set).

        Single GPU: one graph holds the whole step.  Data parallel (``grad_sync`` set): the NCCL all-reduce
        stays outside -- graph A = forward + backward, eager all-reduce, graph B = clip + Adam."""
        from .. import _lib
        self.net.train()
        dev = next(self.net.parameters()).device
        self._static = {k: (v.to(dev).clone() if torch.is_tensor(v) else v) for k, v in batch_map.items()}
        # The chart module keeps the last step's (autograd-connected) outputs, which keeps the parameters'
        # AccumulateGrad nodes alive -- bound to whatever stream that step ran on.  If that was the legacy default
        # stream, a captured backward would have to make it wait on the capturing stream, which CUDA forbids.
        # Dropping those outputs lets the warm-up below re-create the nodes on its side stream.
        self._drop_autograd_state()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self._warmup_loss = None
        with torch.cuda.stream(side):
            for _ in range(warmup):
                out = self.run_net(self._static, None, compute_loss=True)
                self._warmup_loss = out['total_loss'].mean(dim=0).sum()
                self.gradient_update(self._warmup_loss)
                self._warmup_loss = self._warmup_loss.detach()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # data parallel: by default the all-reduce stays between two graphs; CLIORA_GRAPH_ALLREDUCE=1 captures the
        # NCCL call inside the one step graph instead (no eager gap between backward and clip+Adam)
        split = self.grad_sync is not None and os.environ.get('CLIORA_GRAPH_ALLREDUCE', '0') != '1'
        self._graph = torch.cuda.CUDAGraph()
        self._graph_opt = torch.cuda.CUDAGraph() if split else None
        self.optimizer.zero_grad(set_to_none=True)
        before = _lib.launch_count()
        gkw = {} if pool is None else {'pool': pool}
        with torch.cuda.graph(self._graph, **gkw):
            out = self.run_net(self._static, None, compute_loss=True)
            self._static_loss = out['total_loss'].mean(dim=0).sum()
            if split:
                self._static_loss.backward()
            else:
                self.gradient_update(self._static_loss)
            self._static_out = {k: v.detach() for k, v in out.items() if 'loss' in k}
        self.launches_per_step = _lib.launch_count() - before   # library kernels baked into the graph
        self._flat_state = None
        self._graph_grads = [p.grad for p in self.net.parameters() if p.requires_grad]
        if split:
            # no collective here: the captured backward has not run, its gradient tensors hold garbage, and a
            # rank that captures while its peers replay would pair this call with their real all-reduce.
            # The optimizer graph reads the parameters' gradients from ONE flat buffer (views): every replay packs
            # the backward graph's gradient tensors into it and all-reduces it in a single call (parallel.py).
            if hasattr(self.grad_sync, 'bind_flat') and os.environ.get('CLIORA_FLAT_ALLREDUCE', '1') != '0' and \
                    all(g is not None for g in self._graph_grads) and \
                    len(self._graph_grads) == len(getattr(self.grad_sync, 'params', ())):
                self.grad_sync.bind_flat(self._graph_grads)
                self._flat_state = (self.grad_sync._flat, self.grad_sync._flat_src, self.grad_sync._flat_views)
            with torch.cuda.graph(self._graph_opt, **gkw):
                self._optimizer_update()
        self._static_loss = self._static_loss.detach()     # only its value is read from here on
        return self

    def _drop_autograd_state(self):
        d = self.net.diora
        d.reset()
        d._run = None

    # ---- graph replay for a stream of batches whose shape varies (length-bucketed training) ----
    _GRAPH_STATE = ('_static', '_graph', '_graph_opt', '_static_loss', '_static_out', 'launches_per_step',
                    '_graph_grads', '_flat_state')

    @staticmethod
    def _shape_key(batch_map):
        return tuple((k, tuple(v.shape)) for k, v in sorted(batch_map.items()) if torch.is_tensor(v))

    def step_auto(self, batch_map, capture_after=2, max_graphs=64):
        """One training step on a batch of any shape, replaying a CUDA graph whenever one exists for that shape.

        The reference's sampler yields one sentence length per batch (cliora/data/dataloader.py:11-113), so a
        training run sees a few dozen distinct shapes over and over.  A shape is run eagerly until it has been seen
        ``capture_after`` times; that occurrence runs eagerly once more and is then captured, and later
        occurrences replay the graph.  All graphs share one memory pool (they are replayed one at a time and each
        is self-contained), so device memory is the maximum over shapes, not the sum.  Returns the total loss as
        a 0-d device tensor (a copy: valid until you drop it)."""
        if not hasattr(self, '_graphs'):
            self._graphs, self._seen, self._active_key = {}, {}, None
            self._pool = torch.cuda.graph_pool_handle()
            self._auto_stream = torch.cuda.Stream(device=next(self.net.parameters()).device)
        key = self._shape_key(batch_map)
        entry = self._graphs.get(key)
        if entry is None:
            seen = self._seen[key] = self._seen.get(key, 0) + 1
            if seen < capture_after or len(self._graphs) >= max_graphs:
                self._active_key = None
                # eager steps run on a side stream too: autograd state created on the legacy default stream
                # could not take part in a later capture (see capture())
                cur = torch.cuda.current_stream()
                self._auto_stream.wait_stream(cur)
                with torch.cuda.stream(self._auto_stream):
                    loss = self.step(batch_map, train=True, sync_result=False)['total_loss']
                cur.wait_stream(self._auto_stream)
                loss.record_stream(cur)
                return loss
            self.__dict__.pop('_stage', None)            # prefetch slots belong to one captured shape
            self.capture(batch_map, warmup=1, pool=self._pool)    # the warm-up step IS this batch's training step
            self._graphs[key] = {k: getattr(self, k) for k in self._GRAPH_STATE}
            self._active_key = key
            return self._warmup_loss.clone()
        if self._active_key != key:
            self.__dict__.pop('_stage', None)
            for k, v in entry.items():
                setattr(self, k, v)
            params = [p for p in self.net.parameters() if p.requires_grad]
            views = entry['_flat_state'][2] if entry.get('_flat_state') is not None else entry['_graph_grads']
            for prm, g in zip(params, views):     # what this shape's optimizer graph reads: the flat buffer's views
                prm.grad = g                      # (data parallel) or the tensors its backward graph writes
            self._active_key = key
        return self.step_graphed(batch_map).clone()

    def prefetch(self, batch_map):
        """Start the host->device copy of a (pinned) batch on a side stream into one of two staging slots and
        return a handle for ``step_graphed``; the copy overlaps the step that is currently running."""
        dev = next(self.net.parameters()).device
        if not hasattr(self, '_stage'):
            self._stage = [{k: torch.empty_like(v) for k, v in self._static.items() if torch.is_tensor(v)}
                           for _ in range(2)]
            self._stage_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slot = 0
            for e in self._stage_free:
                e.record(torch.cuda.current_stream(dev))
        slot = self._slot
        self._slot ^= 1
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free[slot])     # the step that last read this slot is done
            for k, dst in self._stage[slot].items():
                dst.copy_(batch_map[k], non_blocking=True)
            self._stage_ev[slot].record(self._copy_stream)
        return ('staged', slot)

    def step_graphed(self, batch_map):
        """Replay the captured step on a new batch of the captured shape; returns the (device) total loss.
        ``batch_map`` is a dict of tensors (host or device) or a handle from ``prefetch``."""
        if isinstance(batch_map, tuple) and batch_map[0] == 'staged':
            slot = batch_map[1]
            cur = torch.cuda.current_stream()
            cur.wait_event(self._stage_ev[slot])
            for k, src in self._stage[slot].items():
                self._static[k].copy_(src, non_blocking=True)        # device->device, a few microseconds
            self._stage_free[slot].record(cur)
        else:
            for k, v in batch_map.items():
                if torch.is_tensor(v):
                    self._static[k].copy_(v, non_blocking=True)
        self._graph.replay()
        if self._graph_opt is not None:
            if getattr(self, '_flat_state', None) is not None:      # this graph's own flat gradient buffer
                self.grad_sync._flat, self.grad_sync._flat_src, self.grad_sync._flat_views = self._flat_state
            self.grad_sync()          # pack + ONE fp32 all-reduce (AVG) over NVLink (NCCL)
            self._graph_opt.replay()
        return self._static_loss

    def prepare_info(self, batch_map):
        return {}

    def prepare_result(self, batch_map, model_output):
        """trainer.py:456-463: batch size, length and every ``*loss*`` entry as a Python float."""
        result = {'batch_size': batch_map['batch_size'], 'length': batch_map['length']}
        for k, v in model_output.items():
            if 'loss' in k:
                result[k] = v.mean(dim=0).sum().item()
        return result

    def step(self, batch_map, idx2word=None, train=True, compute_loss=True, sync_result=True):
        self.net.train() if train else self.net.eval()
        with torch.set_grad_enabled(train):
            out = self.run_net(batch_map, idx2word, compute_loss=compute_loss)
        total = out['total_loss'].mean(dim=0).sum()
        if train:
            self.gradient_update(total)
        if not sync_result:
            return {'total_loss': total.detach()}
        result = {'batch_size': batch_map.get('batch_size'), 'length': batch_map.get('length')}
        for k, v in out.items():
            if 'loss' in k:
                result[k] = v.mean(dim=0).sum().item()
        return result


def build_net(options, embeddings=None, random_seed=None):
    """cliora/net/trainer.py:504-582 minus process-group setup (see cliora_b200.parallel)."""
    size = options.hidden_dim
    if options.arch != 'mlp':
        raise NotImplementedError
    if options.obj_feats:
        from .cliora import DioraMLP as Diora
    else:
        from .diora import DioraMLP as Diora
    emb_kind = getattr(options, 'emb', 'none')
    origin_emb = emb_kind == 'none'
    if origin_emb:
        embedding_layer = embeddings
        if options.obj_feats:     # fine-tuning CLIORA from DIORA keeps the word table frozen (trainer.py:538-541)
            embedding_layer.weight.requires_grad = False
    else:                         # pretrained vectors: frozen table, row 0 is padding (trainer.py:542-546)
        table = torch.from_numpy(embeddings) if emb_kind == 'skip' and not torch.is_tensor(embeddings) else embeddings
        embedding_layer = nn.Embedding.from_pretrained(table, freeze=True, padding_idx=0)
    embed = Embed(embedding_layer, input_size=embedding_layer.weight.size(1), size=size)
    image_encoder = ImageEncoder(input_size=2048, size=size)
    diora = Diora(size, outside=True, normalize=options.normalize, compress=False, share=options.share)
    loss_funcs = get_loss_funcs(options, embedding_layer)
    net = Net(embed, image_encoder, diora, obj_feats=options.obj_feats, visualize=False, loss_funcs=loss_funcs)
    if getattr(options, 'load_model_path', None) is not None:
        Trainer.load_model(origin_emb, net, options.load_model_path)
    if options.cuda:
        net.cuda()
        diora.cuda()
    trainer = Trainer(net, k_neg=options.k_neg, ngpus=1, cuda=options.cuda)
    trainer.rank = getattr(options, 'local_rank', None)                 # trainer.py:578-579
    trainer.experiment_name = getattr(options, 'experiment_name', None)
    if getattr(options, 'multigpu', False):
        # the reference wraps the net in DDP here (trainer.py:528-532,572-574); this package keeps the process group
        # outside build_net: wrap with cliora_b200.parallel.GradSync.for_module(trainer.net, world) and set
        # trainer.grad_sync -- say so instead of silently training unsynchronised replicas
        import warnings
        warnings.warn('options.multigpu is set: install trainer.grad_sync = GradSync.for_module(trainer.net, world_size) '
                      '(cliora_b200.parallel) after init_process_group; build_net does not create process groups')
    trainer.init_optimizer(optim.Adam, dict(lr=options.lr, betas=(0.9, 0.999), eps=1e-8))
    return trainer
