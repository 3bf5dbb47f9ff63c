#!/usr/bin/env python
"""bench.py -- train sentences/sec of the CLIORA chart hot path on B200 (BASELINE.json metric).

Workload (config[1] of BASELINE.json): one CLIORA --obj_feats training step, batch 32 per GPU,
length 20, hidden 400, 36x2048 synthetic MAF object features, reconstruction + VG + contrastive loss,
backward, clip 5.0, Adam.  A "step" is one such pass over one synthetic batch.

  value     : sentences/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e       : the same step through the public API (Trainer.step) fed from PINNED HOST buffers, H2D of the
              batch and D2H of the loss inside the timed region
  roofline  : the dominant kernel class of the step, timed per launch with CUDA events on the launching
              stream in a separate instrumented pass (cliora_profile_*), against MEASURED_PEAKS.json
  cpu_baseline : the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py: its own build_net +
              Trainer.step) on this box's host cores -- kind "reference"; the oracle port (oracle/cliora_oracle.py,
              kind "port") only if the staged copy is missing
  gpu_baseline : the same unmodified reference with cuda=True on the same B200 (stock eager PyTorch)
  loss_check   : the product's loss on one batch next to the reference's loss on the same batch and weights
  sub_records  : the other BASELINE.json configs (c1, c3, c4, c5) timed the same way, each with its own roofline

`--impl reference` times the reference's CPU path alone (rank 0 only).
"""
import contextlib
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(B=32, n=20, D=400, R=36, F=2048, V=8000, E=1024, k_neg=100)
METRIC = 'train sentences/sec (fwd+bwd, L=20, D=400)'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tensor=d['bf16_tflops_sustained'], tensor_burst=d['bf16_tflops'], src='measured')
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src='fallback')


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksEventReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksEventReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksEventReasonSwPowerCap: 'sw_power_cap'}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


def make_batch(cfg, seed, device='cpu', pin=False):
    g = torch.Generator().manual_seed(seed)
    sent = torch.randint(0, cfg['V'], (cfg['B'], cfg['n']), generator=g)
    neg = torch.randperm(cfg['V'], generator=g)[:cfg['k_neg']]
    obj = torch.rand(cfg['B'], cfg['R'], cfg['F'], generator=g)
    if pin:
        sent, neg, obj = sent.pin_memory(), neg.pin_memory(), obj.pin_memory()
    return dict(sentences=sent.to(device), neg_samples=neg.to(device), obj_feats=obj.to(device),
                batch_size=cfg['B'], length=cfg['n'])


def build_trainer(cfg, seed=1234):
    from cliora_b200.net.trainer import build_net
    torch.manual_seed(seed)
    opts = argparse.Namespace(arch='mlp', hidden_dim=cfg['D'], k_neg=cfg['k_neg'], margin=1.0, vl_margin=0.2,
                              alpha_contr=1.0, alpha_vg=1.0, vg_loss=True, use_contr=True, use_contr_ce=False,
                              obj_feats=True, normalize='unit', share=True, cuda=True, lr=2e-3)
    emb = torch.nn.Embedding(cfg['V'], cfg['E'])
    trainer = build_net(opts, emb)
    enc = trainer.net.img_encoder
    with torch.no_grad():    # the reference zero-inits ImageEncoder; re-draw so the visual path is live
        for p in enc.parameters():
            p.normal_(0, 0.02)
    return trainer


def ref_kind():
    from oracle import ref_runner
    return 'reference' if ref_runner.available() else 'port'


def run_reference(args, cfg, cuda=False):
    """The reference's own implementation of the step (oracle/_ref: unmodified build_net + Trainer.step), all host
    threads (or cuda=True: stock eager PyTorch on the GPU); the oracle port only when the staged copy is absent."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import ref_runner
    batches = [make_batch(cfg, 99 + i) for i in range(2)]
    if ref_runner.available():
        with contextlib.redirect_stdout(sys.stderr):     # the reference prints from build_net
            v, ms, _ = ref_runner.time_reference(cfg, batches, args.steps, args.warmup, cuda=cuda)
        return v, ms, cores, 'reference'
    from oracle.cliora_oracle import CpuClioraStep
    model = CpuClioraStep(D=cfg['D'], E=cfg['E'], V=cfg['V'], F=cfg['F'], k_neg=cfg['k_neg'],
                          device='cuda' if cuda else 'cpu')
    dev = 'cuda' if cuda else 'cpu'
    batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()} for b in batches]
    for i in range(args.warmup):
        b = batches[i % 2]
        model.step(b['sentences'], b['neg_samples'], b['obj_feats'])
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        b = batches[i % 2]
        model.step(b['sentences'], b['neg_samples'], b['obj_feats'])
    if cuda:
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    return cfg['B'] / dt, dt * 1e3, cores, 'port'


def loss_check(trainer, cfg, dev):
    """Product loss vs the reference's loss on the same batch with the same weights (dropout off on both sides:
    the two draw their masks from different streams).  The checker is the unmodified reference on CPU
    (cliora/net/trainer.py:437-448 run_net) when staged, else the oracle port."""
    from oracle import ref_runner
    batch = make_batch(cfg, 4242)
    net = trainer.net
    p_saved = net.diora.atten_head.dropout.p
    net.diora.atten_head.dropout.p = 0.0
    net.train()
    with torch.no_grad():
        out = trainer.run_net({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}, None, True)
        mine = out['total_loss'].mean(dim=0).sum().item()
    net.diora.atten_head.dropout.p = p_saved
    if ref_runner.available():
        with contextlib.redirect_stdout(sys.stderr):
            ref = ref_runner.build_reference_trainer(cfg, cuda=False)
            ref.net.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()}, strict=True)
            ref.net.diora.atten_head.dropout.p = 0.0
            ref.net.train()
            with torch.no_grad():
                ro = ref.run_net(ref_runner.reference_batch(cfg, batch, 'cpu'), None, compute_loss=True)
        theirs, kind = ro['total_loss'].mean(dim=0).sum().item(), 'reference'
    else:
        from oracle.cliora_oracle import CpuClioraStep
        cpu = CpuClioraStep(D=cfg['D'], E=cfg['E'], V=cfg['V'], F=cfg['F'], k_neg=cfg['k_neg'])
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        with torch.no_grad():
            for k in cpu.P:
                cpu.P[k].copy_(sd['diora.' + k])
            cpu.emb.copy_(sd['embed.embeddings.weight']); cpu.mat.copy_(sd['embed.mat']); cpu.mat1.copy_(sd['embed.mat1'])
            cpu.recon_mat.copy_(sd['reconstruct_softmax_loss.mat'])
            for k in cpu.enc:
                cpu.enc[k].copy_(sd['img_encoder.' + k])
            theirs = cpu.loss(batch['sentences'], batch['neg_samples'], batch['obj_feats'], None)[0].item()
        kind = 'port'
    rel = abs(mine - theirs) / max(abs(theirs), 1e-12)
    return dict(product=mine, checker=theirs, rel_diff=rel, tol=1e-4, ok=bool(rel <= 1e-4), checker_kind=kind,
                note='same batch, same weights, dropout off on both sides, CPU checker')


def roofline_of(prof, nprof, pk, traffic_path=None):
    """Per-kernel-class table + the roofline object of the dominant class from a cliora_profile_* pass."""
    kernels = {}
    tot = sum(v['ms'] for v in prof.values()) or 1.0
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
        per = v['ms'] / v['launches']
        kernels[name] = dict(launches_per_step=v['launches'] // nprof, ms_per_step=v['ms'] / nprof,
                             share=v['ms'] / tot, tflops=v['flops'] / v['ms'] / 1e9 if v['ms'] else 0,
                             gbs=v['bytes'] / v['ms'] / 1e6 if v['ms'] else 0, avg_launch_us=per * 1e3)
    if not prof:
        return None, kernels
    # the dominant KERNEL: the fused level kernels are profiled as four classes (forward / backward x inside / outside
    # pass) of two kernels -- their classes are merged before the comparison
    merged = {}
    for k, v in prof.items():
        key = 'level_fwd' if k.startswith('level_fwd') else 'level_bwd' if k.startswith('level_bwd') else k
        m = merged.setdefault(key, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        for f in ('launches', 'ms', 'flops', 'bytes'):
            m[f] += v[f]
    name, v = max(merged.items(), key=lambda kv: kv[1]['ms'])
    traffic, traffic_note = None, None
    if traffic_path and os.path.exists(traffic_path):   # dram__bytes_read+write of one ncu --set full capture
        t = json.load(open(traffic_path)).get(name)
        if t:
            traffic, traffic_note = t['dram_bytes'], 'ncu capture of: ' + t['launch']
    # Which ceiling binds this kernel class: the larger of flops / tensor peak and algorithmic bytes / HBM peak.
    ach_tf = v['flops'] / v['ms'] / 1e9 if v['ms'] else 0.0
    ach_gb = v['bytes'] / v['ms'] / 1e6 if v['ms'] else 0.0
    frac_t, frac_h = ach_tf / pk['tensor'], ach_gb / pk['hbm']
    tcg = name.startswith('tc_') or 'level_' in name
    tensor_like = 'gemm' in name or 'atten_max' in name or 'level_' in name
    if tensor_like and frac_t >= frac_h:
        roof = dict(kernel=name, bound='tensor', achieved=ach_tf, peak=pk['tensor'], unit='TFLOP/s', frac=frac_t,
                    traffic=traffic, traffic_note=traffic_note, peak_source=pk['src'] + ' bf16 sustained')
    else:
        roof = dict(kernel=name, bound='hbm', achieved=ach_gb, peak=pk['hbm'], unit='GB/s', frac=frac_h,
                    traffic=traffic, traffic_note=traffic_note, peak_source=pk['src'])
    roof['frac_tensor'], roof['frac_hbm'] = frac_t, frac_h
    roof['achieved_tflops'], roof['achieved_gbs'] = ach_tf, ach_gb
    if 'level_' in name:
        roof['note'] = ('fused chart level (gather of the two projection rows of every split, compose GEMM on tcgen05, '
                        'split softmax, weighted sums, normalisation / region attention; the backward one: GY producer, '
                        'GZ GEMM, masked scatter).  Algorithmic bytes = rows x (5D+3) floats forward, (9D+3) backward; '
                        'algorithmic flops = 2 x rows x D x D.  At 40-70 flop/byte the kernel sits left of the ridge '
                        '(bf16 peak / HBM peak = %.0f flop/byte), so the HBM ceiling is the one that binds; frac_tensor is '
                        'against the bf16 peak although the GEMM runs fp32-accurate 3xTF32 (speed of light of that '
                        'scheme: 1/6 of the bf16 peak)' % (pk['tensor'] * 1e3 / pk['hbm']))
    elif tensor_like:
        roof['note'] = ('algorithmic flops = 2*M*N*K per launch; fp32-accurate 3xTF32 on tcgen05: the tensor pipe executes '
                        '3 tf32 UMMAs per algorithmic MAC, tf32 peak = half the bf16 peak, so frac 1/6 would be the speed '
                        'of light of this scheme') if tcg else \
            'algorithmic flops = 2*M*N*K per launch; warp-level mma.sync 3xTF32 / fp32 FMA kernel'
    return roof, kernels


def profile_pass(fn, nprof=3):
    """Eager launches with per-launch CUDA events recorded by the library on the launching stream.  A device-side
    sleep goes first so that the host runs ahead and the kernels execute back to back: otherwise every (start event,
    kernel, stop event) triple also measures the host's launch latency between the calls."""
    from cliora_b200 import _lib
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(0)                                   # how long the HOST needs to enqueue one eager step
    host_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    _lib.profile_start()
    for i in range(nprof):
        torch.cuda._sleep(int(min(0.25, 1.5 * host_s + 0.005) * 1.9e9))
        fn(i)
    return _lib.profile_stop()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='cliora_b200', choices=['cliora_b200', 'reference'])
    ap.add_argument('--batch', type=int, default=CFG['B'], help='sentences per GPU')
    ap.add_argument('--length', type=int, default=CFG['n'])
    ap.add_argument('--cpu-steps', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sub-records', action='store_true', help='skip the c1/c3/c4/c5 sub-records and gpu_baseline')
    ap.add_argument('--chains', type=int, default=None, help='concurrent sentence sub-batches (default: auto)')
    ap.add_argument('--fused', default='auto', choices=['auto', 'on', 'off'],
                    help='fused level kernels (auto: up to batch 32) or the unfused per-level kernel chain')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'tf32', 'bf16'],
                    help="fp32 = fp32-accurate 3xTF32 tensor-core GEMMs (headline); tf32 / bf16 = single-pass reduced precision with their own stated tolerance")
    ap.add_argument('--pdl', type=int, default=None, help='programmatic dependent launch on (1) / off (0)')
    ap.add_argument('--debug-set', default='', help='dev knobs: comma list of key=value passed to cliora_debug_set')
    ap.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()
    if args.gpus > 1 and 'WORLD_SIZE' not in os.environ and args.impl != 'reference':
        # called plainly with --gpus N (the driver launches torch.distributed.run itself): start one rank per GPU
        import socket
        with socket.socket() as sk:
            sk.bind(('127.0.0.1', 0))
            port = sk.getsockname()[1]
        os.execvp(sys.executable, [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node',
                                   str(args.gpus), '--master-addr', '127.0.0.1', '--master-port', str(port),
                                   os.path.abspath(__file__)] + sys.argv[1:])
    cfg = dict(CFG, B=args.batch, n=args.length)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    def workload_of(c):
        return ('CLIORA --obj_feats train step: batch %d/GPU, length %d, hidden %d, %dx%d object feats, '
                'recon+VG+contrastive, clip+Adam' % (c['B'], c['n'], c['D'], c['R'], c['F']))
    workload = workload_of(cfg)

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = max(1, min(args.steps, 6))
        warm = max(1, min(args.warmup, 1))
        a2 = argparse.Namespace(steps=steps, warmup=warm)
        v, ms, cores, kind = run_reference(a2, cfg)
        sample = ('%d warm-up + %d timed full CPU steps of the same workload (one process, batch %d, all %d host '
                  'threads; %s)' % (warm, steps, cfg['B'], cores,
                                    'unmodified reference build_net + Trainer.step from oracle/_ref' if kind == 'reference'
                                    else 'oracle port: oracle/_ref not staged'))
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'sentences/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload, 'device': 'cpu'},
            'cpu_baseline': {'value': v, 'unit': 'sentences/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': v, 'unit': 'sentences/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; cliora_b200 has no CPU path (use --impl reference for the CPU arm)')
    from cliora_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):      # normally shipped prebuilt (python -c 'import __graft_entry__ as g; g.build()')
        if world > 1:
            raise SystemExit('bench.py: build cliora_b200/libcliora_b200.so before a multi-rank run')
        _lib.build()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        dist.barrier()
    _lib.lib()
    if args.pdl is not None:
        _lib.lib().cliora_debug_set(100, args.pdl)
    for kv in filter(None, args.debug_set.split(',')):
        k, v = kv.split('=')
        _lib.lib().cliora_debug_set(int(k), int(v))
    dev = torch.device('cuda', local)
    use_graph = not args.no_graph
    pk = peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def train_record(c, steps, warmup, want_profile, sampler=None, check=False):
        """value / e2e / roofline of one training configuration (the c2 headline and the c5 sub-record)."""
        trainer = build_trainer(c)
        if args.chains is not None:
            trainer.net.diora.chains = args.chains
        trainer.net.diora.precision = args.precision
        trainer.net.diora.fused = {'auto': 'auto', 'on': True, 'off': False}[args.fused]
        sync = None
        if world > 1:
            from cliora_b200.parallel import GradSync
            sync = trainer.grad_sync = GradSync.for_module(trainer.net, world)
            trainer.ngpus = world
        rec = {}
        if check and rank == 0:
            rec['loss_check'] = loss_check(trainer, c, dev)
        # a few distinct resident batches (each rank its own shard, seed + rank), cycled
        resident = [make_batch(c, 1000 + 17 * rank + i, device=dev) for i in range(4)]
        host = [make_batch(c, 2000 + 17 * rank + i, pin=True) for i in range(4)]
        if use_graph:
            trainer.capture(resident[0])

        def step_eager(i):
            trainer.step(resident[i % len(resident)], train=True, sync_result=False)

        def step_resident(i):
            if use_graph:
                trainer.step_graphed(resident[i % len(resident)])   # device->device copy of the batch, then replay
            else:
                step_eager(i)

        h2d = sum(host[0][k].numel() * host[0][k].element_size() for k in ('sentences', 'neg_samples', 'obj_feats'))

        def step_e2e(i):
            hb = host[i % len(host)]
            if use_graph:
                return trainer.step_graphed(hb).item()   # pinned-host -> static device buffers, replay, D2H loss
            b = dict(hb)
            for k in ('sentences', 'neg_samples', 'obj_feats'):
                b[k] = hb[k].to(dev, non_blocking=True)
            out = trainer.step(b, train=True, sync_result=False)
            return out['total_loss'].item()          # D2H read of the step's loss

        for i in range(max(warmup, 3)):
            step_resident(i)
        if sampler is not None:
            sampler.start()
        l0 = _lib.launch_count()
        ms = timed(step_resident, steps)
        launches = _lib.launch_count() - l0
        if use_graph:
            launches = trainer.launches_per_step * steps   # replayed kernels are not re-counted by the library
        if sampler is not None:
            sampler.stop_flag = True
            sampler.join()
        ms_step = ms / steps
        rec.update(value=c['B'] * world * 1e3 / ms_step, ms_per_step=ms_step, gpu_launches=int(launches),
                   launches_per_step=int(launches // max(steps, 1)))
        for i in range(2):
            step_e2e(i)
        if use_graph:
            # pipelined public API: the pinned-host batch of step i+1 is prefetched (H2D on a side stream) while
            # step i runs; every step still pays its own H2D copy and a D2H read of its loss inside the timed region
            pending = [trainer.prefetch(host[0])]

            def step_e2e_pipelined(i):
                nxt = trainer.prefetch(host[(i + 1) % len(host)])
                loss = trainer.step_graphed(pending[0])
                pending[0] = nxt
                return loss.item()
            step_e2e_pipelined(0)
            ms_e2e = timed(step_e2e_pipelined, steps) / steps
        else:
            ms_e2e = timed(step_e2e, steps) / steps
        rec['e2e'] = {'value': c['B'] * world * 1e3 / ms_e2e, 'unit': 'sentences/s', 'ms_per_step': ms_e2e,
                      'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 4}
        if sync is not None:
            rec['ranks_in_sync'] = bool(sync.in_sync())     # parameter checksums agree across ranks after the run
        if want_profile and rank == 0:
            # single-rank pass: no collective may be issued; one stream so per-launch events do not overlap
            saved_sync, trainer.grad_sync = trainer.grad_sync, None
            saved_chains, trainer.net.diora.chains = trainer.net.diora.chains, 1
            prof = profile_pass(step_eager)
            trainer.grad_sync, trainer.net.diora.chains = saved_sync, saved_chains
            rec['roofline'], rec['kernels'] = roofline_of(prof, 3, pk, os.path.join(ROOT, 'profiles', 'r2_ncu_traffic.json'))
        del trainer
        torch.cuda.empty_cache()
        return rec

    sampler = ClockSampler(local)
    head = train_record(cfg, args.steps, args.warmup, True, sampler, check=True)

    # ---- sub-records: the other BASELINE.json configurations, same timing rules, fewer steps ----
    subs = {}
    sub_steps = max(10, min(args.steps, 40))
    if not args.no_sub_records and args.batch == CFG['B'] and args.length == CFG['n']:
        c5 = dict(CFG, B=128)
        r5 = train_record(c5, sub_steps, 3, world == 1)
        r5.update(config={'workload': workload_of(c5), 'global_batch': 128 * world, 'parallelism': 'dp%d' % world},
                  unit='sentences/s', steps=sub_steps)
        r5.pop('kernels', None)
        subs['c5'] = r5
        if world == 1 and args.precision == 'fp32':
            # the bf16-GEMM path (north_star: "a bf16-GEMM path with its own stated tolerance"): a separate record,
            # never the headline
            args.precision = 'bf16'
            rb = train_record(cfg, sub_steps, 3, False)
            args.precision = 'fp32'
            rb.update(config={'workload': workload, 'global_batch': cfg['B'], 'parallelism': 'dp1'}, unit='sentences/s',
                      steps=sub_steps, dtype='bf16',
                      tolerance='bf16 operands in the compose GEMMs of the level kernels (fp32 accumulate), single-pass TF32 '
                                'elsewhere: 3e-2 of max on chart vectors, 1e-2 on scores, 1e-1 on weight gradients '
                                '(tests/test_gpu_chart.py::test_bf16_gemm_mode_has_its_own_tolerance); CKY trees not guaranteed')
            rb.pop('kernels', None)
            subs['c2_bf16'] = rb
        if world == 1 and rank == 0:
            subs.update(other_configs(sub_steps, pk))

    if dist is not None:
        dist.barrier()
    cpu = gpu_base = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        a2 = argparse.Namespace(steps=args.cpu_steps, warmup=1)
        v, _, cores, kind = run_reference(a2, cfg)
        cpu = dict(value=v, unit='sentences/s', cores=cores, kind=kind,
                   sample='1 warm-up + %d timed full CPU steps of the same workload (%s, torch CPU, %d threads)'
                          % (args.cpu_steps, 'unmodified reference Trainer.step from oracle/_ref' if kind == 'reference'
                             else 'oracle port', cores))
        if not args.no_sub_records:
            a3 = argparse.Namespace(steps=10, warmup=3)
            gv, gms, _, gkind = run_reference(a3, cfg, cuda=True)
            gpu_base = dict(value=gv, unit='sentences/s', ms_per_step=gms, kind=gkind,
                            what='the same reference step with cuda=True on this B200 (stock eager PyTorch / cuBLAS '
                                 'fp32), 3 warm-up + 10 timed steps, CUDA events')
    if rank == 0:
        out = {
            'metric': METRIC, 'value': head['value'], 'unit': 'sentences/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'tf32': 'tf32', 'bf16': 'bf16'}[args.precision], 'data': 'synthetic',
            'dtype_note': ('fp32 storage and accumulation; GEMMs are 3xTF32 on tcgen05 (1e-6 of max vs fp64); chart '
                           'tensors and losses within 1e-4 of the reference, gradients within max(1e-4, 2x the fp32 '
                           'reference\'s own distance from fp64) -- see tests/test_gpu_chart.py') if args.precision == 'fp32'
                          else 'reduced-precision GEMM mode with its own stated tolerance (tests/test_gpu_chart.py)',
            'config': {'workload': workload, 'global_batch': cfg['B'] * world, 'parallelism': 'dp%d' % world,
                       'launch': 'cuda-graph replay' if use_graph else 'eager',
                       'l2': 'per-step working set (per-split buffers, > 0.4 GB) exceeds the 126 MB L2; 4 distinct batches cycled'},
            'clocks': sampler.summary(), 'gpu_launches': head['gpu_launches'],
            'launches_per_step': head['launches_per_step'], 'e2e': head['e2e'],
            'roofline': head.get('roofline'), 'kernels': head.get('kernels', {}), 'cpu_baseline': cpu,
            'gpu_baseline': gpu_base, 'loss_check': head.get('loss_check'), 'sub_records': subs,
        }
        if 'ranks_in_sync' in head:
            out['ranks_in_sync'] = head['ranks_in_sync']
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def other_configs(steps, pk):
    """c1 / c3 / c4 of BASELINE.json on one GPU (device-resident inputs, CUDA events), each with the roofline of
    its dominant kernel class.  These are reported next to the headline; their parity lives in tests/."""
    from cliora_b200 import _lib
    from cliora_b200.net.diora import DioraMLP
    from cliora_b200.analysis.cky import ParsePredictor
    dev = torch.device('cuda', torch.cuda.current_device())
    out = {}

    def run(name, B, n, fn, work, nprof=2):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        m.chains, saved = 1, m.chains
        prof = profile_pass(fn, nprof)
        m.chains = saved
        roof, _ = roofline_of(prof, nprof, pk)
        out[name] = dict(value=B * 1e3 / ms, unit='sentences/s', ms_per_step=ms, steps=steps, roofline=roof,
                         config={'workload': work, 'launch': 'eager'})

    torch.manual_seed(1234)
    m = DioraMLP(400).cuda()
    g = torch.Generator().manual_seed(5)
    # c1: DIORA-MLP chart fwd+bwd, batch 32, length 20 (the reference's CPU-runnable case)
    xs = [torch.randn(32, 20, 400, generator=g).to(dev) for _ in range(4)]

    def c1(i):
        x = xs[i % 4].requires_grad_()
        m(x, x)
        (m.outside_h[:, :20].sum() + m.inside_s.sum() + m.outside_s.sum()).backward()
    run('c1', 32, 20, c1, 'DIORA-MLP inside-outside fwd+bwd, hidden 400, batch 32, length 20, text only')
    # c4: long-sentence chart stress, length 64, batch 16
    xl = [torch.randn(16, 64, 400, generator=g).to(dev) for _ in range(2)]

    def c4(i):
        x = xl[i % 2].requires_grad_()
        m(x, x)
        (m.outside_h[:, :64].sum() + m.inside_s.sum() + m.outside_s.sum()).backward()
    run('c4', 16, 64, c4, 'DIORA-MLP inside-outside fwd+bwd, hidden 400, batch 16, length 64, text only')
    # c3: parse inference, batch 256, length 30: inside pass + CKY kernel + trees on the host
    m.eval()
    m.outside = False
    xp = [torch.randn(256, 30, 400, generator=g).to(dev) for _ in range(2)]
    pp = ParsePredictor(m)
    fake = {'sentences': torch.zeros(256, 30, dtype=torch.int64)}

    def c3(i):
        with torch.no_grad():
            m(xp[i % 2], xp[i % 2])
        pp.parse_batch(fake)
    run('c3', 256, 30, c3, 'CKY parse inference batch 256, length 30: inside pass + CKY kernel + nested-tuple trees')
    del m
    torch.cuda.empty_cache()
    return out


if __name__ == '__main__':
    main()
