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
  cpu_baseline : the oracle port of the reference's CPU path (oracle/cliora_oracle.py) on this box's cores

`--impl reference` times that CPU port alone (rank 0 only).
"""
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


def run_reference(args, cfg):
    """CPU arm: the oracle port of the reference's CPU path, all host threads, bounded sample."""
    from oracle.cliora_oracle import CpuClioraStep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = CpuClioraStep(D=cfg['D'], E=cfg['E'], V=cfg['V'], F=cfg['F'], k_neg=cfg['k_neg'])
    batch = make_batch(cfg, 99)
    for _ in range(args.warmup):
        model.step(batch['sentences'], batch['neg_samples'], batch['obj_feats'])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.step(batch['sentences'], batch['neg_samples'], batch['obj_feats'])
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    return cfg['B'] / dt, dt * 1e3, cores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='cliora_b200', choices=['cliora_b200', 'reference'])
    ap.add_argument('--batch', type=int, default=CFG['B'], help='sentences per GPU')
    ap.add_argument('--length', type=int, default=CFG['n'])
    ap.add_argument('--cpu-steps', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--chains', type=int, default=None, help='concurrent sentence sub-batches (default: auto)')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'tf32'],
                    help="fp32 = fp32-accurate 3xTF32 tensor-core GEMMs (headline); tf32 = single-pass TF32, tolerance 1e-2")
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
    workload = ('CLIORA --obj_feats train step: batch %d/GPU, length %d, hidden %d, %dx%d object feats, '
                'recon+VG+contrastive, clip+Adam' % (cfg['B'], cfg['n'], cfg['D'], cfg['R'], cfg['F']))

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = max(1, min(args.steps, 6))
        warm = max(1, min(args.warmup, 1))
        a2 = argparse.Namespace(steps=steps, warmup=warm)
        v, ms, cores = run_reference(a2, cfg)
        sample = '%d warm-up + %d timed full CPU steps of the same workload' % (warm, steps)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'sentences/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload, 'device': 'cpu'},
            'cpu_baseline': {'value': v, 'unit': 'sentences/s', 'cores': cores, 'kind': 'port', 'sample': sample},
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

    trainer = build_trainer(cfg)
    if args.chains is not None:
        trainer.net.diora.chains = args.chains
    trainer.net.diora.precision = args.precision
    if world > 1:
        from cliora_b200.parallel import GradSync
        trainer.grad_sync = GradSync([p for p in trainer.net.parameters() if p.requires_grad], world)
        trainer.ngpus = world
    dev = torch.device('cuda', local)
    # a few distinct resident batches (each rank its own shard, seed + rank), cycled
    resident = [make_batch(cfg, 1000 + 17 * rank + i, device=dev) for i in range(4)]
    host = [make_batch(cfg, 2000 + 17 * rank + i, pin=True) for i in range(4)]

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

    use_graph = not args.no_graph
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

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - l0
    if use_graph:
        launches = trainer.launches_per_step * args.steps   # replayed kernels are not re-counted by the library
    sampler.stop_flag = True
    sampler.join()
    ms_step = ms / args.steps
    value = cfg['B'] * world * 1e3 / ms_step

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
        ms_e2e = timed(step_e2e_pipelined, args.steps) / args.steps
    else:
        ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e = cfg['B'] * world * 1e3 / ms_e2e
    # ---- roofline pass: per-kernel-class CUDA-event timing of the same step (rank 0) ----
    roof, kernels = None, {}
    if rank == 0:
        pk = peaks()
        nprof = 3
        saved_sync, trainer.grad_sync = trainer.grad_sync, None   # single-rank pass: no collective may be issued
        saved_chains, trainer.net.diora.chains = trainer.net.diora.chains, 1   # one stream: per-launch events do not overlap
        _lib.profile_start()
        for i in range(nprof):
            # Eager launches: the per-kernel events are recorded by the library at launch time.  A device-side
            # sleep goes first so that the host runs ahead and the kernels execute back to back: otherwise every
            # (start event, kernel, stop event) triple also measures the host's launch latency between the calls.
            torch.cuda._sleep(int(0.02 * 1.9e9))
            step_eager(i)
        prof = _lib.profile_stop()
        trainer.grad_sync = saved_sync
        trainer.net.diora.chains = saved_chains
        tot = sum(v['ms'] for v in prof.values()) or 1.0
        for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
            per = v['ms'] / v['launches']
            kernels[name] = dict(launches_per_step=v['launches'] // nprof, ms_per_step=v['ms'] / nprof,
                                 share=v['ms'] / tot, tflops=v['flops'] / v['ms'] / 1e9 if v['ms'] else 0,
                                 gbs=v['bytes'] / v['ms'] / 1e6 if v['ms'] else 0, avg_launch_us=per * 1e3)
        top = max(prof.items(), key=lambda kv: kv[1]['ms'])
        name, v = top
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r1_ncu_traffic.json')
        if os.path.exists(tpath):      # dram__bytes_read+write of one ncu --set full capture of this kernel class
            t = json.load(open(tpath)).get(name)
            if t:
                traffic, traffic_note = t['dram_bytes'], 'ncu capture of: ' + t['launch']
        if 'gemm' in name or 'atten_max' in name:
            ach = v['flops'] / v['ms'] / 1e9
            tcg = name.startswith('tc_')
            roof = dict(kernel=name, bound='tensor', achieved=ach, peak=pk['tensor'], unit='TFLOP/s',
                        frac=ach / pk['tensor'], traffic=traffic, traffic_note=traffic_note,
                        peak_source=pk['src'] + ' bf16 sustained',
                        note=('algorithmic flops = 2*M*N*K per launch; fp32-accurate 3xTF32 on tcgen05: the tensor '
                              'pipe executes 3 tf32 UMMAs per algorithmic MAC, tf32 peak = half the bf16 peak, so '
                              'frac 1/6 would be the speed of light of this scheme') if tcg else
                        'algorithmic flops = 2*M*N*K per launch; warp-level mma.sync 3xTF32 / fp32 FMA kernel')
        else:
            ach = v['bytes'] / v['ms'] / 1e6
            roof = dict(kernel=name, bound='hbm', achieved=ach, peak=pk['hbm'], unit='GB/s', frac=ach / pk['hbm'],
                        traffic=traffic, traffic_note=traffic_note, peak_source=pk['src'])

    if dist is not None:
        dist.barrier()
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        a2 = argparse.Namespace(steps=args.cpu_steps, warmup=1)
        v, _, cores = run_reference(a2, cfg)
        cpu = dict(value=v, unit='sentences/s', cores=cores, kind='port',
                   sample='1 warm-up + %d timed full CPU steps of the same workload (oracle port, torch CPU)' % args.cpu_steps)
    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': 'sentences/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32' if args.precision == 'fp32' else 'tf32', 'data': 'synthetic',
            'config': {'workload': workload, 'global_batch': cfg['B'] * world, 'parallelism': 'dp%d' % world,
                       'launch': 'cuda-graph replay' if use_graph else 'eager',
                       'l2': 'per-step working set (~0.6 GB of per-split buffers) exceeds the 126 MB L2; 4 distinct batches cycled'},
            'clocks': sampler.summary(), 'gpu_launches': int(launches),
            'e2e': {'value': e2e, 'unit': 'sentences/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': 4},
            'roofline': roof, 'kernels': kernels, 'cpu_baseline': cpu,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
