"""Round-2 probe (NOT run yet): do the two sentence chains overlap better when each owns half of the SMs?

DESIGN.md section 8 item 2: two chains beat one by only 4 % because every kernel is shaped to fill the GPU and
the chains queue for the same SMs, while each kernel is individually latency-bound.  CUDA green contexts can
split the 148 SMs into two partitions; a stream created in a green context only runs on that partition.  This
probe builds two partitions with the driver API (cuda-python is in the image), wraps their streams as
``torch.cuda.ExternalStream`` and plants them in the chart's stream pool, then times eager and graph-captured
steps with and without the partitioning.

    timeout 300 python dev/proto/green_ctx_probe.py

Open questions this answers: (1) does stream capture accept green-context streams as forked capture streams
(if not, only the eager numbers print), (2) is the step faster with 74 + 74 SMs than with both chains on 148.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cuda.bindings import driver as cu  # noqa: E402


def ok(res):
    err, *vals = res
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError('driver call failed: %s' % (err,))
    return vals[0] if len(vals) == 1 else vals


def green_streams(n_groups=2, per_group=4):
    torch.cuda.init()
    torch.zeros(1, device='cuda')                                   # primary context exists
    dev = ok(cu.cuDeviceGet(0))
    sm = ok(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    total = sm.sm.smCount
    min_count = (total // n_groups) // 2 * 2                        # SM counts come in multiples of 2 on sm_100
    groups, n_out, remaining = ok(cu.cuDevSmResourceSplitByCount(n_groups, sm, 0, min_count))
    print('SMs: %d total -> %s (+%d left over)' % (total, [g.sm.smCount for g in groups[:n_out]],
                                                    remaining.sm.smCount if remaining else 0))
    streams = []
    for g in groups[:n_out]:
        desc = ok(cu.cuDevResourceGenerateDesc([g], 1))
        gctx = ok(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        row = []
        for _ in range(per_group):
            s = ok(cu.cuGreenCtxStreamCreate(gctx, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
            row.append(torch.cuda.ExternalStream(int(s)))
        streams.append(row)
    return streams


def build(cfg_batch=32):
    import bench
    cfg = dict(bench.CFG, B=cfg_batch, n=20)
    trainer = bench.build_trainer(cfg)
    batches = [bench.make_batch(cfg, 100 + i, device='cuda') for i in range(4)]
    return trainer, batches


def timed(fn, steps=30):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    from cliora_b200.net import chart
    trainer, batches = build()
    trainer.net.diora.chains = 2
    eager = lambda i: trainer.step(batches[i % 4], train=True, sync_result=False)
    print('eager, torch stream pool        : %.3f ms/step' % timed(eager))
    gs = green_streams(2, 2)
    # chart._streams(dev, k) hands out pool[:k]: chain i runs on pool[i], its auxiliary stream on pool[chains + i]
    chart._side_streams[torch.cuda.current_device()] = [gs[0][0], gs[1][0], gs[0][1], gs[1][1]]
    print('eager, one SM partition per chain: %.3f ms/step' % timed(eager))
    try:
        trainer.capture(batches[0], warmup=2)
        graphed = lambda i: trainer.step_graphed(batches[i % 4])
        print('graph, one SM partition per chain: %.3f ms/step' % timed(graphed, 60))
    except Exception as e:       # capture across green contexts may not be allowed
        print('graph capture with green-context streams failed: %r' % (e,))


if __name__ == '__main__':
    t0 = time.time()
    main()
    print('done in %.1f s' % (time.time() - t0))
