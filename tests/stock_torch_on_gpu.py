"""Like-for-like GPU baseline (last row): the reference's dense eager-PyTorch
formulation (oracle/cliora_oracle.py::CpuClioraStep, which restates cliora/net/trainer.py:272-304,450-455)
run on the same B200 with stock ATen/cuBLAS kernels.  Not a test and not part of the product: a one-off
measurement script that lives under tests/ because only tests may import the oracle.

    python tests/stock_torch_on_gpu.py [--steps 20] [--warmup 5] [--tf32]

Prints one JSON line; the committed result is profiles/r1_stock_torch_on_b200.json.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cliora_oracle import CpuClioraStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--length', type=int, default=20)
    ap.add_argument('--tf32', action='store_true', help='allow TF32 cuBLAS (default: fp32, like the reference)')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.backends.cuda.matmul.allow_tf32 = args.tf32
    torch.backends.cudnn.allow_tf32 = args.tf32
    B, n, D, R, F, V, E, K = args.batch, args.length, 400, 36, 2048, 8000, 1024, 100
    model = CpuClioraStep(D=D, E=E, V=V, F=F, k_neg=K, device=dev)
    g = torch.Generator().manual_seed(99)
    batches = []
    for _ in range(4):
        batches.append((torch.randint(0, V, (B, n), generator=g).to(dev),
                        torch.randperm(V, generator=g)[:K].to(dev),
                        torch.rand(B, R, F, generator=g).to(dev)))

    def step(i):
        s, neg, obj = batches[i % 4]
        model.opt.zero_grad()
        total, _ = model.loss(s, neg, obj)
        total.backward()
        torch.nn.utils.clip_grad_norm_(model.params, 5.0)
        model.opt.step()

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({'what': 'dense eager PyTorch restatement of the reference step on cuda:0 (stock ATen/cuBLAS)',
                      'value': B / ms * 1e3, 'unit': 'sentences/s', 'ms_per_step': ms, 'steps': args.steps,
                      'warmup': args.warmup, 'batch': B, 'length': n, 'matmul_tf32': bool(args.tf32),
                      'gpu': torch.cuda.get_device_name(0)}))


if __name__ == '__main__':
    main()
