"""Round-2 harness for dev/proto/tc_gemm_f32a.cu (NOT run yet: written after round 1's GPU budget was spent).

    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC \\
         -o dev/proto/libproto_f32a.so dev/proto/tc_gemm_f32a.cu
    timeout 120 python dev/proto/test_f32a.py          # wrap in timeout: a barrier mistake would hang

Checks the prototype against float64 and against the shipped pair-operand kernel, then times both inside a CUDA
graph (dependent launches) at the one-wave shape of the compose GEMM.
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cliora_b200 import _lib as L  # noqa: E402

proto = ctypes.CDLL(os.path.join(ROOT, 'dev', 'proto', 'libproto_f32a.so'))
vp = ctypes.c_void_p
proto.proto_tc_linear_f32a.argtypes = [ctypes.c_int] * 3 + [vp, vp, vp, vp]
lib = L.lib()


def pair(x):
    out = torch.empty(2, *x.shape, device='cuda')
    L.check(lib.cliora_split_tf32(L.ptr(x), x.numel(), L.ptr(out), L.stream()), 'split')
    return out


def run(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    Wp, Ap = pair(W), pair(A)
    C1 = torch.zeros(M, N, device='cuda')
    C2 = torch.zeros(M, N, device='cuda')
    rc = proto.proto_tc_linear_f32a(M, N, K, L.ptr(A), L.ptr(Wp), L.ptr(C1), L.stream())
    assert rc == 0, rc
    L.check(lib.cliora_tc_linear(M, N, K, L.ptr(Ap), L.ptr(Wp), None, 0, L.ptr(C2), L.stream()), 'tc')
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t()
    e1 = float((C1.double() - ref).abs().max() / ref.abs().max())
    e2 = float((C2.double() - ref).abs().max() / ref.abs().max())
    return e1, e2


def bench(fn, reps=40):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)


if __name__ == '__main__':
    for M, N, K in [(128, 80, 32), (128, 80, 400), (300, 400, 400), (3200, 400, 400), (777, 396, 404)]:
        e1, e2 = run(M, N, K)
        print('M=%5d N=%3d K=%3d   f32a-proto err %.2e   pair kernel err %.2e' % (M, N, K, e1, e2))
        assert e1 < 5e-6, 'prototype is wrong'
    M, N, K = 3200, 400, 400
    A = torch.randn(M, K).cuda(); W = torch.randn(N, K).cuda(); Wp, Ap = pair(W), pair(A)
    C = torch.empty(M, N, device='cuda')
    t1 = bench(lambda: proto.proto_tc_linear_f32a(M, N, K, L.ptr(A), L.ptr(Wp), L.ptr(C), L.stream()))
    t2 = bench(lambda: lib.cliora_tc_linear(M, N, K, L.ptr(Ap), L.ptr(Wp), None, 0, L.ptr(C), L.stream()))
    print('one-wave compose shape: f32a-proto %.2f us   pair kernel %.2f us' % (t1, t2))
