"""Development probe (GPU): pure device time of the SIMT vs tcgen05 GEMM at chart-level shapes (CUDA-graph replay)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200 import _lib as L
lib = L.lib()
N = K = 400
W = torch.randn(N, K).cuda(); b = torch.randn(N).cuda()
Wp = torch.empty(2, N, K, device='cuda')
L.check(lib.cliora_split_tf32(L.ptr(W), W.numel(), L.ptr(Wp), L.stream()), 's')
for cfg in (1, 3, 2):
  lib.cliora_debug_set(2, cfg)
  print('tile config', {1: 'narrow 80x4', 2: 'wide 256x2', 3: 'mid 160x3'}[cfg])
  for M in (608, 3200, 6720, 12160, 48640):
      A = torch.randn(M, K).cuda(); Ap = torch.empty(2, M, K, device='cuda')
      L.check(lib.cliora_split_tf32(L.ptr(A), A.numel(), L.ptr(Ap), L.stream()), 's')
      C = torch.empty(M, N, device='cuda')
      res = {}
      for name, fn in (('simt', lambda: lib.cliora_linear(M, N, K, L.ptr(A), L.ptr(W), L.ptr(b), 1, L.ptr(C), L.stream())),
                       ('tc', lambda: lib.cliora_tc_linear(M, N, K, L.ptr(Ap), L.ptr(Wp), L.ptr(b), 1, L.ptr(C), L.stream()))):
          s = torch.cuda.Stream()
          with torch.cuda.stream(s):
              for _ in range(3):
                  L.check(fn(), name)
              torch.cuda.synchronize()
              g = torch.cuda.CUDAGraph()
              with torch.cuda.graph(g, stream=s):
                  for _ in range(20):
                      L.check(fn(), name)
              g.replay(); torch.cuda.synchronize()
              e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              e0.record(s)
              for _ in range(5):
                  g.replay()
              e1.record(s); torch.cuda.synchronize()
              us = e0.elapsed_time(e1) * 1e3 / 100
          res[name] = us
      fl = 2.0 * M * N * K
      print('M=%6d  simt %7.1f us (%5.1f TF/s)   tc %7.1f us (%5.1f TF/s, x3 passes = %6.1f tf32 TF/s)' % (
          M, res['simt'], fl / res['simt'] / 1e6, res['tc'], fl / res['tc'] / 1e6, 3 * fl / res['tc'] / 1e6))
