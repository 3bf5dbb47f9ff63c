"""Development probe (GPU): decode how the MN-major UMMA reads the smem tiles (single CTA, exact small integers)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200 import _lib as L
lib = L.lib()
M, Ka, Kb = 32, 128, 224
def run(A, B):
    Ap = torch.empty(2, M, Ka, device='cuda'); Bp = torch.empty(2, M, Kb, device='cuda')
    L.check(lib.cliora_split_tf32(L.ptr(A), A.numel(), L.ptr(Ap), L.stream()), 's')
    L.check(lib.cliora_split_tf32(L.ptr(B), B.numel(), L.ptr(Bp), L.stream()), 's')
    C = torch.zeros(Ka, Kb, device='cuda')
    scratch = torch.empty(int(lib.cliora_tc_matmul_tn_scratch_floats(M, Ka, Kb)) + 8, device='cuda')
    L.check(lib.cliora_tc_matmul_tn(M, Ka, Kb, L.ptr(Ap), L.ptr(Bp), L.ptr(C), 0, L.ptr(scratch), L.stream()), 'tn')
    torch.cuda.synchronize()
    return C.cpu()
torch.set_printoptions(linewidth=250, edgeitems=40, precision=0, sci_mode=False)
for r in (0, 1, 8, 9):
    A = torch.zeros(M, Ka, device='cuda'); B = torch.zeros(M, Kb, device='cuda')
    A[r] = torch.arange(1, Ka + 1, device='cuda').float()      # A[r, i] = i + 1
    B[r] = 1.0
    C = run(A, B)
    ref = (A.t() @ B).cpu()
    print('--- only reduction row r=%d nonzero; expect C[i,j] = i+1.  max|C-ref| = %.1f' % (r, (C - ref).abs().max()))
    print('C[:, 0]   :', C[:, 0][:40].tolist())
    print('C[0:4, :8]:', C[0:4, :8].tolist())
    A = torch.zeros(M, Ka, device='cuda'); B = torch.zeros(M, Kb, device='cuda')
    A[r] = 1.0
    B[r] = torch.arange(1, Kb + 1, device='cuda').float()
    C = run(A, B)
    print('B decode: C[0, :40]:', C[0, :40].tolist(), ' max err %.1f' % (C - (A.t() @ B).cpu()).abs().max())
