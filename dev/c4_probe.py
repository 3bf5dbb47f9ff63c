"""Dev probe (GPU): the c4 configuration alone (DIORA-MLP, length 64, hidden 400, batch 16, forward + backward, eager), for
`ncu -k regex:level_ ...` captures of the long-sentence regime."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cliora_b200.net.diora import DioraMLP
torch.manual_seed(0)
m = DioraMLP(400).cuda()
x = torch.randn(16, 64, 400, device='cuda', requires_grad=True)
for _ in range(int(os.environ.get('REPS', 2))):
    for p in m.parameters():
        p.grad = None
    m(x, x)
    (m.outside_h[:, :64].sum() + m.inside_s.sum() + m.outside_s.sum()).backward()
torch.cuda.synchronize()
print('c4 probe done')
