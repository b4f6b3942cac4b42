"""cProfile of the host side of one decoder training step (forward_img + backward, B=32 x 2048, grid-64)."""
import cProfile, pstats, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtaco_b200.conv_onet.models import decoder_dict
torch.manual_seed(0)
B, N, R = 32, 2048, 64
dec = decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, hidden_size=32).cuda().train()
feat = torch.randn(B, 32, R, R, R, device='cuda').contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
p = torch.rand(B, N, 3, device='cuda') - 0.5
c_img = torch.randn(B, N, 32, device='cuda', requires_grad=True)
r = torch.randn(B, N, device='cuda')
def step():
    dec.zero_grad(set_to_none=True); feat.grad = None; c_img.grad = None
    (dec.forward_img(p, {'grid': feat}, c_img) * r).sum().backward()
for _ in range(5): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host per step %.1f us, incl. drain %.1f us' % ((t1 - t0) / 50 * 1e6, (t2 - t0) / 50 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(30); print(s.getvalue()[:6000])
