"""cProfile of the host side of a flat LocalDecoder.forward call (training shape B=32 x 2048, no grad)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtaco_b200.conv_onet.models import decoder_dict
torch.manual_seed(0)
dec = decoder_dict['simple_local'](dim=3, c_dim=32, hidden_size=32).cuda().eval()
p = (torch.rand(32, 2048, 3, device='cuda') - 0.5) * 1.1
c = {'grid': torch.randn(32, 32, 64, 64, 64, device='cuda').permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)}
ci = torch.randn(32, 2048, 32, device='cuda')
with torch.no_grad():
    for _ in range(20):
        dec(p, c)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(200):
        dec(p, c)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('host time per call %.1f us; incl. drain %.1f us' % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        dec(p, c)
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22)
    print(s.getvalue()[:4500])
