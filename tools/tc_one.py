import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtaco_b200.conv_onet.models import decoder_dict
torch.manual_seed(0)
dec = decoder_dict['simple_local'](dim=3, c_dim=32, hidden_size=32).cuda().eval()
dec.kernel_variant = int(os.environ.get('VARIANT', '7'))
nx = int(os.environ.get('NX', '256'))
c = {'grid': torch.randn(1, 32, 64, 64, 64, device='cuda')}
out = torch.empty(nx, nx, nx, device='cuda')
with torch.no_grad():
    for _ in range(3):
        dec.forward_dense(c, nx, out=out)
torch.cuda.synchronize()
