"""Time the dense tcgen05 decoder (256^3, grid-64) and check it against the FFMA2 SIMT kernel."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtaco_b200.conv_onet.models import decoder_dict

torch.manual_seed(0)
dec = decoder_dict['simple_local'](dim=3, c_dim=32, hidden_size=32).cuda().eval()
with torch.no_grad():
    for n, p in dec.named_parameters():
        if n.endswith('fc_1.weight'):
            p.normal_(0, 0.1)
nx = int(os.environ.get('NX', '256'))
c = {'grid': torch.randn(1, 32, 64, 64, 64, device='cuda')}
out = torch.empty(nx, nx, nx, device='cuda')
ref = torch.empty(nx, nx, nx, device='cuda')
with torch.no_grad():
    dec.kernel_variant = 1
    dec.forward_dense(c, nx, out=ref)
    for variant in [int(v) for v in os.environ.get('VARIANTS', '2,4').split(',')]:
        dec.kernel_variant = variant
        for _ in range(3):
            dec.forward_dense(c, nx, out=out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            dec.forward_dense(c, nx, out=out)
        b.record()
        torch.cuda.synchronize()
        d = (out - ref).abs() / ref.abs().clamp(min=1)
        print('variant %d dense %d^3: %.3f ms  rel err vs fp32 SIMT: max %.2e mean %.2e'
              % (variant, nx, a.elapsed_time(b) / 10, d.max().item(), d.mean().item()))
