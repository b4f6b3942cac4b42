"""Quick on-GPU probe: FP32 FMA peak (scalar vs packed) and dense-decode timing per variant."""
import ctypes as C
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vtaco_b200 import _abi
from vtaco_b200.conv_onet.models import decoder_dict

L = _abi.lib()
res = {}
for v in (0, 1):
    r = C.c_double(0)
    _abi.check(L.vtaco_fp32_peak(v, 4096, C.byref(r), _abi.stream_ptr()), 'peak')
    res['fp32_peak_v%d_tflops' % v] = r.value / 1e12
torch.manual_seed(0)
dec = decoder_dict['simple_local'](dim=3, c_dim=32, hidden_size=32).cuda().eval()
with torch.no_grad():
    for b in dec.blocks:
        b.fc_1.weight.normal_(0, 0.1)
nx = int(os.environ.get('NX', '256'))
c = {'grid': torch.randn(1, 32, 64, 64, 64, device='cuda')}
tri = {k: torch.randn(1, 32, 64, 64, device='cuda') for k in ('xz', 'xy', 'yz')}
out = torch.empty(nx, nx, nx, device='cuda')
for name, feats in (('grid', c), ('tri', tri)):
    for v in (0, 1, 2, 3):
        dec.kernel_variant = v
        with torch.no_grad():
            for _ in range(2):
                dec.forward_dense(feats, nx, out=out)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(5):
                dec.forward_dense(feats, nx, out=out)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res['dense%d_%s_v%d_ms' % (nx, name, v)] = ms
        res['dense%d_%s_v%d_gpts' % (nx, name, v)] = nx ** 3 / ms / 1e6
        res['dense%d_%s_v%d_tflops' % (nx, name, v)] = nx ** 3 * 30976 / ms / 1e9
        if v == 0:
            ref_out = out.clone()
        else:
            res['dense%d_%s_v%d_maxrelerr_vs_v0' % (nx, name, v)] = float(((out - ref_out).abs() / ref_out.abs().clamp(min=1)).max())
# flat random queries, training shape and large
for B, N in ((32, 2048), (1, 4000000)):
    p = (torch.rand(B, N, 3, device='cuda') - 0.5) * 1.1
    cc = {'grid': torch.randn(B, 32, 64, 64, 64, device='cuda')}
    ci = torch.randn(B, N, 32, device='cuda')
    dec.kernel_variant = 2
    with torch.no_grad():
        for _ in range(2):
            dec(p, cc)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            dec(p, cc)
        e1.record()
        torch.cuda.synchronize()
    res['flat_tc_%dx%d_ms' % (B, N)] = e0.elapsed_time(e1) / 5
    res['flat_tc_%dx%d_gpts' % (B, N)] = B * N / (e0.elapsed_time(e1) / 5) / 1e6
    dec.kernel_variant = 1
    with torch.no_grad():
        for _ in range(2):
            dec.forward_img(p, cc, ci)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            dec.forward_img(p, cc, ci)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res['flat_%dx%d_ms' % (B, N)] = ms
    res['flat_%dx%d_gpts' % (B, N)] = B * N / ms / 1e6
print(json.dumps(res, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/probe.json', 'w'), indent=1)
