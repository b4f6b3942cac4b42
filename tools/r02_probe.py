"""Round-2 probe (GPU box): forward vs forward_img(c_img tensor) on the tcgen05 kernel, weight re-pack
cost, cuBLAS TF32 / BF16 burst peaks (roofline denominators).  Writes gpurun_out/r02_probe.json."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vtaco_b200.conv_onet.models import decoder_dict


def timed(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


res = {'device': torch.cuda.get_device_name(0)}
torch.manual_seed(0)
dec = decoder_dict['simple_local'](dim=3, c_dim=32, hidden_size=32).cuda().eval()
with torch.no_grad():
    for b in dec.blocks:
        b.fc_1.weight.normal_(0, 0.1)
with torch.no_grad():
    for B, N in ((1, 100000), (1, 1000000), (32, 2048), (1, 10000000)):
        p = (torch.rand(B, N, 3, device='cuda') - 0.5) * 1.1
        c = {'grid': torch.randn(B, 32, 64, 64, 64, device='cuda')}
        ci = torch.randn(B, N, 32, device='cuda')
        for v in (7, 5, 1):
            dec.kernel_variant = v
            a = timed(lambda: dec(p, c))
            b = timed(lambda: dec.forward_img(p, c, ci))
            res['flat_%dx%d_v%d' % (B, N, v)] = {'forward_ms': a, 'forward_img_ms': b, 'forward_gpts': B * N / a / 1e6,
                                                 'forward_img_gpts': B * N / b / 1e6}
        del p, c, ci
    dec.kernel_variant = 7
    def repack():
        dec.invalidate()
        dec._packed_weights_tc(mixed=2)
    res['repack_ms'] = timed(repack)
# tensor peaks: cuBLAS burst, best of 10 (same method as MEASURED_PEAKS.json's bf16 figure)
n = 8192
for name, dt, tf32 in (('bf16', torch.bfloat16, False), ('tf32', torch.float32, True), ('fp32', torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device='cuda', dtype=dt)
    b = torch.randn(n, n, device='cuda', dtype=dt)
    best = 1e9
    for _ in range(3):
        a @ b
    for _ in range(10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res['cublas_%s_tflops' % name] = 2 * n ** 3 / best / 1e9
torch.backends.cuda.matmul.allow_tf32 = False
print(json.dumps(res, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/r02_probe.json', 'w'), indent=1)
