"""2-D U-Net of the feature planes: our fused path vs the torch.nn modules (cuDNN TF32 / fp32), eager and in a CUDA graph."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vtaco_b200.encoder.unet import UNet


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


res = {}
torch.manual_seed(0)
net = UNet(32, in_channels=32, depth=4, merge_mode='concat', start_filts=32).cuda().eval()
with torch.no_grad():
    for B, R in ((3, 32), (3, 64), (3, 128), (96, 32)):
        x = torch.randn(B, 32, R, R, device='cuda').contiguous(memory_format=torch.channels_last)
        r = {}
        net.fused = True
        r['fused_eager_ms'] = timed(lambda: net(x))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = net(x)
        r['fused_graph_ms'] = timed(g.replay)
        net.fused = False
        torch.backends.cudnn.allow_tf32 = True
        r['torch_tf32_eager_ms'] = timed(lambda: net(x))
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            y2 = net(x)
        r['torch_tf32_graph_ms'] = timed(g2.replay)
        torch.backends.cudnn.allow_tf32 = False
        r['torch_fp32_eager_ms'] = timed(lambda: net(x))
        torch.backends.cudnn.allow_tf32 = True
        res['B%d_R%d' % (B, R)] = r
print(json.dumps(res, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/r02_unet2d_probe.json', 'w'), indent=1)
