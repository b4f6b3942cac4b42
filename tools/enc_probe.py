import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_models, synthetic_scene
dev = torch.device('cuda')
net = build_models(dev)
cloud = torch.from_numpy(synthetic_scene(0)[0])[None].to(dev)
enc = net.encoder
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad():
    print('pointnet part ms', timeit(lambda: enc.pointnet_features(cloud)))
    f = enc.pointnet_features(cloud)['grid']
    print('unet3d on channels_last_3d view ms', timeit(lambda: enc.unet3d(f)))
    fc = f.contiguous()
    print('unet3d on contiguous ms', timeit(lambda: enc.unet3d(fc)))
    print('relayout to contiguous ms', timeit(lambda: f.contiguous()))
    torch.backends.cudnn.benchmark = True
    print('unet3d contiguous + cudnn.benchmark ms', timeit(lambda: enc.unet3d(fc)))
    print('unet3d channels_last + cudnn.benchmark ms', timeit(lambda: enc.unet3d(f)))
    o = enc.unet3d(fc); print(o.shape, o.stride())
