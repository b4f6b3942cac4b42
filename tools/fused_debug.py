import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_multigpu_gpu import _build
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
dev = torch.device('cuda', int(os.environ['LOCAL_RANK'])); torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
nx = 64
gen, c = _build(dev, nx)
single, sk = gen.eval_lattice(c, group=False); single = single.clone()
print(rank, 'multicast ptr', hex(getattr(gen, '_fused', None).grid_multicast) if getattr(gen, '_fused', None) else None)
for ex in ('nccl', 'fused', 'fused'):
    g, k = gen.eval_lattice(c, group=dist.group.WORLD, exchange=ex)
    torch.cuda.synchronize()
    diff = (g != single)
    idx = diff.nonzero()
    print(rank, ex, 'mismatch', int(diff.sum()), 'of', g.numel(), 'x range', (int(idx[:,0].min()), int(idx[:,0].max())) if len(idx) else None,
          'nan', int(torch.isnan(g).sum()), 'keys', k.tolist(), 'single keys', sk.tolist(), flush=True)
    dist.barrier()
print(rank, 'multicast ptr', hex(gen._fused.grid_multicast), flush=True)
dist.destroy_process_group()
