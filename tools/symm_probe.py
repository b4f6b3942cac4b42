import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
dev = torch.device('cuda', int(os.environ['LOCAL_RANK'])); torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
try:
    t = symm.empty(1024 * 1024, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    print(rank, 'rendezvous ok', 'ptrs', [hex(p) for p in hdl.buffer_ptrs], flush=True)
    t.fill_(float(rank))
    hdl.barrier()
    peer = (rank + 1) % world
    pb = hdl.get_buffer(peer, (1024 * 1024,), torch.float32)
    print(rank, 'peer value', pb[:4].tolist(), flush=True)
    pb[10:20] = 100.0 + rank          # remote write
    hdl.barrier()
    torch.cuda.synchronize()
    print(rank, 'mine after remote write', t[8:12].tolist(), flush=True)
    # barrier latency
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(100): hdl.barrier()
    torch.cuda.synchronize(); print(rank, 'barrier us', (time.perf_counter() - t0) * 1e4, flush=True)
except Exception as e:
    print(rank, 'SYMM FAILED', repr(e), flush=True)
dist.destroy_process_group()
