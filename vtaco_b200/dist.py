"""Multi-GPU plumbing of the dense extraction: x-slab partition of the lattice and the
exchanges that follow a sharded decode —
  * `MeshExchange` (default): per-slab marching cubes, iso-level agreement and gather of the mesh pieces by
    device-side signalling over symmetric memory (csrc/exchange.cu);
  * `FusedExchange`, `RootExchange`: the round-1 logit exchanges (decoder epilogue stores / bulk peer copies);
  * `all_gather_slabs`, `all_reduce_minmax`: plain torch.distributed collectives (NCCL over NVLink on the GPU
    box, gloo in the CPU tests).
The reference is single-GPU (train.py:29) — there is nothing to mirror here."""
import torch
import torch.distributed as dist


def rank_world(group=None):
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    if group is False:
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def slab(nx, rank, world):
    """rows [x0, x1) of the lattice owned by `rank`: equal slabs of ceil(nx/world) rows
    (the last ranks may own fewer / none), so every slab is one contiguous, equally
    strided block of the flat (x-slowest) logit array."""
    per = (nx + world - 1) // world
    x0 = min(nx, rank * per)
    return x0, min(nx, x0 + per)


def slab_root(nx, rank, world, root_rows):
    """x-rows when rank 0 also runs marching cubes: the root decodes `root_rows` rows (even,
    >= 0), the remaining rows are split evenly (in bricks of 2 rows) over ranks 1..world-1."""
    if world == 1:
        return 0, nx
    root_rows = max(0, min(nx, int(root_rows) // 2 * 2))
    rest = nx - root_rows
    per = -(-rest // (world - 1))
    per += per & 1
    if rank == 0:
        return 0, root_rows
    x0 = min(nx, root_rows + (rank - 1) * per)
    return x0, min(nx, x0 + per)


def all_gather_slabs(grid, nx, group=None):
    """in-place all-gather of the x-slabs of `grid` (nx,nx,nx)."""
    rank, world = rank_world(group)
    if world == 1:
        return grid
    per = (nx + world - 1) // world
    if per * world == nx:
        x0, x1 = slab(nx, rank, world)
        dist.all_gather_into_tensor(grid.view(-1), grid[x0:x1].reshape(-1), group=group)
    else:  # ragged tail: gather into a padded buffer
        plane = nx * nx
        buf = grid.new_empty(per * world * plane)
        x0, x1 = slab(nx, rank, world)
        mine = grid.new_zeros(per * plane)
        mine[:(x1 - x0) * plane] = grid[x0:x1].reshape(-1)
        dist.all_gather_into_tensor(buf, mine, group=group)
        grid.view(-1).copy_(buf[:nx * plane])
    return grid


def all_reduce_minmax(keys, group=None):
    """keys int32[2] = ordered-int (min, max): min-reduce [0], max-reduce [1] in one call
    by negating the first entry."""
    rank, world = rank_world(group)
    if world == 1:
        return keys
    k = keys.to(torch.int64)
    k[0] = -k[0]
    dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
    k[0] = -k[0]
    keys.copy_(k.to(torch.int32))
    return keys


class FusedExchange(object):
    """Logit grid + iso-level key table in torch symmetric memory (peer-mapped over NVLink).

    The decoder kernel stores every logit of its slab straight into the grids of ALL ranks
    (`vtaco_decoder_args.logits_peers`), so the all-gather is fused into the decoder epilogue and
    overlaps the math; the (min,max) key pairs are published into a per-rank slot of every
    peer's table by a one-thread kernel.  Two symmetric-memory barriers per step order the
    writers against the marching-cubes readers.  No NCCL collective is on the data path."""

    def __init__(self, nx, device, group, use_multicast=True):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError('fused exchange supports up to 8 ranks (one NVSwitch domain)')
        self.grid = symm.empty((nx, nx, nx), dtype=torch.float32, device=device)
        self.table = symm.empty((2 * self.world,), dtype=torch.int32, device=device)
        self.h_grid = symm.rendezvous(self.grid, group)
        self.h_table = symm.rendezvous(self.table, group)
        self.grid_ptrs = [int(p) + int(getattr(self.h_grid, 'offset', 0)) for p in self.h_grid.buffer_ptrs]
        self.table_ptrs = [int(p) + int(getattr(self.h_table, 'offset', 0)) for p in self.h_table.buffer_ptrs]
        assert self.grid_ptrs[self.rank] == self.grid.data_ptr(), 'symmetric buffer pointer mismatch'
        mc = int(getattr(self.h_grid, 'multicast_ptr', 0) or 0)
        self.grid_multicast = (mc + int(getattr(self.h_grid, 'offset', 0))) if (mc and use_multicast) else 0
        self._tabs = (C.c_void_p * self.world)(*self.table_ptrs)
        self.device = device

    def barrier(self):
        self.h_grid.barrier()

    def publish(self, keys):
        from . import _abi
        with torch.cuda.device(self.device):
            st = _abi.lib().vtaco_publish_keys(_abi.ptr(keys), self._tabs, self.world, self.rank,
                                               _abi.stream_ptr(self.device))
        _abi.check(st, 'publish_keys')


class RootExchange(object):
    """Gather-to-root variant of the fused exchange for a stream of extractions.

    Only rank 0 extracts the mesh, so peers decode into a local grid and push their logit slab
    into rank 0's grid with ONE bulk peer-to-peer copy (copy engine over NVLink; measured on
    8 x B200: fine-grained 128-byte stores from the SMs sustain only ~140 GB/s of NVLink ingress
    per GPU, bulk copies ~700 GB/s) and publish their (min,max) keys into rank 0's table.  The
    grid and table are double-buffered: while rank 0 runs marching cubes on buffer b the peers
    already decode the next lattice into buffer b^1, and ONE symmetric-memory barrier per step
    (placed after the decode on every rank) orders both hazards:
      * peers pass barrier(s+1) only after rank 0 arrived there, i.e. after it finished marching
        cubes on the buffer they are about to overwrite at step s+2;
      * rank 0 passes barrier(s) only after every peer's decode(s) kernel (and key publish) ended.
    Rank 0 is given fewer lattice rows (`root_rows`) so that decode_0 + marching cubes takes as
    long as a peer's decode."""

    def __init__(self, nx, device, group):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError('fused exchange supports up to 8 ranks (one NVSwitch domain)')
        self.device = device
        self.grids, self.tables, self.root_grid_ptr, self._tabs, self._handles = [], [], [], [], []
        self.root_view = []
        for _ in range(2):
            g = symm.empty((nx, nx, nx), dtype=torch.float32, device=device)
            t = symm.empty((16,), dtype=torch.int32, device=device)
            hg, ht = symm.rendezvous(g, group), symm.rendezvous(t, group)
            self.grids.append(g)
            self.tables.append(t)
            self._handles.append((hg, ht))
            self.root_grid_ptr.append(int(hg.buffer_ptrs[0]) + int(getattr(hg, 'offset', 0)))
            self.root_view.append(hg.get_buffer(0, (nx, nx, nx), torch.float32))   # rank 0's grid, peer-mapped
            self._tabs.append((C.c_void_p * 1)(int(ht.buffer_ptrs[0]) + int(getattr(ht, 'offset', 0))))
        self.parity = 0

    def barrier(self):
        self._handles[0][0].barrier()

    def publish(self, keys, b):
        from . import _abi
        with torch.cuda.device(self.device):
            st = _abi.lib().vtaco_publish_keys(_abi.ptr(keys), self._tabs[b], 1, self.rank,
                                               _abi.stream_ptr(self.device))
        _abi.check(st, 'publish_keys')


class MeshExchange(object):
    """Sharded extraction: every rank runs marching cubes on its own x-slab (+ 2 halo rows) and only
    MESH PIECES cross NVLink (~12 B per vertex / face instead of 4 B per lattice point); reference
    call site of the pieces: generation.py:268-272.  Device-side signalling through a control
    block in torch symmetric memory (csrc/exchange.cu: vtaco_exchange_level / vtaco_exchange_mesh)
    — no host synchronisation and no collective on the data path, the whole step is one CUDA graph.

    gather='root': the pieces are concatenated in rank 0's buffers only; 'all': in every rank's."""

    def __init__(self, device, group, vertex_capacity, face_capacity, gather='root'):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _abi
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError('mesh exchange supports up to 8 ranks (one NVSwitch domain)')
        self.device = device
        self.gather = gather
        self.ctrl = symm.empty((_abi.EXCHANGE_CTRL_BYTES // 4,), dtype=torch.int32, device=device)
        self.ctrl.zero_()
        self.h_ctrl = symm.rendezvous(self.ctrl, group)
        self.ex = _abi.Exchange()
        for r, p in enumerate(self.h_ctrl.buffer_ptrs):
            self.ex.ctrl[r] = int(p) + int(getattr(self.h_ctrl, 'offset', 0))
        assert self.ex.ctrl[self.rank] == self.ctrl.data_ptr(), 'symmetric buffer pointer mismatch'
        self.ex.world, self.ex.rank = self.world, self.rank
        self.level_ptr = self.ctrl.data_ptr() + _abi.EXCHANGE_LEVEL_OFFSET
        self.totals = torch.zeros(2, dtype=torch.int64, device=device)
        self.verts = self.faces = None
        self._alloc(vertex_capacity, face_capacity)
        torch.cuda.synchronize(device)
        self.h_ctrl.barrier()          # every control block is zeroed before anyone signals
        torch.cuda.synchronize(device)

    def _alloc(self, vcap, fcap):
        """(re)allocate the destination buffers — collective: every rank calls it with the same sizes."""
        import torch.distributed._symmetric_memory as symm
        self.verts = symm.empty((int(vcap), 3), dtype=torch.float32, device=self.device)
        self.faces = symm.empty((int(fcap), 3), dtype=torch.int32, device=self.device)
        hv, hf = symm.rendezvous(self.verts, self.group), symm.rendezvous(self.faces, self.group)
        self._handles = (hv, hf)
        vp = [int(p) + int(getattr(hv, 'offset', 0)) for p in hv.buffer_ptrs]
        fp = [int(p) + int(getattr(hf, 'offset', 0)) for p in hf.buffer_ptrs]
        self._dst = [(vp[r], fp[r]) if (self.gather == 'all' or r == 0) else (0, 0) for r in range(self.world)]

    def ensure_capacity(self, total_v, total_f):
        """grow the destination buffers to hold the given mesh totals (same on every rank: they
        come out of the all-gathered counts) — collective."""
        if total_v > self.verts.shape[0] or total_f > self.faces.shape[0]:
            torch.cuda.synchronize(self.device)
            self.h_ctrl.barrier()
            torch.cuda.synchronize(self.device)
            self._alloc(max(int(total_v * 1.5) + 16, self.verts.shape[0]), max(int(total_f * 1.5) + 16, self.faces.shape[0]))
            return True
        return False

    def shared_grid(self, nx):
        """(nx,nx,nx) lattice grid in symmetric memory and every rank's address of it (collective)."""
        import torch.distributed._symmetric_memory as symm
        grid = symm.empty((nx, nx, nx), dtype=torch.float32, device=self.device)
        h = symm.rendezvous(grid, self.group)
        ptrs = [int(p) + int(getattr(h, 'offset', 0)) for p in h.buffer_ptrs]
        assert ptrs[self.rank] == grid.data_ptr(), 'symmetric buffer pointer mismatch'
        self._grid_handle = h
        return grid, ptrs

    def level(self, keys):
        from . import _abi
        import ctypes as C
        with torch.cuda.device(self.device):
            st = _abi.lib().vtaco_exchange_level(C.byref(self.ex), _abi.ptr(keys), _abi.stream_ptr(self.device))
        _abi.check(st, 'exchange_level')

    def push(self, counts, verts, faces):
        """all-gather the (V,F) counts and concatenate this rank's piece into the destination(s)."""
        from . import _abi
        import ctypes as C
        m = _abi.MeshPiece()
        m.counts, m.vertices, m.faces = counts.data_ptr(), verts.data_ptr(), faces.data_ptr()
        for r, (v, f) in enumerate(self._dst):
            m.dst_vertices[r], m.dst_faces[r] = v or None, f or None
        m.vertex_capacity, m.face_capacity = self.verts.shape[0], self.faces.shape[0]
        m.total_counts = self.totals.data_ptr()
        with torch.cuda.device(self.device):
            st = _abi.lib().vtaco_exchange_mesh(C.byref(self.ex), C.byref(m), _abi.stream_ptr(self.device))
        _abi.check(st, 'exchange_mesh')

    def timed_out(self):
        """True if a device-side wait gave up (a peer never arrived); synchronises."""
        from . import _abi
        return bool(self.ctrl[_abi.EXCHANGE_ERR_OFFSET // 4].item())

    def has_result(self):
        return self.gather == 'all' or self.rank == 0
