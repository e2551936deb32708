"""Host-side sharding for multi-GPU runs (one process per GPU, torch.distributed for the plumbing).

The reference is single-GPU (one GL context, main.cpp:44); SURVEY.md 8e lists where the path shards:
  * cone tracing: image rows / whole views are independent given the grid (grid replicated, no collective);
  * voxelisation of large dynamic meshes: contiguous triangle ranges per rank into a private integer
    accumulator, then ONE exchange step: all-reduce(sum) of the uint32 accumulator.  Integer sums are
    order independent, so the sharded result is bit-identical to the single-GPU one.
"""
from __future__ import annotations

import numpy as np


def triangle_range(n_tris: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) triangle range of `rank`."""
    base, rem = divmod(int(n_tris), int(world))
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


def triangle_share(ctx, n_tris: int, rank: int, world: int, interleave: bool = True):
    """Triangle share of `rank` for sharded voxelisation; returns the [begin, end) to pass to voxelize_range /
    voxelize_shared.  interleave=True deals blocks of 128 triangles round-robin (TriangleInterleave / TrianglePhase):
    contiguous ranges of a real mesh differ several-fold in fragments per triangle, interleaved shares do not."""
    if interleave:
        ctx.set_i("TriangleInterleave", world); ctx.set_i("TrianglePhase", rank)
        return 0, int(n_tris)
    ctx.set_i("TriangleInterleave", 1); ctx.set_i("TrianglePhase", 0)
    return triangle_range(n_tris, rank, world)


def row_band(height: int, rank: int, world: int, align: int = 8):
    """[begin, end) rows of the frame for `rank`; bands are multiples of `align` rows (cone_trace's warp
    tile height) except possibly the last."""
    blocks = (height + align - 1) // align
    b0, b1 = triangle_range(blocks, rank, world)
    return min(b0 * align, height), min(b1 * align, height)


def row_band_equal(height: int, rank: int, world: int, align: int = 8):
    """Row bands of EQUAL nominal size (a multiple of `align`), the last ones clipped to the frame: what a single
    all-gather of the bands needs.  Returns (begin, end, rows_per_band)."""
    blocks = (height + align - 1) // align
    per = ((blocks + world - 1) // world) * align
    return min(rank * per, height), min((rank + 1) * per, height), per


def row_strips_for_rank(height: int, rank: int, world: int, strip: int = 8):
    """Interleaved row sharding (RowInterleave / RowPhase; what vct_comm_init deals by default): the frame is cut into
    strips of `strip` rows -- cone_trace's block height -- and rank r renders strips r, r + world, ...  Returns the
    [begin, end) row ranges of `rank`.  Mirrors row_owned() / launch_cone() in csrc/vct_cone.cu."""
    out = []
    g = rank
    while g * strip < height:
        out.append((g * strip, min((g + 1) * strip, height)))
        g += world
    return out


def views_for_rank(n_views: int, rank: int, world: int):
    """Round-robin view assignment (light-probe bake, BASELINE config 5)."""
    return list(range(rank, n_views, world))


class _DevicePointer:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def accumulator_tensor(ctx, device):
    """The context's integer accumulator (4 x uint32 per voxel) as an int32 torch tensor sharing memory.
    int32 two's-complement addition is bit-identical to uint32 addition, and NCCL has no uint32 sum in torch."""
    import torch
    ptr, n = ctx.accum_buffer()
    return torch.as_tensor(_DevicePointer(ptr, n), device=device)


def allreduce_accumulator(acc, group=None):
    """The one exchange step of triangle-sharded voxelisation: sum of the per-rank accumulators."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def pack_accumulator(counts: np.ndarray, sums: np.ndarray) -> np.ndarray:
    """(V,V,V) counts + (V,V,V,3) sums -> the device layout: per voxel two little-endian u64 words
    (r<<32 | g), (b<<32 | count), i.e. uint32 order [g, r, count, b]."""
    out = np.empty(counts.shape + (4,), dtype=np.uint32)
    out[..., 0] = sums[..., 1]
    out[..., 1] = sums[..., 0]
    out[..., 2] = counts
    out[..., 3] = sums[..., 2]
    return out


def unpack_accumulator(acc: np.ndarray):
    acc = acc.reshape(-1, 4)
    counts = acc[:, 2].copy()
    sums = np.stack([acc[:, 1], acc[:, 0], acc[:, 3]], -1)
    return counts, sums


def pack_exchange_records(counts, sums):
    """Host mirror of the inbox record format (vox_push_inbox, csrc/vct_voxelize.cu): one uint32[4] per voxel with
    count > 0 = {r | c0<<24, g | c1<<24, b | c2<<24, voxel}, 24-bit channel sums and a 24-bit count split into bytes
    c0..c2.  Values of 2^24 or more do not fit (the library reports VCT_ERR_OVERFLOW; here OverflowError)."""
    c = np.ascontiguousarray(counts, dtype=np.uint32).reshape(-1)
    s = np.ascontiguousarray(sums, dtype=np.uint32).reshape(-1, 3)
    v = np.nonzero(c)[0].astype(np.uint32)
    cc, ss = c[v], s[v]
    if (cc >> 24).any() or (ss >> 24).any():
        raise OverflowError("a voxel sum or count needs more than 24 bits")
    rec = np.empty((len(v), 4), dtype=np.uint32)
    for ch in range(3):
        rec[:, ch] = ss[:, ch] | (((cc >> (8 * ch)) & 0xFF) << 24)
    rec[:, 3] = v
    return rec


def merge_exchange_records(counts, sums, records):
    """Adds another rank's records into (counts, sums) in place (vox_merge_inbox): a rank's records name distinct
    voxels, so plain indexed adds suffice."""
    r = np.asarray(records, dtype=np.uint32).reshape(-1, 4)
    c = counts.reshape(-1); s = sums.reshape(-1, 3)
    v = r[:, 3]
    assert len(np.unique(v)) == len(v)
    n = (r[:, 0] >> 24) | ((r[:, 1] >> 24) << 8) | ((r[:, 2] >> 24) << 16)
    c[v] += n
    for ch in range(3):
        s[v, ch] += r[:, ch] & 0xFFFFFF
    return counts, sums


def default_session():
    """A session name every rank of one launch agrees on (torchrun exports MASTER_ADDR / MASTER_PORT)."""
    import os
    return f"{os.environ.get('MASTER_ADDR', 'local')}_{os.environ.get('MASTER_PORT', '0')}"


class SharedAccumulator:
    """Sharded voxelisation / sharded frames through the LIBRARY's multi-GPU layer (vct_comm_init, csrc/vct_comm.cu):
    symmetric segments, NVSwitch multicast mapping, handle exchange between the processes and the device-side barrier
    all live behind the C ABI; nothing here touches torch.  rank / world default to the torchrun environment
    (RANK / WORLD_SIZE), a world of one works on a single GPU.
    exchange = "inbox" (default; touched voxels as records, multimem.st) or "reduce" (multimem.red into a dense
    symmetric accumulator).  flags: capi.COMM_NO_MULTICAST, capi.COMM_KEEP_SHARES."""

    def __init__(self, ctx, device=None, group=None, exchange="inbox", rank=None, world=None, session=None, flags=0):
        import os
        self.ctx = ctx
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else int(world)
        self.reduce = exchange == "reduce"
        ctx.set_i("SharedExchange", 1 if self.reduce else 0)
        ctx.comm_init(self.rank, self.world, session or default_session(), flags)
        self.info = ctx.comm_info()

    def barrier(self):
        self.ctx.comm_barrier()

    def frame(self, host_rgba=None):
        """One sharded frame, pipelined inside the library (vct_frame_sharded): rows of every rank land in rank 0's
        frame ring; rank 0 receives the frame in `host_rgba`.  Asynchronous: call wait() before reading."""
        self.ctx.frame_sharded(host_rgba)

    def wait(self):
        self.ctx.frame_sharded_wait()

    def frame_voxels(self, tri_begin, tri_end):
        """One sharded voxelisation, serial form: voxelise this rank's triangle range and multicast what it touched,
        barrier, merge / resolve + mip the local copy.  (The inbox is double buffered by frame parity, so one barrier
        per frame is enough; the in-switch reduction needs a second one before the next frame may add into the
        accumulator.)"""
        self.ctx.voxelize_shared(tri_begin, tri_end)
        self.barrier()
        self.ctx.resolve_shared()
        if self.reduce:
            self.barrier()

    def close(self):
        self.ctx.comm_destroy()
