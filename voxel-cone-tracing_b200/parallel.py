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


class SharedAccumulator:
    """Symmetric exchange buffer for the fused triangle-sharded voxelisation (vct_voxelize_shared / vct_resolve_shared).

    Plumbing only: the memory comes from torch symmetric memory (same allocation size on every rank, mapped on the
    peers and through an NVSwitch multicast address), the barrier is the symmetric-memory signal-pad barrier on the
    current stream.  With world size 1 it is a plain zeroed device buffer and the barrier is a no-op.
    exchange = "inbox" (default) or "reduce" (multimem.red into a dense symmetric accumulator)."""

    def __init__(self, ctx, device, group=None, exchange="inbox"):
        import torch
        import torch.distributed as dist
        self.ctx = ctx
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.reduce = exchange == "reduce"
        self._xstream = None
        ctx.set_i("SharedExchange", 1 if self.reduce else 0)
        ctx.set_i("SharedWorld", self.world)
        ctx.set_i("SharedRank", self.rank)
        nbytes = ctx.shared_accum_bytes()
        n64 = (nbytes + 7) // 8
        self.hdl = None
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(n64, dtype=torch.int64, device=device)
            g = group or dist.group.WORLD
            self.hdl = symm_mem.rendezvous(self.buf, group=g.group_name if hasattr(g, "group_name") else g)
            self.buf.zero_()
            torch.cuda.synchronize(device)
            self.hdl.barrier()
            mc = int(self.hdl.multicast_ptr or 0)
            if not mc:
                raise RuntimeError("no multicast mapping for symmetric memory on this system")
            ctx.set_shared_accum(self.buf.data_ptr(), mc)
        else:
            self.buf = torch.zeros(n64, dtype=torch.int64, device=device)
            ctx.set_shared_accum(self.buf.data_ptr(), 0)

    def barrier(self):
        if self.hdl is not None:
            self.hdl.barrier()

    def frame(self, tri_begin, tri_end, host_rgba=None):
        """One sharded frame, pipelined (vct_frame_shared_begin / _end; inbox exchange only): the cross-rank barrier is
        enqueued on the library's exchange stream, so the voxel / exchange / visibility stages of the next frame run
        beside this frame's cone_trace.  Renders rows RowBegin..RowEnd of this rank."""
        import torch
        self.ctx.frame_shared_begin(tri_begin, tri_end)
        if self.hdl is not None:
            if self._xstream is None:
                self._xstream = torch.cuda.ExternalStream(self.ctx.exchange_stream(), device=self.buf.device)
            with torch.cuda.stream(self._xstream):
                self.hdl.barrier()
        self.ctx.frame_shared_end(host_rgba)

    def frame_voxels(self, tri_begin, tri_end):
        """One sharded voxelisation: voxelise this rank's triangle range and multicast what it touched, barrier, merge /
        resolve + mip the local copy.  (The inbox is double buffered by frame parity, so one barrier per frame is
        enough; the in-switch reduction needs a second one before the next frame may add into the accumulator.)"""
        self.ctx.voxelize_shared(tri_begin, tri_end)
        self.barrier()
        self.ctx.resolve_shared()
        if self.reduce:
            self.barrier()
