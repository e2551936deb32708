"""torchrun worker (one process per GPU): fused triangle-sharded voxelisation through the symmetric/multicast
accumulator must give every rank the SAME grid as a single-GPU voxelisation, bit for bit; row-band frames gathered
from all ranks must equal the single-GPU frame."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vct_b200  # noqa: E402
from vct_b200 import parallel, scenes, uniforms  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sc = scenes.atrium(detail=0.3, tex_size=64)
    H, W = 360, 640
    u = uniforms.scene_uniforms(sc, V=128, width=W, height=H, shadow_map_size=2048, coverage="conservative")
    c = vct_b200.Context(local)
    c.set_stream(stream.cuda_stream)
    c.set_uniforms(u); c.load_scene(sc)
    c.draw_depth()
    # single-GPU reference on every rank
    c.draw_voxels(); c.render(); c.sync()
    ref_counts_occ = c.grid(0)[..., 3] > 0
    ref_grid = [c.grid(l) for l in range(8)]
    ref_frame = c.read_frame()
    # fused sharded path, three frames (exercises the mask-driven clear / stale-voxel removal on both slots)
    shared = parallel.SharedAccumulator(c, dev, exchange=sys.argv[1] if len(sys.argv) > 1 else "inbox")
    ok = True
    for it in range(4):
        tb, te = parallel.triangle_share(c, sc.n_tris, rank, world, interleave=it >= 2)   # contiguous, then interleaved
        shared.frame_voxels(tb, te)
        c.sync()
        for l in range(8):
            ok &= bool(np.array_equal(c.grid(l), ref_grid[l]))
    b0, b1 = parallel.row_band(H, rank, world)
    c.set_i("RowBegin", b0); c.set_i("RowEnd", b1)
    c.render(); c.sync()
    band = torch.from_numpy(c.read_frame()[b0:b1].copy()).to(dev)
    bands = [torch.empty((parallel.row_band(H, r, world)[1] - parallel.row_band(H, r, world)[0], W, 4), dtype=torch.uint8, device=dev)
             for r in range(world)]
    for r in range(world):
        src = band if r == rank else bands[r]
        dist.broadcast(src, src=r)
        if r == rank:
            bands[r] = band
    frame = torch.cat(bands, 0).cpu().numpy()
    ok &= bool(np.array_equal(frame, ref_frame))
    if not shared.reduce:
        # pipelined sharded frames (vct_frame_shared_begin / barrier on the exchange stream / _end)
        c.set_i("PipelineFrames", 1)
        for it in range(4):
            shared.frame(tb, te)
        c.sync()
        ok &= bool(np.array_equal(c.read_frame()[b0:b1], ref_frame[b0:b1]))
        for l in range(8):
            ok &= bool(np.array_equal(c.grid(l), ref_grid[l]))
    ok &= bool(ref_counts_occ.sum() > 10000)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_SHARED_OK" if int(flag.item()) == 1 else "MGPU_SHARED_MISMATCH", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
