"""One rank of a multi-process sharded run, driven ONLY through the C ABI (no torch, no NCCL): the library's own
vct_comm_init does the handle exchange, the multicast binding and the device-side barriers.

    python tests/mgpu_comm_worker.py <rank> <world> <session> <device> [nomc] [reduce]

Every rank first renders the single-GPU reference (grids + frame), then runs sharded frames with a moving camera; the
sharded grids must equal the reference on EVERY rank bit for bit and rank 0 must receive the reference frame in its
host buffer.  Prints MGPU_COMM_OK <rank> on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vct_b200  # noqa: E402
import vct_b200.glmath as gm  # noqa: E402
from vct_b200 import capi, parallel, scenes, uniforms  # noqa: E402


def main():
    rank, world, session, device = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    flags = (capi.COMM_NO_MULTICAST if "nomc" in sys.argv[5:] else 0) | (capi.COMM_ROW_BANDS if "bands" in sys.argv[5:] else 0)
    reduce = "reduce" in sys.argv[5:]
    sc = scenes.atrium(detail=0.3, tex_size=64)
    H, W, V = 360, 640, 128
    u = uniforms.scene_uniforms(sc, V=V, width=W, height=H, shadow_map_size=2048, coverage="conservative")
    c = vct_b200.Context(device)
    c.set_uniforms(u); c.load_scene(sc); c.draw_depth()

    def camera(i):
        view = gm.view_matrix(sc.camera_pos, sc.yaw + 4.0 * i, sc.pitch)
        c.set_mat4("ModelViewMatrix", gm.colmajor((view @ gm.scale(0.05)).astype(np.float32)))

    ref = []
    for i in range(5):
        camera(i); c.frame(); c.sync()
        ref.append(c.read_frame())
    ref_grid = [c.grid(l) for l in range(8)]
    ok = bool((ref_grid[0][..., 3] > 0).sum() > 10000)
    ref_depth = c.depth()
    shard_shadow = "shadow" in sys.argv[5:]
    if shard_shadow:
        c.set_i("ShardShadowMap", 1)            # before comm_init: the segment gets a D24 image

    shared = parallel.SharedAccumulator(c, rank=rank, world=world, session=session, flags=flags,
                                        exchange="reduce" if reduce else "inbox")
    info = shared.info
    assert info["rank"] == rank and info["world"] == world
    if shard_shadow:
        # triangle-sharded shadow map: every rank rasterises its share, the depth fragments are min-reduced into every
        # rank's image (multimem.red.min.u32 / peer atomics): the map must equal the single-GPU one on EVERY rank
        for _ in range(2):
            c.draw_depth(); c.sync()
            same = bool(np.array_equal(c.depth(), ref_depth))
            if not same:
                print(f"rank {rank}: sharded shadow map differs in {(c.depth() != ref_depth).sum()} texels", flush=True)
            ok &= same
    if reduce:
        for it in range(3):
            shared.frame_voxels(0, sc.n_tris)
            c.sync()
            for l in range(8):
                ok &= bool(np.array_equal(c.grid(l), ref_grid[l]))
    else:
        hosts = [np.zeros((H, W, 4), np.uint8) for _ in range(5)]
        for i in range(5):                       # five frames in flight through the three-slot ring
            camera(i)
            shared.frame(hosts[i] if rank == 0 else None)
        shared.wait()
        for l in range(8):
            ok &= bool(np.array_equal(c.grid(l), ref_grid[l]))
        if rank == 0:
            for i in range(5):
                same = bool(np.array_equal(hosts[i], ref[i]))
                if not same:
                    bad = np.argwhere((hosts[i] != ref[i]).any(-1))
                    print(f"rank 0 frame {i}: {len(bad)} pixels differ, rows {bad[:, 0].min()}..{bad[:, 0].max()}", flush=True)
                ok &= same
        # serial form on the same segments
        shared.frame_voxels(0, sc.n_tris); c.sync()
        for l in range(8):
            ok &= bool(np.array_equal(c.grid(l), ref_grid[l]))
    shared.barrier(); c.sync()
    print(("MGPU_COMM_OK" if ok else "MGPU_COMM_MISMATCH") + f" {rank} multicast={info['multicast']}", flush=True)
    shared.close(); c.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
