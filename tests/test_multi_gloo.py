"""world_size-2 gloo test of the triangle-sharded voxelisation exchange (BASELINE config 4):
each rank accumulates its triangle range, the uint32 accumulators are summed with all_reduce, and the
result must be bit-identical to the unsharded one.  The oracle stands in for the device here (CPU box)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import vct_b200  # noqa: F401
    from vct_b200 import parallel, scenes, uniforms
    from oracle.oracle_py import Oracle
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = scenes.atrium(detail=0.12, tex_size=32)
    u = uniforms.scene_uniforms(sc, V=32, width=64, height=64, shadow_map_size=256, coverage="conservative")
    o = Oracle(); o.set_uniforms(u); o.load_scene(sc)
    o.draw_depth()
    tb, te = parallel.triangle_range(sc.n_tris, rank, world)
    o.draw_voxels_range(tb, te, clear_first=True)
    acc = parallel.pack_accumulator(o.counts(), o.sums())
    t = torch.from_numpy(acc.view(np.int32).reshape(-1))
    parallel.allreduce_accumulator(t)
    counts, sums = parallel.unpack_accumulator(t.numpy().view(np.uint32))
    o.set_accum(counts, sums)
    o.resolve_and_mip()
    # rows of the frame: every rank renders its band, bands are gathered on rank 0
    o.render()
    b0, b1 = parallel.row_band(64, rank, world)
    band = torch.from_numpy(o.frame()[b0:b1].copy())
    bands = [torch.empty_like(band) for _ in range(world)] if rank == 0 else None
    dist.gather(band, bands, dst=0)
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), counts=counts, sums=sums, grid0=o.grid(0), grid3=o.grid(3),
                 frame=np.concatenate([b.numpy() for b in bands], 0))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_triangle_sharded_voxelisation_allreduce_is_bit_exact(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    from vct_b200 import scenes, uniforms
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npz")
    sc = scenes.atrium(detail=0.12, tex_size=32)
    u = uniforms.scene_uniforms(sc, V=32, width=64, height=64, shadow_map_size=256, coverage="conservative")
    o = oracle_mod.Oracle(); o.set_uniforms(u); o.load_scene(sc)
    o.draw_depth(); o.draw_voxels(); o.render()
    assert np.array_equal(got["counts"], o.counts().reshape(-1))
    assert np.array_equal(got["sums"], o.sums().reshape(-1, 3))
    assert np.array_equal(got["grid0"], o.grid(0)) and np.array_equal(got["grid3"], o.grid(3))
    assert np.array_equal(got["frame"], o.frame())
    assert got["counts"].sum() > 1000


def _worker_inbox(rank, world, port, out_dir):
    """CPU model of the fused exchange (vct_voxelize_shared / vct_resolve_shared, inbox flavour): interleaved triangle
    shares, every rank sends 16-byte records of the voxels it touched to all others, merges what it receives, resolves
    locally; equal row bands gathered by ONE all-gather."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import vct_b200  # noqa: F401
    from vct_b200 import parallel, scenes, uniforms
    from oracle.oracle_py import Oracle
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = scenes.atrium(detail=0.12, tex_size=32)
    H = W = 72
    u = uniforms.scene_uniforms(sc, V=32, width=W, height=H, shadow_map_size=256, coverage="conservative")
    o = Oracle(); o.set_uniforms(u); o.load_scene(sc)
    o.draw_depth()
    first = True
    for b0 in range(rank * 128, sc.n_tris, world * 128):      # TriangleInterleave = world, TrianglePhase = rank
        o.draw_voxels_range(b0, min(b0 + 128, sc.n_tris), clear_first=first)
        first = False
    counts, sums = o.counts().copy(), o.sums().copy()
    mine = parallel.pack_exchange_records(counts, sums)
    n = torch.tensor([len(mine)], dtype=torch.int64)
    ns = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(ns, n)
    cap = int(max(int(x) for x in ns))
    padded = torch.zeros((cap, 4), dtype=torch.int32)
    padded[:len(mine)] = torch.from_numpy(mine.view(np.int32))
    inbox = [torch.zeros((cap, 4), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(inbox, padded)                            # = the multicast store into every rank's inbox row
    for r in range(world):
        if r != rank:
            parallel.merge_exchange_records(counts, sums, inbox[r].numpy().view(np.uint32)[:int(ns[r])])
    o.set_accum(counts, sums)
    o.resolve_and_mip()
    o.render()
    y0, y1, per = parallel.row_band_equal(H, rank, world)
    band = torch.zeros((per, W, 4), dtype=torch.uint8)
    band[:y1 - y0] = torch.from_numpy(o.frame()[y0:y1].copy())
    full = torch.zeros((world * per, W, 4), dtype=torch.uint8)
    dist.all_gather_into_tensor(full, band)
    # the library's default row deal (vct_comm_init): interleaved 8-row strips, every rank stores its strips straight
    # into rank 0's frame slot.  Disjoint writes are modelled by a sum of frames that are zero outside the own strips.
    strips = parallel.row_strips_for_rank(H, rank, world)
    own = torch.zeros((H, W, 4), dtype=torch.int32)
    rows = torch.zeros(H, dtype=torch.int32)
    for a, b in strips:
        own[a:b] = torch.from_numpy(o.frame()[a:b].astype(np.int32))
        rows[a:b] += 1
    dist.all_reduce(own)
    dist.all_reduce(rows)
    np.savez(os.path.join(out_dir, f"inbox_{rank}.npz"), counts=counts, sums=sums, grid0=o.grid(0), grid3=o.grid(3),
             frame=full.numpy()[:H], n_records=np.array([int(x) for x in ns]),
             strip_frame=own.numpy().astype(np.uint8), strip_rows=rows.numpy(), n_strips=len(strips))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_inbox_exchange_model_is_bit_exact_on_every_rank(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    from vct_b200 import parallel, scenes, uniforms
    port = _free_port()
    mp.spawn(_worker_inbox, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sc = scenes.atrium(detail=0.12, tex_size=32)
    u = uniforms.scene_uniforms(sc, V=32, width=72, height=72, shadow_map_size=256, coverage="conservative")
    o = oracle_mod.Oracle(); o.set_uniforms(u); o.load_scene(sc)
    o.draw_depth(); o.draw_voxels(); o.render()
    for rank in range(2):
        got = np.load(tmp_path / f"inbox_{rank}.npz")
        assert np.array_equal(got["counts"], o.counts()) and np.array_equal(got["sums"], o.sums())
        assert np.array_equal(got["grid0"], o.grid(0)) and np.array_equal(got["grid3"], o.grid(3))
        assert np.array_equal(got["frame"], o.frame())
        assert 0 < got["n_records"].min() and got["n_records"].sum() >= int((o.counts() > 0).sum())
        assert np.array_equal(got["strip_rows"], np.ones(72, dtype=np.int32))      # every row owned exactly once
        assert np.array_equal(got["strip_frame"], o.frame())
        assert int(got["n_strips"]) == (5 if rank == 0 else 4)                    # 72 rows = 9 strips, dealt 5 + 4
    # record format: round trip and the 24-bit limit
    c = np.zeros(8, dtype=np.uint32); s = np.zeros((8, 3), dtype=np.uint32)
    c[3], s[3] = 65793, (65793 * 255, 1, 0)
    rec = parallel.pack_exchange_records(c, s)
    assert rec.shape == (1, 4) and rec[0, 3] == 3
    c2 = np.zeros(8, dtype=np.uint32); s2 = np.zeros((8, 3), dtype=np.uint32)
    parallel.merge_exchange_records(c2, s2, rec)
    assert np.array_equal(c2, c) and np.array_equal(s2, s)
    c[3], s[3] = 65794, (65794 * 255, 0, 0)                   # 16 777 470 >= 2^24
    with pytest.raises(OverflowError):
        parallel.pack_exchange_records(c, s)
