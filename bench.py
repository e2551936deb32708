#!/usr/bin/env python
"""bench.py -- full frames/s (voxelize + mip + cone-trace) of the B200 path on BASELINE.json config 2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode views|tiles|trishard]

One process per GPU (torchrun for N > 1).  A step is one full frame: sparse clear -> voxelise (coverage +
PCF-lit shading) -> resolve -> mip pyramid -> primary visibility -> cone trace, on the synthetic 260 K-triangle
atrium at 256^3 / 1920x1080 (`config.workload`).  The shadow map is static (as in the reference, which draws it
once at init) and is not part of the step.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "full frames/s (voxelize+mip+cone-trace), 256^3 grid, 1920x1080"
UNIT = "frames/s"
WORKLOAD = ("config2: synthetic Sponza-scale atrium 259608 tris (seed 1234), 256^3 RGBA8 grid, 1920x1080, "
            "6 diffuse + 1 specular cone (reference table), 2 bounces (reference semantics), conservative coverage, "
            "4096^2 shadow map")


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=None, choices=["views", "tiles", "trishard", "shard"],
                    help="N>1: views = one camera per rank, grid replicated (weak scaling, default for config 2); tiles = row bands, "
                         "voxelisation replicated; trishard = triangle ranges + NCCL all-reduce of the accumulator; shard = triangle "
                         "ranges exchanged over NVSwitch multicast by the library's own kernels (see --exchange) + row bands")
    ap.add_argument("--serial-shard", action="store_true", help="--mode shard: voxelise / exchange / trace back to back on one stream "
                    "instead of the pipelined vct_frame_shared_begin/_end")
    ap.add_argument("--contiguous", action="store_true", help="triangle sharding by contiguous ranges instead of interleaved blocks")
    ap.add_argument("--exchange", default="inbox", choices=["inbox", "reduce"],
                    help="--mode shard: inbox = touched voxels multicast as records (multimem.st) and merged locally; reduce = "
                         "multimem.red into a dense symmetric accumulator (reduced in the switch)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = headline (default); 3 = 512^3 fp16 grid, 4K, 9+1 cones, row bands; "
                         "4 = dynamic 1M-triangle mesh re-voxelised every frame, triangle-sharded + all-reduce; 5 = light-probe bake: 64 "
                         "camera views at 1024^2 from one 3-bounce voxelisation per bake, views round-robin over the ranks")
    ap.add_argument("--detail", type=float, default=1.0, help="scene tessellation scale (1.0 = config 2)")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--coverage", default="conservative")
    ap.add_argument("--cones", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flush", action="store_true", help="flush L2 between timed steps (per-step events, frames not pipelined); "
                    "default: no flush -- the per-frame working set (~220 MB, two alternating frame slots) exceeds the 126 MB L2")
    a = ap.parse_args(argv)
    if a.config == 3:
        a.grid, a.width, a.height = 512, 3840, 2160
        a.cones = a.cones or "9+1"
        a.mode = a.mode or ("shard" if a.gpus > 1 else "views")
    elif a.config == 4:
        a.mode = a.mode or ("shard" if a.gpus > 1 else "views")
    elif a.config == 5:
        a.width, a.height, a.mode = 1024, 1024, "probes"
    a.cones = a.cones or "6+1"
    a.mode = a.mode or "views"
    return a


def make_scene_and_uniforms(args):
    import vct_b200  # noqa: F401
    from vct_b200 import scenes, uniforms
    if args.config == 4:
        sc = scenes.dynamic_knot()
    else:
        sc = scenes.atrium(detail=args.detail)
    u = uniforms.scene_uniforms(sc, V=args.grid, width=args.width, height=args.height, shadow_map_size=4096,
                                coverage=args.coverage, cones=args.cones, grid_format=1 if args.config == 3 else 0,
                                bounces=3 if args.config == 5 else 2)
    return sc, u


def camera_for(step, rank):
    """Per-step input: a slow camera pan (every step has a new view matrix, as the reference's loop does)."""
    # every rank renders its own stream of frames: the same pan, 17 steps apart per rank (distinct frames of
    # near-identical cost, so that per-GPU work stays fixed as N grows = weak scaling)
    yaw = -90.0 + 0.05 * (step + 17 * rank)
    pos = (0.0, 4.0, 0.0)
    return pos, yaw


def set_camera(ctx, args, step, rank):
    import vct_b200.glmath as gm
    pos, yaw = camera_for(step, rank)
    view = gm.view_matrix(pos, yaw, 0.0)
    ctx.set_mat4("ModelViewMatrix", gm.colmajor((view @ gm.scale(0.05)).astype(np.float32)))
    ctx.set_3f("CameraPosition", pos)
    return 64 + 12          # bytes of per-step input handed to the device (as kernel parameters)


class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import vct_b200
    from vct_b200 import parallel

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)          # one explicit stream for our kernels, torch's events and NCCL
    torch.cuda.set_stream(stream)

    vct_b200.load_library()                         # mandatory extension: raises if missing
    sc, u = make_scene_and_uniforms(args)
    ctx = vct_b200.Context(local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_uniforms(u)
    ctx.load_scene(sc)
    H = args.height
    band = None
    if args.mode in ("tiles", "shard") and world > 1:
        b0, b1, per = parallel.row_band_equal(H, rank, world)     # equal nominal bands: ONE all-gather per frame
        ctx.set_i("RowBegin", b0); ctx.set_i("RowEnd", b1)
        band = (b0, b1, per)
    ctx.draw_depth()                                # static light: once, like the reference's init
    ctx.sync()
    tri_rng = (parallel.triangle_share(ctx, sc.n_tris, rank, world, interleave=not args.contiguous)
               if args.mode in ("trishard", "shard") else None)
    shared = parallel.SharedAccumulator(ctx, dev, exchange=args.exchange) if args.mode == "shard" else None
    acc = parallel.accumulator_tensor(ctx, dev) if args.mode == "trishard" else None
    gather_buf = None
    if band is not None:
        fptr, fbytes = ctx.frame_buffer()
        frame_t = torch.as_tensor(parallel._DevicePointer(fptr, fbytes, "|u1"), device=dev)
        chunk = band[2] * args.width * 4
        gather_buf = torch.zeros(world * chunk, dtype=torch.uint8, device=dev)    # bands in frame order, padded at the end
        my_band = torch.zeros(chunk, dtype=torch.uint8, device=dev)
        own0, own1 = band[0] * args.width * 4, band[1] * args.width * 4

    dyn = None
    if args.config == 4:
        # the mesh is re-generated on the device every frame (time-varying sine displacement along the normal)
        base = torch.from_numpy(np.ascontiguousarray(sc.verts[:, :3])).to(dev)
        nrm = torch.from_numpy(np.ascontiguousarray(sc.verts[:, 3:6])).to(dev)
        phase = (base[:, 0] * 0.004 + base[:, 2] * 0.003)
        dyn = (base, nrm, phase, torch.empty_like(base))

    probes = None
    if args.config == 5:
        import vct_b200.glmath as gm
        from vct_b200 import scenes as _scenes
        cams = _scenes.probe_cameras(64)
        probes = []
        for k in parallel.views_for_rank(len(cams), rank, world):
            pos, yaw, pitch = cams[k]
            view = gm.view_matrix(pos, yaw, pitch)
            probes.append((gm.colmajor((view @ gm.scale(0.05)).astype(np.float32)), tuple(float(x) for x in pos)))
        probe_hosts = [torch.empty((args.height, args.width, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]

    def bake(to_host):
        """Config 5: one 3-bounce voxelisation, then this rank's probe views rendered from it."""
        ctx.draw_voxels()
        for n, (mv, pos) in enumerate(probes):
            ctx.set_mat4("ModelViewMatrix", mv); ctx.set_3f("CameraPosition", pos)
            ctx.render(probe_hosts[n & 1] if to_host else None)

    def prepare(i):
        """Per-step input: the camera (configs 2, 3) or the re-generated mesh and its shadow map (config 4)."""
        cam_rank = rank if args.mode == "views" else 0
        if args.config == 5:
            return 76 * len(probes)
        if args.config != 4:
            return set_camera(ctx, args, i, cam_rank)
        base, nrm, phase, out = dyn
        torch.add(base, nrm * (12.0 * torch.sin(phase + 0.21 * i)).unsqueeze(1), out=out)
        ctx.update_positions(device_ptr=out.data_ptr(), n_verts=out.shape[0])
        ctx.draw_depth()                            # the light-space depth map follows the mesh
        return 4                                    # the time parameter; the mesh itself is generated on the device

    def step(i, host_out=None):
        prepare(i)
        if probes is not None:
            bake(host_out is not None)
            return
        if args.mode == "trishard":
            ctx.voxelize_range(tri_rng[0], tri_rng[1], clear_first=True)
            parallel.allreduce_accumulator(acc)
            ctx.resolve_and_mip()
            ctx.render(host_out)
        else:
            if args.mode == "shard":
                if args.exchange == "inbox" and not args.serial_shard:
                    shared.frame(tri_rng[0], tri_rng[1])      # pipelined: exchange on the library's voxel stream
                else:
                    shared.frame_voxels(tri_rng[0], tri_rng[1])   # exchange over NVSwitch multicast (multimem.st inbox / multimem.red)
                    ctx.render(None)
            else:
                ctx.frame(None if gather_buf is not None else host_out)
            if gather_buf is not None:              # row bands -> every rank holds the full frame
                my_band[:own1 - own0].copy_(frame_t[own0:own1])
                dist.all_gather_into_tensor(gather_buf, my_band)
                if host_out is not None:
                    host_out.view(-1).copy_(gather_buf[:H * args.width * 4], non_blocking=False)
            elif host_out is not None and args.mode == "shard":
                ctx.sync()
                host_out.copy_(torch.from_numpy(ctx.read_frame()))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush else None   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    # (1) timed region: per-pass event recording off (it costs ~30 us/frame).  K steps between one pair of CUDA
    # events on the launching stream, barrier + synchronize on both sides.  Consecutive frames pipeline inside the
    # library (the next frame's voxel/visibility stages run beside cone_trace); every frame does all of its work.
    ctx.set_i("Profile", 0)
    ctx.set_i("PipelineFrames", 0 if args.flush else 1)
    launches0 = ctx.kernel_launches()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
        e1.record(stream)
        barrier()
        total_ms = e0.elapsed_time(e1)
    else:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)                   # untimed: evicts the previous frame's lines from L2
            ev[i][0].record(stream)
            step(args.warmup + i)
            ev[i][1].record(stream)
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop()
    launches = ctx.kernel_launches() - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    frames = args.steps * (world if args.mode == "views" else 64 if args.mode == "probes" else 1)
    value = frames / (total_ms * 1e-3)
    # (2) per-pass breakdown and the dominant kernel's average duration: same steps again with pass events on
    ctx.set_i("Profile", 1)
    pass_names = ["vox_clear", "vox_cover", "vox_shade", "resolve", "mip", "visibility", "cone"]
    if args.mode == "shard":
        pass_names.insert(3, "exchange_push")
        if args.exchange == "inbox":
            pass_names.insert(4, "exchange_merge")
    pass_sum = {p: 0.0 for p in pass_names}
    samples_sum = 0
    n_prof = min(args.steps, 30)
    for i in range(n_prof):
        if flush is not None:
            flush.fill_(i & 0xFF)
        step(args.warmup + i)
        torch.cuda.synchronize()
        for p in pass_names:
            try:
                pass_sum[p] += ctx.pass_time_us(p)
            except Exception:
                pass
        samples_sum += ctx.cone_samples()
    ctx.set_i("Profile", 0)
    barrier()
    passes_max = None
    if world > 1 and args.mode != "views":         # sharded work: the slowest rank sets the pace of every phase
        pass_t = torch.tensor([pass_sum[p] / n_prof for p in pass_names], dtype=torch.float64, device=dev)
        dist.all_reduce(pass_t, op=dist.ReduceOp.MAX)
        passes_max = {p: round(float(v), 2) for p, v in zip(pass_names, pass_t.tolist())}

    # ---- end to end through the public API with HOST frame buffers (pinned); every frame's D2H copy is inside
    # the timed region.  Render loops use the pipelined call (vct_frame_async / vct_frame_wait: double-buffered,
    # the copy of frame i overlaps the rendering of frame i+1 -- what glfwSwapBuffers gives the reference's loop).
    hosts = [torch.empty((args.height, args.width, 4), dtype=torch.uint8).pin_memory() for _ in range(3)]
    pipelined = args.mode == "views"
    for i in range(3):
        step(i, hosts[0])
    if pipelined:                                   # warm the asynchronous path too (ring buffers, copy stream)
        for i in range(6):
            ctx.frame_async(hosts[i % 3])
            if i >= 2:
                ctx.frame_wait()
        ctx.frame_wait(); ctx.frame_wait()
    barrier()
    t0 = time.perf_counter()
    h2d = 0
    cam_rank = rank if args.mode == "views" else 0
    for i in range(args.steps):
        if pipelined:
            h2d = prepare(args.warmup + i)
            ctx.frame_async(hosts[i % 3])
            if i >= 2:
                ctx.frame_wait()                    # frame i-2 has arrived in host memory (two frames stay queued)
        else:
            h2d = 76 * len(probes) if probes is not None else 76 if args.config != 4 else 4
            step(args.warmup + i, hosts[i % 3])     # returns after the frame is in host memory
    if pipelined:
        ctx.frame_wait(); ctx.frame_wait()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("VCT_BENCH_DEBUG"):
        print(f"[rank {rank}] device loop {total_ms / args.steps:.4f} ms/step (max over ranks), e2e loop {e2e_s / args.steps * 1e3:.4f} ms/step", file=sys.stderr, flush=True)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = frames / float(te.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peaks_kind = measured_peaks()
    K = args.steps
    cone_us = pass_sum["cone"] / n_prof
    samples_per_launch = samples_sum / n_prof
    tex_peak_frac_lod = ctx.bench_tex3d(V=args.grid, n_samples=1 << 28, pattern=0, lod=0.5, iters=3)
    tex_peak_int_lod = ctx.bench_tex3d(V=args.grid, n_samples=1 << 28, pattern=0, lod=0.0, iters=3)
    achieved_gs = samples_per_launch / (cone_us * 1e-6) * 1e-9 if cone_us > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "cone_trace_traffic.json")
    if os.path.exists(prof):
        traffic = json.load(open(prof)).get("dram_bytes_per_launch")
    mip_us = pass_sum["mip"] / n_prof
    mip_bytes = sum((args.grid >> l) ** 3 * 4 for l in range(args.grid.bit_length()))     # read L0 once + write L1..: 73.14 MiB @256
    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": round(total_ms / K, 4), "higher_is_better": True,
        "scaling": "weak" if args.mode == "views" else "strong",     # probes / tiles / shards: total work fixed "vs_baseline": None, "dtype": "u8/u32 grid, f32 shading",
        "data": "synthetic",
        "config": {"workload": WORKLOAD if (args.config == 2 and args.detail == 1.0 and args.grid == 256) else f"config{args.config}: {sc.name} {sc.n_tris} tris, V={args.grid} {'RGBA16F' if args.config == 3 else 'RGBA8'}, {args.width}x{args.height}, cones {args.cones}",
                   "mode": args.mode + ("/" + args.exchange if args.mode == "shard" else ""), "l2": ("flushed between steps by an untimed 256 MiB write; frames not pipelined" if args.flush else "no flush: per-frame working set ~220 MB (64 MiB level 0 + mips, shadow texels, accumulator lines, queues, vertex cache, visibility) in two alternating frame slots exceeds the 126 MB L2"),
                   "timing": "one CUDA-event pair around the K steps on the launching stream, barrier+synchronize both sides; max over ranks",
                   "step": ("one bake = clear+voxelize+resolve+mip+reinject+mip once, then visibility+cone-trace of 64 probe views (round-robin over ranks); value counts views"
                            if args.mode == "probes" else "clear+voxelize+resolve+mip+visibility+cone-trace (shadow map static, drawn once)")},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": args.width * args.height * 4 * (len(probes) if probes is not None else 1),
                "note": ("vct_frame_async(host_rgba)+vct_frame_wait, three pinned host frame buffers (two frames queued), every frame copied to the host inside the timed region; per-step input = view matrix + camera position (kernel parameters)"
                         if pipelined else "each step returns after its frame(s) reached pinned host memory (synchronous copy inside the timed region)")},
        "gpu_launches": int(launches),
        "passes_us": {p: round(pass_sum[p] / n_prof, 2) for p in pass_names},
        "passes_us_max_over_ranks": passes_max,
        "cone_samples_per_frame": int(samples_per_launch),
        "gcone_samples_per_s": round(achieved_gs, 2),
        # dominant kernel.  It is bound by the texture pipe (tex3DLod wavefronts), not by HBM (28 MB of DRAM traffic per
        # launch: the pyramid is L1/L2 resident) and not by tensor cores (no dense contraction in this path), so
        # "bound" says "texture"; achieved/peak are ALGORITHMIC texel bytes (64 B = 2 levels x 8 texels x 4 B per
        # sample) per second, the peak being the tex3DLod rate measured in this run x 64 B.
        "roofline": {"kernel": "cone_trace", "bound": "texture", "achieved": round(achieved_gs * 64.0, 1),
                     "peak": round(tex_peak_frac_lod * 64.0, 1), "unit": "GB/s",
                     "frac": round(achieved_gs / tex_peak_frac_lod, 4) if tex_peak_frac_lod else None,
                     "traffic": traffic,
                     "achieved_gsamples_per_s": round(achieved_gs, 2), "peak_gsamples_per_s": round(tex_peak_frac_lod, 2),
                     "peak_single_level_gsamples_per_s": round(tex_peak_int_lod, 2),
                     "peak_source": "vct_bench_tex3d measured in this run: trilinear + mip-linear tex3DLod, RGBA8 256^3 pyramid, coherent walks (L2-resident)",
                     "algorithmic_bytes_per_sample": 64, "samples_per_launch": int(samples_per_launch),
                     "kernel_us": round(cone_us, 2)},
        "roofline_mip": {"kernel": "mip_fused3+mip_tail", "bound": "hbm",
                         "achieved": round(mip_bytes / (mip_us * 1e-6) * 1e-9, 1) if mip_us > 0 else None,
                         "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                         "frac": round(mip_bytes / (mip_us * 1e-6) * 1e-9 / peaks.get("hbm_gbs"), 4) if mip_us > 0 else None,
                         "peak_source": f"MEASURED_PEAKS.json ({peaks_kind})", "algorithmic_bytes": mip_bytes},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, sc, u, budget_s=15.0)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU legs
def host_cores():
    """threads the OpenMP oracle actually gets (affinity mask), not the machine total"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count()


def cpu_baseline(args, sc, u, budget_s):
    """The oracle ("port": the reference's GLSL cannot run here) timed on the host cores, on a
    bounded sample of the same workload: as many full config frames as fit in ~budget_s (at least one)."""
    from oracle.oracle_py import Oracle
    o = Oracle(); o.set_uniforms(u); o.load_scene(sc)
    o.draw_depth()
    t0 = time.perf_counter(); n = 0
    while True:
        o.draw_voxels(); o.render(); n += 1
        if time.perf_counter() - t0 > budget_s or n >= 50:
            break
    dt = time.perf_counter() - t0
    o.close()
    return {"value": round(n / dt, 4), "unit": UNIT, "cores": host_cores(), "kind": "port",
            "sample": f"{n} full frame(s) of the same workload (voxelize+mip+cone-trace) in {dt:.1f} s, OpenMP over all host cores"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle.oracle_py import Oracle
    sc, u = make_scene_and_uniforms(args)
    o = Oracle(); o.set_uniforms(u); o.load_scene(sc)
    o.draw_depth()
    H = args.height
    # size each step so that the whole run fits in ~150 s: full voxelisation + a band of rows, scaled
    t0 = time.perf_counter(); o.draw_voxels(); t_vox = time.perf_counter() - t0
    t0 = time.perf_counter(); o.render_rows(H // 2 - 8, H // 2 + 8); t_rows16 = time.perf_counter() - t0
    n_steps = args.steps + args.warmup
    per_step = max(150.0 / max(n_steps, 1), 0.05)
    rows = int(np.clip((per_step - t_vox) / max(t_rows16 / 16.0, 1e-6), 8, H)) // 8 * 8
    rows = max(8, min(rows, H))
    y0 = (H - rows) // 2
    for i in range(args.warmup):
        o.draw_voxels(); o.render_rows(y0, y0 + rows)
    tv = tr = 0.0
    for i in range(args.steps):
        t0 = time.perf_counter(); o.draw_voxels(); t1 = time.perf_counter(); o.render_rows(y0, y0 + rows); t2 = time.perf_counter()
        tv += t1 - t0; tr += t2 - t1
    K = max(args.steps, 1)
    frame_s = tv / K + (tr / K) * (H / rows)        # time of a full frame extrapolated from the row band
    value = 1.0 / frame_s
    sample = (f"each step = full voxelize+mip ({tv / K * 1e3:.0f} ms) + cone trace of {rows} of {H} rows "
              f"({tr / K * 1e3:.0f} ms), extrapolated linearly to the full frame; CPU oracle (port), OpenMP, {host_cores()} cores")
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(frame_s * 1e3, 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 grid, f32 shading", "data": "synthetic",
           "config": {"workload": WORKLOAD, "mode": "cpu"},
           "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
