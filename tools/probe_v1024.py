"""Dense storage beyond 512^3: config-2 scene at V = 512 and V = 1024, RGBA8 and RGBA16F, 1080p.
Reports frames/s, per-pass times, device memory in use and sanity properties.  usage: probe_v1024.py [frames]"""
import json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import vct_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
names = ["vox_clear", "vox_cover", "vox_shade", "resolve", "mip", "visibility", "cone"]


def mem_used_mib():
    out = subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip()
    return int(out.splitlines()[0]) if out else -1


frames = {}
for V, fmt in ((512, 0), (1024, 0), (1024, 1)):
    a = bench.parse(["--config", "2"])
    sc, u = bench.make_scene_and_uniforms(a)
    u["VoxelDimensions"] = V
    u["GridFormat"] = fmt
    c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
    c.set_i("MaxFragments", 96 << 20); c.set_i("MaxTileItems", 16 << 20)      # ~30 M fragments at 1024^3
    c.draw_depth()
    for _ in range(3):
        c.frame(); c.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        c.frame()
    c.sync()
    dt = (time.perf_counter() - t0) / n
    img = c.read_frame()
    frames[(V, fmt)] = img
    c.set_i("PipelineFrames", 0); c.set_i("OverlapVisibility", 0)
    for _ in range(2):
        c.frame(); c.sync()
    W, H = int(u["screen_width"]), int(u["screen_height"])
    rec = dict(V=V, grid_format="RGBA16F" if fmt else "RGBA8", frames_per_s=round(1.0 / dt, 1), ms_per_frame=round(dt * 1e3, 3),
               passes_us={k: round(c.pass_time_us(k), 1) for k in names}, fragments=int(c.fragment_count()),
               occupied_voxels=int(c.occupied_voxels()), cone_samples_per_pixel=round(c.cone_samples() / (W * H), 2),
               device_mem_used_mib=mem_used_mib(), background_fraction=round(float((img[..., :3] == 128).all(-1).mean()), 4))
    print(json.dumps(rec), flush=True)
    c.close()

ref = frames[(512, 0)].astype(np.float64)
for key in ((1024, 0), (1024, 1)):
    mse = ((frames[key].astype(np.float64) - ref) ** 2).mean()
    print(f"V={key[0]} fmt={key[1]} vs V=512 RGBA8: psnr {10 * np.log10(255 ** 2 / max(mse, 1e-12)):.1f} dB (different grids: similarity only)")
print("RGBA8 vs RGBA16F at V=1024:", f"{10 * np.log10(255 ** 2 / max(((frames[(1024, 0)].astype(float) - frames[(1024, 1)].astype(float)) ** 2).mean(), 1e-12)):.1f} dB")
