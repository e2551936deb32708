"""Per-pass timings on config 2 (no oracle).  usage: perf_passes.py [frames]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sc = scenes.atrium()
u = uniforms.scene_uniforms(sc, V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.sync()
names = ["vox_clear", "vox_cover", "vox_shade", "resolve", "mip", "visibility", "cone", "frame"]
acc = {k: [] for k in names}
for i in range(n + 5):
    c.frame(); c.sync()
    if i >= 5:
        for k in names: acc[k].append(c.pass_time_us(k))
print("depth %.1f us" % c.pass_time_us("depth"))
print("  ".join(f"{k} {np.median(v):.1f}" for k, v in acc.items()))
tot = np.median(acc["frame"])
print(f"frame {tot:.1f} us -> {1e6 / tot:.1f} fps ; cone samples {c.cone_samples()} -> {c.cone_samples() / np.median(acc['cone']) * 1e-3:.1f} Gs/s; frags {c.fragment_count()} occupied {c.occupied_voxels()}")
c.set_i("Profile", 0)
for i in range(5): c.frame()
c.sync(); t0 = time.perf_counter()
N = 200
for i in range(N): c.frame()
c.sync(); dt = (time.perf_counter() - t0) / N
print(f"back-to-back throughput (pipelined, Profile=0): {dt * 1e6:.1f} us/frame -> {1 / dt:.1f} fps")
c.set_i("PipelineFrames", 0)
for i in range(5): c.frame()
c.sync(); t0 = time.perf_counter()
for i in range(N): c.frame()
c.sync(); dt = (time.perf_counter() - t0) / N
print(f"back-to-back throughput (PipelineFrames=0): {dt * 1e6:.1f} us/frame -> {1 / dt:.1f} fps")
c.set_i("PipelineFrames", 1); c.set_i("Profile", 1)
c.frame(); c.sync()
f = c.read_frame()
import zlib
print("frame crc", zlib.crc32(f.tobytes()), "grid0 crc", zlib.crc32(c.grid(0).tobytes()), "counts crc", zlib.crc32(c.counts().tobytes()))
