"""One rank's share of an N-way sharded frame, on ONE GPU (profiling aid): triangles and rows are dealt as for rank r of
N, the exchange runs against a world of one (so vox_merge_inbox has nothing to merge).  usage: rank_view.py <config> <N> [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import vct_b200
from vct_b200 import capi, parallel
cfg, N = sys.argv[1], int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 6
a = bench.parse(["--config", cfg])
sc, u = bench.make_scene_and_uniforms(a)
c = vct_b200.Context(0); c.set_uniforms(u)
for kv in filter(None, os.environ.get("VCT_TUNE", "").split(",")):
    k, v = kv.split("="); c.set_i(k, int(v))
c.load_scene(sc); c.draw_depth()
c.set_i("TriangleInterleave", N); c.set_i("TrianglePhase", 1); c.set_i("RowInterleave", N); c.set_i("RowPhase", 1)
shared = parallel.SharedAccumulator(c, rank=0, world=1, session="rv", flags=capi.COMM_KEEP_SHARES)
c.set_i("PipelineFrames", 1); c.set_i("Profile", 1)
for i in range(frames):
    bench.set_camera(c, sc, i, 0); shared.frame(None)
shared.wait()
names = ["vox_clear", "vox_cover", "vox_shade", "exchange_push", "resolve", "mip", "visibility", "cone"]
t0 = None
for back in (1, 0):
    row = []
    for n in names:
        b, e = c.pass_timeline(back, n)
        t0 = b if t0 is None else t0
        row.append(f"{n} {b - t0:6.0f}-{e - t0:6.0f}")
    print(f"frame -{back}: " + " | ".join(row))

import time
c.set_i("Profile", 0)
for i in range(5):
    bench.set_camera(c, sc, i, 0); shared.frame(None)
shared.wait(); t0 = time.perf_counter()
for i in range(60):
    bench.set_camera(c, sc, i, 0); shared.frame(None)
shared.wait(); dt = (time.perf_counter() - t0) / 60
print(f"steady-state period {dt * 1e6:.0f} us/frame  (tune: {os.environ.get('VCT_TUNE', '')})")
t0 = time.perf_counter()
for i in range(60):
    bench.set_camera(c, sc, i, 0); shared.frame(None)
t_enq = (time.perf_counter() - t0) / 60
shared.wait()
t0 = time.perf_counter()
for i in range(60):
    bench.set_camera(c, sc, i, 0)
t_cam = (time.perf_counter() - t0) / 60
print(f"host enqueue time {t_enq * 1e6:.0f} us/frame (of which camera set-up in Python {t_cam * 1e6:.0f} us)")
