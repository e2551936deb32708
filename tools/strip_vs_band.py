"""cone_trace cost of one rank's rows at config 3: interleaved 8-row strips vs contiguous bands (one GPU, full grid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, vct_b200
a = bench.parse(["--config", sys.argv[1] if len(sys.argv) > 1 else "3"])
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sc, u = bench.make_scene_and_uniforms(a)
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc); c.draw_depth(); c.draw_voxels(); c.sync()
H = a.height
def cone_time():
    t = []
    for i in range(8):
        c.render(); c.sync()
        if i >= 3: t.append(c.pass_time_us("cone"))
    return float(np.median(t)), c.cone_samples()
full, n_full = cone_time()
print(f"whole frame: {full:.0f} us, {n_full} samples -> ideal 1/{N} share {full / N:.0f} us")
strips = []
for r in range(N):
    c.set_i("RowInterleave", N); c.set_i("RowPhase", r)
    strips.append(cone_time())
c.set_i("RowInterleave", 0); c.set_i("RowPhase", 0)
print("strips:", " ".join(f"{t:.0f}" for t, _ in strips), f"| max {max(t for t, _ in strips):.0f} sum {sum(t for t, _ in strips):.0f}")
bands = []
per = ((H + 7) // 8 + N - 1) // N * 8
for r in range(N):
    c.set_i("RowBegin", min(r * per, H)); c.set_i("RowEnd", min((r + 1) * per, H))
    bands.append(cone_time())
print("bands: ", " ".join(f"{t:.0f}" for t, _ in bands), f"| max {max(t for t, _ in bands):.0f} sum {sum(t for t, _ in bands):.0f}")
