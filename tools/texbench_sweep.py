import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
c = vct_b200.Context(0)
for pattern in (0, 1):
    for lod in (0.0, 0.207, 0.5, 1.0, 1.314, 2.421, 3.0, 3.528, 4.635, 5.742, 6.85, 8.0):
        g = c.bench_tex3d(V=256, n_samples=1 << 27, pattern=pattern, lod=lod, iters=3)
        print(f"pattern {pattern} lod {lod:5.3f}: {g:7.1f} Gsamples/s")
