"""tex3DLod throughput sweep (roofline denominators): pyramid size = residency (64^3 RGBA8 1 MiB ... 512^3 RGBA16F 1.17 GiB),
format, access pattern (coherent cone-like walks / random), one-level vs two-level filtering."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
c = vct_b200.Context(0)
print(f"{'V':>4s} {'format':8s} {'pyramid':>10s} {'pattern':9s} {'lod':>4s} {'Gsamples/s':>11s} {'algorithmic TB/s':>17s}")
for V in (64, 128, 256, 512):
    for fmt, name, bpt in ((0, "RGBA8", 4), (1, "RGBA16F", 8)):
        size = sum((V >> l) ** 3 * bpt for l in range(V.bit_length()))
        for pattern, pname in ((0, "coherent"), (1, "random")):
            for lod in (0.0, 0.5, 2.5):
                g = c.bench_tex3d(V=V, n_samples=1 << 27, pattern=pattern, lod=lod, iters=3, grid_format=fmt)
                per = (8 if lod == int(lod) else 16) * bpt
                print(f"{V:4d} {name:8s} {size / 2**20:8.1f}Mi {pname:9s} {lod:4.1f} {g:11.1f} {g * per / 1000:17.2f}", flush=True)
