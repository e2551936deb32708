"""Cost of each row band of a frame on one GPU (how well `tiles` / `shard` sharding balances).
usage: python tools/band_cost.py [--config 3] [--bands 8]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--bands", type=int, default=8)
    a = ap.parse_args()
    args = bench.parse(["--config", str(a.config)])
    import vct_b200
    from vct_b200 import parallel
    sc, u = bench.make_scene_and_uniforms(args)
    c = vct_b200.Context(0)
    c.set_uniforms(u); c.load_scene(sc)
    c.set_i("Profile", 1); c.set_i("PipelineFrames", 0)
    bench.set_camera(c, args, 0, 0)
    c.draw_depth(); c.draw_voxels(); c.sync()
    print("occupied voxels", c.occupied_voxels(), "fragments", c.fragment_count())
    H = args.height
    rows = []
    for nb in sorted({a.bands, 2}):
        for r in range(nb):
            b0, b1 = parallel.row_band(H, r, nb)
            c.set_i("RowBegin", b0); c.set_i("RowEnd", b1)
            t = []
            for it in range(5):
                c.render(None); c.sync()
                t.append((c.pass_time_us("visibility"), c.pass_time_us("cone")))
            t = np.median(np.array(t), axis=0)
            rows.append((nb, r, b0, b1, t[0], t[1]))
            print(f"bands={nb} rank={r} rows[{b0},{b1}) visibility {t[0]:.1f} us cone {t[1]:.1f} us", flush=True)
    for nb in sorted({a.bands, 2}):
        cone = [x[5] for x in rows if x[0] == nb]
        print(f"bands={nb}: cone max {max(cone):.1f} mean {np.mean(cone):.1f} -> balance {np.mean(cone) / max(cone):.3f}")


if __name__ == "__main__":
    main()
