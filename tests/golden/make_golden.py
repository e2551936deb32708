"""Regenerates the committed golden fixtures from the CPU oracle:  python tests/golden/make_golden.py

The reference cannot run here and reads nothing back, so there are no reference-made
vectors; these fixtures pin the oracle's own output (regression pins, checked on CPU by test_golden.py) and
give the GPU tests a committed target that does not depend on the oracle being rebuilt identically."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vct_b200  # noqa: E402,F401
from vct_b200 import scenes, uniforms  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (scene factory, uniform kwargs)
    "cornell_v32_msaa4": (lambda: scenes.cornell(), dict(V=32, width=96, height=96, shadow_map_size=512, coverage="msaa4")),
    "cornell_v32_center": (lambda: scenes.cornell(), dict(V=32, width=96, height=96, shadow_map_size=512, coverage="center")),
    "atrium_v32_conservative": (lambda: scenes.atrium(detail=0.1, tex_size=32),
                                dict(V=32, width=96, height=54, shadow_map_size=512, coverage="conservative")),
}


def run_case(name):
    factory, kw = CASES[name]
    sc = factory()
    u = uniforms.scene_uniforms(sc, **kw)
    o = Oracle(); o.set_uniforms(u); o.load_scene(sc)
    o.draw_depth(); o.draw_voxels(); o.render()
    d = o.depth()
    out = dict(depth_crc=np.uint32(zlib.crc32(d.tobytes())), depth_lit=np.uint32((d < 0xFFFFFF).sum()),
               counts=o.counts().astype(np.uint16), sums=o.sums(), grid0=o.grid(0), grid1=o.grid(1), grid2=o.grid(2),
               grid4=o.grid(4), visibility=o.visibility(), frame=o.frame(), cone_samples=np.uint64(o.cone_samples()))
    o.close()
    return sc, u, out


if __name__ == "__main__":
    for name in CASES:
        _, _, out = run_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else int(v)) for k, v in out.items()})
