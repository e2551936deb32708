#!/usr/bin/env python
"""Render a scene to a PNG through the public API (needs a CUDA device).

    python tools/render_scene.py --scene atrium --grid 256 --size 1920x1080 --out frame.png
    python tools/render_scene.py --obj path/to/model.obj --scale 0.05 --camera 0,4,0 --yaw -90

The reference's loop (main.cpp:77-94) with a file instead of a window."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200  # noqa: E402
from vct_b200 import objloader, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell", choices=["cornell", "atrium", "knot"])
    ap.add_argument("--obj", default=None, help="Wavefront OBJ (+MTL) instead of a built-in scene; model units as in the file")
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--size", default="1280x720")
    ap.add_argument("--camera", default=None, help="x,y,z world position")
    ap.add_argument("--yaw", type=float, default=None)
    ap.add_argument("--pitch", type=float, default=None)
    ap.add_argument("--coverage", default="msaa4", choices=["center", "msaa4", "conservative"])
    ap.add_argument("--bounces", type=int, default=2)
    ap.add_argument("--out", default="frame.png")
    a = ap.parse_args()
    w, h = (int(x) for x in a.size.lower().split("x"))
    sc = objloader.load_obj(a.obj) if a.obj else {"cornell": scenes.cornell, "atrium": scenes.atrium, "knot": scenes.dynamic_knot}[a.scene]()
    cam = vct_b200.Camera(tuple(float(x) for x in a.camera.split(",")) if a.camera else sc.camera_pos,
                          a.yaw if a.yaw is not None else sc.yaw, a.pitch if a.pitch is not None else sc.pitch, sc.fov_deg)
    r = vct_b200.Voxel_Cone_Tracing(w, h, None, device=0, VoxelDimensions=a.grid, camera=cam)
    r.CoveragePolicy, r.Bounces = a.coverage, a.bounces
    r.init_voxel_cone_tracing(sc)              # upload + DrawDepthTexture + DrawVoxelTexture
    frame = np.empty((h, w, 4), dtype=np.uint8)
    r.Render(frame)
    objloader.save_frame_png(frame, a.out)
    c = r.ctx
    print(f"{sc.name}: {sc.n_tris} triangles, {c.occupied_voxels()} occupied voxels of {a.grid}^3, {c.cone_samples()} cone samples -> {a.out}")
    for p in ("depth", "vox_cover", "vox_shade", "resolve", "mip", "visibility", "cone"):
        print(f"  {p:11s} {c.pass_time_us(p):9.1f} us")


if __name__ == "__main__":
    main()
