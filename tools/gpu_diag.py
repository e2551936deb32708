"""Diagnostic run on a GPU box: CUDA path vs oracle on small scenes, prints per-stage agreement."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
from oracle.oracle_py import Oracle

OUT = "gpurun_out"
os.makedirs(OUT, exist_ok=True)


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = (d ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def compare(scene, tag, **kw):
    u = uniforms.scene_uniforms(scene, **kw)
    o = Oracle(); o.set_uniforms(u); o.load_scene(scene)
    c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(scene)
    t = time.time(); o.draw_depth(); to = time.time() - t
    c.draw_depth(); c.sync()
    dg, do = c.depth(), o.depth()
    print(f"[{tag}] depth: equal={np.array_equal(dg, do)} mismatches={(dg != do).sum()} maxdiff={np.abs(dg.astype(np.int64) - do.astype(np.int64)).max()} oracle {to:.2f}s gpu {c.pass_time_us('depth'):.1f}us")
    t = time.time(); o.draw_voxels(); to = time.time() - t
    c.draw_voxels(); c.sync()
    cg, co = c.counts(), o.counts()
    print(f"[{tag}] counts: equal={np.array_equal(cg, co)} mismatches={(cg != co).sum()} occupied gpu={int((cg > 0).sum())} oracle={int((co > 0).sum())} frags gpu={int(cg.sum())} oracle={int(co.sum())} touched={c.occupied_voxels()} oracle {to:.2f}s")
    sg, so = c.sums(), o.sums()
    ds = np.abs(sg.astype(np.int64) - so.astype(np.int64))
    print(f"[{tag}] sums: equal={np.array_equal(sg, so)} mismatched cells={(ds.max(-1) > 0).sum()} maxdiff={ds.max()}")
    for l in range(0, o.V.bit_length()):
        gg, go = c.grid(l), o.grid(l)
        d = np.abs(gg.astype(int) - go.astype(int))
        print(f"[{tag}] grid L{l}: equal={np.array_equal(gg, go)} maxdiff={d.max()} n>0={(d > 0).sum()} n>2={(d > 2).sum()}")
    print(f"[{tag}] voxel passes us: clear {c.pass_time_us('vox_clear'):.1f} cover {c.pass_time_us('vox_cover'):.1f} shade {c.pass_time_us('vox_shade'):.1f} resolve {c.pass_time_us('resolve'):.1f} mip {c.pass_time_us('mip'):.1f}")
    t = time.time(); o.render(); to = time.time() - t
    c.render(); c.sync()
    vg, vo = c.visibility(), o.visibility()
    print(f"[{tag}] visibility: equal={np.array_equal(vg, vo)} mismatches={(vg != vo).sum()} of {vg.size}")
    fg, fo = c.read_frame(), o.frame()
    d = np.abs(fg.astype(int) - fo.astype(int)).max(-1)
    print(f"[{tag}] frame: psnr={psnr(fg[..., :3], fo[..., :3]):.2f} dB  maxdiff={d.max()}  frac<=2: {(d <= 2).mean():.5f}  cone samples gpu={c.cone_samples()} oracle={o.cone_samples()}  oracle {to:.2f}s vis {c.pass_time_us('visibility'):.1f}us cone {c.pass_time_us('cone'):.1f}us")
    try:
        from PIL import Image
        Image.fromarray(fg[::-1, :, :3]).save(f"{OUT}/{tag}_gpu.png")
        Image.fromarray(fo[::-1, :, :3]).save(f"{OUT}/{tag}_oracle.png")
        Image.fromarray((np.clip(d * 40, 0, 255)).astype(np.uint8)[::-1]).save(f"{OUT}/{tag}_diff.png")
    except Exception as e:
        print("png failed", e)
    c.close(); o.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["cornell", "atrium_small"]
    if "cornell" in which:
        for cov in ("center", "msaa4", "conservative"):
            compare(scenes.cornell(), f"cornell_{cov}", V=64, width=256, height=256, shadow_map_size=1024, coverage=cov)
    if "atrium_small" in which:
        compare(scenes.atrium(detail=0.3, tex_size=128), "atrium_small", V=128, width=640, height=360, shadow_map_size=2048, coverage="conservative")
    if "atrium" in which:
        compare(scenes.atrium(), "atrium", V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
