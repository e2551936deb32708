import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
from vct_b200 import scenes, uniforms
sc = scenes.atrium()
u = uniforms.scene_uniforms(sc, V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.sync(); print("shadow items", c.debug_counter(0))
c.draw_voxels(); c.sync(); print("vox items", c.debug_counter(0), "frags", c.fragment_count())
c.render(); c.sync(); print("vis items", c.debug_counter(0))
vis = c.visibility()
ids, cnt = np.unique(vis[vis != 0xFFFFFFFF], return_counts=True)
print("visible tris", len(ids), "max px/tri", cnt.max(), "median", np.median(cnt), "tris>10k px", (cnt > 10000).sum(), "px in those", cnt[cnt > 10000].sum())
# bbox tile estimate on host for visibility
