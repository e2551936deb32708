import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vct_b200
import vct_b200.glmath as gm
from vct_b200 import scenes, uniforms
sc = scenes.atrium()
u = uniforms.scene_uniforms(sc, V=256, width=1920, height=1080, shadow_map_size=4096, coverage="conservative")
c = vct_b200.Context(0); c.set_uniforms(u); c.load_scene(sc)
c.draw_depth(); c.sync()
hosts = [torch.empty((1080, 1920, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
def cam(i):
    view = gm.view_matrix((0, 4, 0), -90.0 + 0.05 * i, 0.0)
    c.set_mat4("ModelViewMatrix", gm.colmajor((view @ gm.scale(0.05)).astype(np.float32)))
    c.set_3f("CameraPosition", (0, 4, 0))
N = 200
for prof in (1, 0):
    c.set_i("Profile", prof)
    for i in range(10): c.frame()
    c.sync()
    t0 = time.perf_counter()
    for i in range(N): c.frame()
    c.sync(); t1 = time.perf_counter()
    print(f"profile={prof} device-only back-to-back: {(t1 - t0) / N * 1e6:.1f} us/frame")
    t0 = time.perf_counter()
    for i in range(N): cam(i); c.frame()
    c.sync(); t1 = time.perf_counter()
    print(f"profile={prof} + camera update: {(t1 - t0) / N * 1e6:.1f} us/frame")
    t0 = time.perf_counter()
    for i in range(N):
        cam(i); c.frame_async(hosts[i & 1])
        if i >= 1: c.frame_wait()
    c.frame_wait(); t1 = time.perf_counter()
    print(f"profile={prof} pipelined e2e: {(t1 - t0) / N * 1e6:.1f} us/frame")
    t0 = time.perf_counter()
    for i in range(N): cam(i); c.frame(hosts[0])
    t1 = time.perf_counter()
    print(f"profile={prof} synchronous e2e: {(t1 - t0) / N * 1e6:.1f} us/frame")
    t0 = time.perf_counter()
    for i in range(N): cam(i)
    t1 = time.perf_counter()
    print(f"   host camera math alone: {(t1 - t0) / N * 1e6:.1f} us")
