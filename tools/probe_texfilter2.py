import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
V = 8
c = vct_b200.Context(0)
c.set_i("VoxelDimensions", V)
rng = np.random.default_rng(1)
g = rng.integers(0, 256, (V, V, V, 4), dtype=np.uint8)
c.upload_grid_level0(g)
vws = 150.0 / V
G = g.astype(np.float64) / 255

def emu(P, mode, bits=8):
    u = (P / np.float32(75.0)) * np.float32(0.5) + np.float32(0.5)
    xx = u.astype(np.float64) * V - 0.5
    i0 = np.floor(xx).astype(int); fr = xx - i0
    s = float(1 << bits)
    if mode == "round": a = np.floor(fr * s + 0.5) / s
    elif mode == "floor": a = np.floor(fr * s) / s
    else: a = fr
    i1 = (i0 + 1) % V; i0 = i0 % V
    def at(ix, iy, iz): return G[iz, iy, ix]
    ax, ay, az = a[:, 0:1], a[:, 1:2], a[:, 2:3]
    c00 = at(i0[:,0], i0[:,1], i0[:,2]) * (1-ax) + at(i1[:,0], i0[:,1], i0[:,2]) * ax
    c10 = at(i0[:,0], i1[:,1], i0[:,2]) * (1-ax) + at(i1[:,0], i1[:,1], i0[:,2]) * ax
    c01 = at(i0[:,0], i0[:,1], i1[:,2]) * (1-ax) + at(i1[:,0], i0[:,1], i1[:,2]) * ax
    c11 = at(i0[:,0], i1[:,1], i1[:,2]) * (1-ax) + at(i1[:,0], i1[:,1], i1[:,2]) * ax
    c0 = c00 * (1-ay) + c10 * ay; c1 = c01 * (1-ay) + c11 * ay
    return c0 * (1-az) + c1 * az

n = 20000
centre = lambda k: (k + 0.5) * vws - 75.0
for axes in ([0], [1], [2], [0, 1], [0, 2], [1, 2], [0, 1, 2]):
    P = np.stack([centre(rng.integers(0, V, n)) for _ in range(3)], 1)
    for a in axes:
        P[:, a] = rng.uniform(-74, 74, n)
    P = P.astype(np.float32)
    out = c.sample_voxels(P, 0.0).astype(np.float64)
    msg = f"axes {axes}:"
    for mode, bits in (("exact", 8), ("round", 8), ("floor", 8), ("round", 9), ("round", 7), ("round", 6)):
        e = np.abs(out - emu(P, mode, bits))
        msg += f"  {mode}{bits}: max {e.max():.5f} mean {e.mean():.6f}"
    print(msg)
# inspect a few 3-axis samples
P = rng.uniform(-74, 74, (5, 3)).astype(np.float32)
out = c.sample_voxels(P, 0.0)
print(np.c_[P, out[:, 0], emu(P, "round")[:, 0], emu(P, "exact")[:, 0]])
