"""Characterises the hardware trilinear filter (weight quantisation, rounding, mip fraction) on the GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vct_b200
V = 8
c = vct_b200.Context(0)
c.set_i("VoxelDimensions", V)
g = np.zeros((V, V, V, 4), dtype=np.uint8)
g[:, :, 3] = 0; g[:, :, 4] = 255          # step between x=3 and x=4, constant in y,z
c.upload_grid_level0(g)
vws = 150.0 / V
# sweep x from texel centre 3 to texel centre 4 in 1/4096 texel steps, y/z at a texel centre
N = 4097
t = np.arange(N) / 4096.0
x = ((3 + 0.5 + t) * vws - 75.0).astype(np.float32)
pos = np.stack([x, np.full(N, (2 + 0.5) * vws - 75.0), np.full(N, (5 + 0.5) * vws - 75.0)], 1)
out = c.sample_voxels(pos, 0.0)[:, 0]
vals = np.unique(out)
print("distinct values over one texel:", len(vals))
w = out  # = weight alpha since texels are 0 and 1
q = w * 256
print("max |alpha*256 - round|:", np.abs(q - np.rint(q)).max())
# where does alpha switch from k/256 to (k+1)/256 ?  compare with t
k = np.rint(q).astype(int)
for kk in (0, 1, 2, 127, 128, 255):
    idx = np.nonzero(k == kk)[0]
    if len(idx): print(f"alpha={kk}/256 for t in [{t[idx[0]]:.6f}, {t[idx[-1]]:.6f}]  (t*256 in [{t[idx[0]]*256:.4f}, {t[idx[-1]]*256:.4f}])")
# non-trivial texel values: check lerp rounding: texels 37 and 201
g[:, :, 3] = 37; g[:, :, 4] = 201
c.upload_grid_level0(g)
out = c.sample_voxels(pos, 0.0)[:, 0]
a = k / 256.0
pred = (37 + a * (201 - 37)) / 255.0
print("lerp with quantised alpha: max err vs fp32 formula:", np.abs(out - pred).max(), " (1/255 =", 1 / 255, ")")
pred2 = ((1 - a) * 37 + a * 201) / 255.0
print("   alt formula:", np.abs(out - pred2).max())
# mip fraction: L0 = alternating 0/255 planes in x so that L1 = 128 (rounded 127.5) everywhere
g[:] = 0
g[:, :, 1::2] = 255
c.upload_grid_level0(g)
l1 = c.grid(1)
print("L1 unique:", np.unique(l1))
p0 = np.array([[(2 + 0.5) * vws - 75.0, (2 + 0.5) * vws - 75.0, (5 + 0.5) * vws - 75.0]], dtype=np.float32)  # texel x=2 -> value 0
M = 2049
f = np.arange(M) / 2048.0
out = c.sample_voxels(np.repeat(p0, M, 0), f.astype(np.float32))[:, 0]
fm = out / (l1.flat[0] / 255.0)
print("distinct mip fractions in [0,1]:", len(np.unique(out)))
qf = fm * 256
print("max |f*256 - round|:", np.abs(qf - np.rint(qf)).max())
kf = np.rint(qf).astype(int)
for kk in (0, 1, 2, 128, 255, 256):
    idx = np.nonzero(kf == kk)[0]
    if len(idx): print(f"mipfrac={kk}/256 for lod in [{f[idx[0]]:.6f}, {f[idx[-1]]:.6f}] (lod*256 in [{f[idx[0]]*256:.4f}, {f[idx[-1]]*256:.4f}])")
# 3D: random texels, random positions, compare with emulation alpha = round(frac*256)/256 vs floor
rng = np.random.default_rng(1)
g = rng.integers(0, 256, (V, V, V, 4), dtype=np.uint8)
c.upload_grid_level0(g)
n = 20000
P = rng.uniform(-70, 70, (n, 3)).astype(np.float32)
out = c.sample_voxels(P, 0.0)
def emu(P, mode):
    u = (P / np.float32(75.0)) * np.float32(0.5) + np.float32(0.5)
    xx = u * V - 0.5
    i0 = np.floor(xx).astype(int); fr = xx - i0
    if mode == "round": a = np.floor(fr * 256 + 0.5) / 256
    elif mode == "floor": a = np.floor(fr * 256) / 256
    else: a = fr
    i1 = (i0 + 1) % V; i0 = i0 % V
    G = g.astype(np.float64) / 255
    def at(ix, iy, iz): return G[iz, iy, ix]
    ax, ay, az = a[:, 0:1], a[:, 1:2], a[:, 2:3]
    c00 = at(i0[:,0], i0[:,1], i0[:,2]) * (1-ax) + at(i1[:,0], i0[:,1], i0[:,2]) * ax
    c10 = at(i0[:,0], i1[:,1], i0[:,2]) * (1-ax) + at(i1[:,0], i1[:,1], i0[:,2]) * ax
    c01 = at(i0[:,0], i0[:,1], i1[:,2]) * (1-ax) + at(i1[:,0], i0[:,1], i1[:,2]) * ax
    c11 = at(i0[:,0], i1[:,1], i1[:,2]) * (1-ax) + at(i1[:,0], i1[:,1], i1[:,2]) * ax
    c0 = c00 * (1-ay) + c10 * ay; c1 = c01 * (1-ay) + c11 * ay
    return c0 * (1-az) + c1 * az
for mode in ("exact", "round", "floor"):
    e = np.abs(out - emu(P, mode))
    print(f"3D random: emulation '{mode}': max err {e.max():.6f} mean {e.mean():.7f}")
