// vct_mip.cu -- M1: 2x2x2 box-filter pyramid of the RGBA8 voxel texture.
// Replaces glGenerateMipmap(GL_TEXTURE_3D) (Voxel_Cone_Tracing.h:246-248).  Rounding rule (the GL one is
// driver defined): child = (sum of 8 parents + 4) >> 3 per channel.
//
//   mip_fused3 : reads a 32x8x8 block of level L once (16-byte surface loads, 4 texels per thread along
//                x) and writes levels L+1, L+2, L+3, staging the intermediate levels in shared memory.
//                HBM traffic = read L once + write the three children.
//   mip_tail   : one CTA reduces a level of <= 32^3 texels down to 1^3 entirely in shared memory.
// At V = 256 the pyramid is built by three launches: fused3(0 -> 1,2,3), fused3(3 -> 4,5,6), tail(6 -> 7,8).
#include <cuda_fp16.h>

#include "vct_internal.h"

namespace vct {

// packed byte arithmetic: even/odd bytes of a uchar4 widened into 16-bit lanes
__device__ __forceinline__ void acc_px(uint32_t p, uint32_t& even, uint32_t& odd) {
  even += p & 0x00FF00FFu;
  odd += (p >> 8) & 0x00FF00FFu;
}
__device__ __forceinline__ uint32_t finish_px(uint32_t even, uint32_t odd) {
  even = ((even + 0x00040004u) >> 3) & 0x00FF00FFu;
  odd = ((odd + 0x00040004u) >> 3) & 0x00FF00FFu;
  return even | (odd << 8);
}

// block = (8,4,4) threads, covers 32x8x8 texels of `src`
__global__ void __launch_bounds__(128) mip_fused3(cudaSurfaceObject_t src, cudaSurfaceObject_t d1,
                                                  cudaSurfaceObject_t d2, cudaSurfaceObject_t d3,
                                                  const unsigned char* __restrict__ dirty,
                                                  const unsigned char* __restrict__ dirty_prev, int bricks_x, int n_bricks_x) {
  __shared__ uint32_t s1[4][4][16];
  __shared__ uint32_t s2[2][2][8];
  const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
  // a block walks `bricks_x` consecutive 32x8x8 source bricks along x.  Sparse build (level 0 only): bricks that did not
  // change since the pyramid was last built are skipped -- one block per brick meant tens of thousands of blocks that
  // exit at once, and their launch overhead WAS the kernel time at 512^3.
  for (int kb = 0; kb < bricks_x; ++kb) {
  const int brick_x = blockIdx.x * bricks_x + kb;
  if (brick_x >= n_bricks_x) break;
  if (dirty) {
    const uint32_t brick = (blockIdx.z * gridDim.y + blockIdx.y) * n_bricks_x + brick_x;
    if (!(dirty[brick] | dirty_prev[brick])) continue;
  }
  const int bx = brick_x * 32, by = blockIdx.y * 8, bz = blockIdx.z * 8;
  uint32_t e0 = 0, o0 = 0, e1 = 0, o1 = 0;
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      uint4 q = surf3Dread<uint4>(src, (bx + tx * 4) * 4, by + ty * 2 + dy, bz + tz * 2 + dz);
      acc_px(q.x, e0, o0); acc_px(q.y, e0, o0);
      acc_px(q.z, e1, o1); acc_px(q.w, e1, o1);
    }
  const uint32_t c0 = finish_px(e0, o0), c1 = finish_px(e1, o1);
  const int l1x = (bx >> 1) + tx * 2, l1y = (by >> 1) + ty, l1z = (bz >> 1) + tz;
  surf3Dwrite(make_uint2(c0, c1), d1, l1x * 4, l1y, l1z);
  s1[tz][ty][tx * 2] = c0;
  s1[tz][ty][tx * 2 + 1] = c1;
  __syncthreads();
  const int t = (tz * 4 + ty) * 8 + tx;
  if (t < 32) {   // level +2: 8x2x2
    const int x = t & 7, y = (t >> 3) & 1, z = t >> 4;
    uint32_t e = 0, o = 0;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        acc_px(s1[z * 2 + dz][y * 2 + dy][x * 2], e, o);
        acc_px(s1[z * 2 + dz][y * 2 + dy][x * 2 + 1], e, o);
      }
    const uint32_t c = finish_px(e, o);
    surf3Dwrite(c, d2, ((bx >> 2) + x) * 4, (by >> 2) + y, (bz >> 2) + z);
    s2[z][y][x] = c;
  }
  __syncthreads();
  if (t < 4) {    // level +3: 4x1x1
    uint32_t e = 0, o = 0;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        acc_px(s2[dz][dy][t * 2], e, o);
        acc_px(s2[dz][dy][t * 2 + 1], e, o);
      }
    surf3Dwrite(finish_px(e, o), d3, ((bx >> 3) + t) * 4, by >> 3, bz >> 3);
  }
  __syncthreads();     // the staging buffers are reused by the next brick
  }
}

struct TailSurfaces { cudaSurfaceObject_t s[8]; };   // s[0] = source level, s[1..] = children

// one CTA; n0 = source size (<= 32), n_out = number of child levels to produce (log2(n0))
__global__ void __launch_bounds__(1024) mip_tail(TailSurfaces lv, int n0, int n_out) {
  extern __shared__ uint32_t sm[];
  uint32_t* cur = sm;                       // n0^3
  uint32_t* nxt = sm + n0 * n0 * n0;        // (n0/2)^3, then ping-pong inside the first buffer
  const int tid = threadIdx.x;
  for (int i = tid; i < n0 * n0 * n0; i += blockDim.x) {
    int x = i % n0, y = (i / n0) % n0, z = i / (n0 * n0);
    cur[i] = surf3Dread<uint32_t>(lv.s[0], x * 4, y, z);
  }
  __syncthreads();
  int n = n0;
  for (int l = 1; l <= n_out; ++l) {
    const int h = n >> 1;
    for (int i = tid; i < h * h * h; i += blockDim.x) {
      int x = i % h, y = (i / h) % h, z = i / (h * h);
      uint32_t e = 0, o = 0;
#pragma unroll
      for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const uint32_t* row = cur + ((2 * z + dz) * n + (2 * y + dy)) * n + 2 * x;
          acc_px(row[0], e, o);
          acc_px(row[1], e, o);
        }
      const uint32_t c = finish_px(e, o);
      nxt[i] = c;
      surf3Dwrite(c, lv.s[l], x * 4, y, z);
    }
    __syncthreads();
    uint32_t* t = cur; cur = nxt; nxt = t;   // the old source buffer is large enough for every later level
    n = h;
  }
}

// generic single level (used when a level is too small for fused3 and too big for tail; not hit for
// power-of-two V >= 8, kept for V < 32 odd cases)
__global__ void mip_one(cudaSurfaceObject_t src, cudaSurfaceObject_t dst, int h) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, z = blockIdx.z;
  if (x >= h || y >= h) return;
  uint32_t e = 0, o = 0;
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      uint2 q = surf3Dread<uint2>(src, (2 * x) * 4, 2 * y + dy, 2 * z + dz);
      acc_px(q.x, e, o);
      acc_px(q.y, e, o);
    }
  surf3Dwrite(finish_px(e, o), dst, x * 4, y, z);
}

// RGBA16F level: fp32 sum of the 8 parents in the fixed order ((a00 + a10) + a01) + a11 (a_yz = the x pair at
// row 2y+dy, slice 2z+dz), times 0.125, rounded to half.  One thread per child texel; every level is read once and
// written once (1.29 x level 0 in total), which is what bounds it: HBM.
__global__ void mip_level_f16(cudaSurfaceObject_t src, cudaSurfaceObject_t dst, int h, const unsigned char* __restrict__ dirty,
                              const unsigned char* __restrict__ dirty_prev, int shift, int V) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, z = blockIdx.z;
  if (x >= h || y >= h) return;
  // sparse build (child levels 1..3: a child texel lies inside one 32x8x8 level-0 brick)
  if (dirty) {
    const uint32_t brick = brick_of(x << shift, y << shift, z << shift, V);
    if (!(dirty[brick] | dirty_prev[brick])) return;
  }
  float acc[4];
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const uint4 q = surf3Dread<uint4>(src, (2 * x) * 8, 2 * y + dy, 2 * z + dz);   // two texels
      const unsigned w0[4] = {q.x & 0xFFFFu, q.x >> 16, q.y & 0xFFFFu, q.y >> 16};
      const unsigned w1[4] = {q.z & 0xFFFFu, q.z >> 16, q.w & 0xFFFFu, q.w >> 16};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const float a = __half2float(__ushort_as_half((unsigned short)w0[ch])) + __half2float(__ushort_as_half((unsigned short)w1[ch]));
        acc[ch] = (dz == 0 && dy == 0) ? a : acc[ch] + a;
      }
    }
  unsigned short o[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) o[ch] = __half_as_ushort(__float2half_rn(acc[ch] * 0.125f));
  surf3Dwrite(make_uint2((unsigned)o[0] | ((unsigned)o[1] << 16), (unsigned)o[2] | ((unsigned)o[3] << 16)), dst, x * 8, y, z);
}

// RGBA16F, three levels per launch: a block of (16,4,4) threads reads a 32x8x8 brick of `src` once (16-byte surface
// loads: two texels) and writes its 16x4x4, 8x2x2 and 4x1x1 children, staging the intermediate levels in shared
// memory AS HALVES -- every level is computed from the rounded texels of the level above, in the order
// ((a00 + a10) + a01) + a11 of mip_level_f16, so the result is bit-identical to the level-by-level build.
__device__ __forceinline__ void f16_acc(const uint2 t, float acc[4], bool first) {
  const unsigned w[4] = {t.x & 0xFFFFu, t.x >> 16, t.y & 0xFFFFu, t.y >> 16};
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const float a = __half2float(__ushort_as_half((unsigned short)w[ch]));
    acc[ch] = first ? a : acc[ch] + a;
  }
}
__device__ __forceinline__ uint2 f16_pack(const float acc[4]) {
  unsigned short o[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) o[ch] = __half_as_ushort(__float2half_rn(acc[ch] * 0.125f));
  return make_uint2((unsigned)o[0] | ((unsigned)o[1] << 16), (unsigned)o[2] | ((unsigned)o[3] << 16));
}
// the 2x2x2 parents of child (x, y, z) in a staged level: x pair first, then rows, then slices
template <int NX, int NY>
__device__ __forceinline__ uint2 f16_reduce_staged(const uint2* s, int x, int y, int z) {
  float acc[4], pair[4];
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const uint2* row = s + ((2 * z + dz) * NY + (2 * y + dy)) * NX + 2 * x;
      f16_acc(row[0], pair, true);
      f16_acc(row[1], pair, false);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) acc[ch] = (dz == 0 && dy == 0) ? pair[ch] : acc[ch] + pair[ch];
    }
  return f16_pack(acc);
}

__global__ void __launch_bounds__(256) mip_fused3_f16(cudaSurfaceObject_t src, cudaSurfaceObject_t d1, cudaSurfaceObject_t d2,
                                                      cudaSurfaceObject_t d3, const unsigned char* __restrict__ dirty,
                                                      const unsigned char* __restrict__ dirty_prev, int bricks_x, int n_bricks_x) {
  __shared__ uint2 s1[4 * 4 * 16];
  __shared__ uint2 s2[2 * 2 * 8];
  const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
  for (int kb = 0; kb < bricks_x; ++kb) {      // see mip_fused3: a block walks several bricks along x
  const int brick_x = blockIdx.x * bricks_x + kb;
  if (brick_x >= n_bricks_x) break;
  if (dirty) {    // sparse build (level 0 only): this brick did not change since the pyramid was last built
    const uint32_t brick = (blockIdx.z * gridDim.y + blockIdx.y) * n_bricks_x + brick_x;
    if (!(dirty[brick] | dirty_prev[brick])) continue;
  }
  const int bx = brick_x * 32, by = blockIdx.y * 8, bz = blockIdx.z * 8;
  float acc[4], pair[4];
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const uint4 q = surf3Dread<uint4>(src, (bx + tx * 2) * 8, by + ty * 2 + dy, bz + tz * 2 + dz);   // two texels
      f16_acc(make_uint2(q.x, q.y), pair, true);
      f16_acc(make_uint2(q.z, q.w), pair, false);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) acc[ch] = (dz == 0 && dy == 0) ? pair[ch] : acc[ch] + pair[ch];
    }
  const uint2 c1 = f16_pack(acc);
  surf3Dwrite(c1, d1, ((bx >> 1) + tx) * 8, (by >> 1) + ty, (bz >> 1) + tz);
  s1[(tz * 4 + ty) * 16 + tx] = c1;
  __syncthreads();
  const int t = (tz * 4 + ty) * 16 + tx;
  if (t < 32) {   // level +2: 8x2x2
    const int x = t & 7, y = (t >> 3) & 1, z = t >> 4;
    const uint2 c2 = f16_reduce_staged<16, 4>(s1, x, y, z);
    surf3Dwrite(c2, d2, ((bx >> 2) + x) * 8, (by >> 2) + y, (bz >> 2) + z);
    s2[(z * 2 + y) * 8 + x] = c2;
  }
  __syncthreads();
  if (t < 4)      // level +3: 4x1x1
    surf3Dwrite(f16_reduce_staged<8, 2>(s2, t, 0, 0), d3, ((bx >> 3) + t) * 8, by >> 3, bz >> 3);
  __syncthreads();     // the staging buffers are reused by the next brick
  }
}

// Sparse build: bricks per block such that the launch has ~8 K blocks.  With one brick per block a 512^3 level has
// 65 536 blocks, most of which exit at once -- their launch overhead was the kernel time (82 -> 67 us); at 256^3 there
// are 8 192 bricks and walking several per block only serialises the dirty ones (22 -> 32 us), so it stays at one.
static int sparse_bricks_per_block(bool sparse, int n) {
  if (!sparse) return 1;
  const long long bricks = (long long)(n / 32) * (n / 8) * (n / 8);
  int per = 1;
  while (per < 8 && bricks / per > 8192 && (n / 32) % (per * 2) == 0) per *= 2;
  return per;
}

int launch_mip(vct_context* c) {
  int rc = ensure_grid(c); if (rc) return rc;
  PassTimer timer(c, VCT_PASS_MIP);
  const int levels = c->P.levels;
  vct_context::GridBuf& gb = c->grid[c->cur];
  const unsigned char* dirty = (gb.dirty_valid && c->P.V >= 32 && !c->dense_resolve) ? gb.dirty_now : nullptr;
  gb.mips_current = true;
  if (c->grid_format == 1) {
    int l = 0, n = c->P.V;
    while (n >= 32 && l + 3 < levels) {      // three levels per launch while the source level has whole 32x8x8 bricks
      const int nbx = n / 32, per = sparse_bricks_per_block(l == 0 && dirty, n);
      dim3 b(16, 4, 4), g((nbx + per - 1) / per, n / 8, n / 8);
      mip_fused3_f16<<<g, b, 0, c->stream>>>(gb.surf[l], gb.surf[l + 1], gb.surf[l + 2], gb.surf[l + 3], l == 0 ? dirty : nullptr, gb.dirty_prev, per, nbx);
      c->launches += 1;
      l += 3; n >>= 3;
    }
    for (; l + 1 < levels; ++l, n >>= 1) {
      const int h = n >> 1;
      dim3 b(32, 4), g((h + 31) / 32, (h + 3) / 4, h);
      mip_level_f16<<<g, b, 0, c->stream>>>(gb.surf[l], gb.surf[l + 1], h, l + 1 <= 3 ? dirty : nullptr, gb.dirty_prev, l + 1, c->P.V);
      c->launches += 1;
    }
    VCT_CUDA(c, cudaGetLastError());
    return VCT_OK;
  }
  int l = 0, n = c->P.V;
  while (n >= 32 && l + 3 < levels) {
    const int nbx = n / 32, per = sparse_bricks_per_block(l == 0 && dirty, n);
    dim3 b(8, 4, 4), g((nbx + per - 1) / per, n / 8, n / 8);
    mip_fused3<<<g, b, 0, c->stream>>>(gb.surf[l], gb.surf[l + 1], gb.surf[l + 2], gb.surf[l + 3], l == 0 ? dirty : nullptr, gb.dirty_prev, per, nbx);
    c->launches += 1;
    l += 3; n >>= 3;
  }
  while (n > 32) {   // V not reducible by fused3 steps down to <= 32 (e.g. 64 -> 8 is fine; defensive)
    dim3 b(32, 8), g((n / 2 + 31) / 32, (n / 2 + 7) / 8, n / 2);
    mip_one<<<g, b, 0, c->stream>>>(c->grid[c->cur].surf[l], c->grid[c->cur].surf[l + 1], n / 2);
    c->launches += 1;
    l += 1; n >>= 1;
  }
  if (l < levels - 1) {
    TailSurfaces ts{};
    int n_out = levels - 1 - l;
    for (int k = 0; k <= n_out; ++k) ts.s[k] = c->grid[c->cur].surf[l + k];
    size_t smem = ((size_t)n * n * n + (size_t)(n / 2) * (n / 2) * (n / 2)) * 4;
    if (smem > 48 * 1024)
      VCT_CUDA(c, cudaFuncSetAttribute(mip_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    mip_tail<<<1, 1024, smem, c->stream>>>(ts, n, n_out);
    c->launches += 1;
  }
  VCT_CUDA(c, cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct
