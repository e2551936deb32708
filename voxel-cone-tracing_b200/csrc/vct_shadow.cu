// vct_shadow.cu -- S1: light-space depth map.
// Replaces DrawDepthTexture (Voxel_Cone_Tracing.h:192-211) + Shader/Shadow.vs:7-10 + the GL depth
// rasteriser: back faces culled, depth test LESS, cleared to 1.0, GL_DEPTH_COMPONENT24.
// Triangle-parallel exact integer rasterisation (vct_raster.cuh) with atomicMin on the 24-bit depth.
#include "vct_raster.cuh"

namespace vct {

struct ShadowPass {
  Params P;                  // uniform block, by value: lives in the kernel's constant bank
  const float* verts; const uint32_t* idx;
  uint32_t* depth;
  // sharded shadow map (triangle ranges + min-reduction of the depth image): mode 1 = every depth fragment is
  // min-reduced into ALL ranks' images by one multimem.red.min.u32 (performed in the NVSwitch), mode 2 = one atomicMin
  // per rank through the peer mappings; mode 0 = the private image
  int mode; int world; size_t seg_words;

  static constexpr bool kAppends = false;
  static constexpr bool kWarpMedium = true;
  struct Setup { RasterTri t; float z0, z1, z2; };

  __device__ __forceinline__ bool setup(uint32_t tri, Setup& s, int& i0, int& i1, int& j0, int& j1) const {
    const int S = P.S;
    float wx[3], wy[3], wz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* v = verts + (size_t)__ldg(&idx[tri * 3 + k]) * 14;
      F4 c = mul_mat_vec(P.depth_mvp, __ldg(v), __ldg(v + 1), __ldg(v + 2), 1.0f);
      if (!(c.w > 0.0f)) return false;
      float nx = c.x / c.w, ny = c.y / c.w, nz = c.z / c.w;
      wx[k] = (nx * 0.5f + 0.5f) * (float)S;
      wy[k] = (ny * 0.5f + 0.5f) * (float)S;
      wz[k] = nz * 0.5f + 0.5f;
    }
    if (!setup_raster(wx, wy, &s.t)) return false;
    if (s.t.flipped) return false;   // GL_CULL_FACE / GL_BACK, Voxel_Cone_Tracing.h:194
    s.z0 = wz[0]; s.z1 = wz[1]; s.z2 = wz[2];
    return raster_bbox(s.t, 0, S, S, &i0, &i1, &j0, &j1);
  }

  __device__ __forceinline__ void emit(const Setup& s, int i, int j) const {
    float l1, l2;
    s.t.lambdas(i, j, &l1, &l2);
    float z = interp3(s.z0, s.z1, s.z2, l1, l2);
    if (!(z >= 0.0f) || z > 1.0f) return;   // near/far clip
    uint32_t d = (uint32_t)__float2int_rn(z * 16777215.0f);
    const size_t texel = (size_t)j * P.S + i;
    if (mode == 1) asm volatile("multimem.red.relaxed.sys.global.min.u32 [%0], %1;" ::"l"(depth + texel), "r"(d) : "memory");
    else if (mode == 2) { for (int p = 0; p < world; ++p) atomicMin(depth + (size_t)p * seg_words + texel, d); }
    else atomicMin(&depth[texel], d);
  }

  __device__ __forceinline__ bool setup_full(uint32_t tri, Setup& s, int& i0, int& i1, int& j0, int& j1) const {
    return setup(tri, s, i0, i1, j0, j1);
  }
  __device__ __forceinline__ bool tile_may_cover(const Setup& s, int x0, int y0, int x1, int y1) const {
    return tile_may_cover_exact(s.t, x0, y0, x1, y1);
  }
  __device__ __forceinline__ void tile_rows(int&, int&, int&) const {}      // every tile row of the box

  __device__ __forceinline__ void small(const Setup& s, uint32_t, bool active, int i0, int i1, int j0, int j1) const {
    if (!active) return;
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i)
        if (s.t.covered(i, j, 0)) emit(s, i, j);
  }

  __device__ __forceinline__ void pixel(const Setup& s, uint32_t, int i, int j, bool in_bbox) const {
    if (in_bbox && s.t.covered(i, j, 0)) emit(s, i, j);
  }
};

__global__ void fill_u32(uint32_t* p, size_t n, uint32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// The D24 map as a 2D float array for tex2Dgather (voxel shading): four texels per thread, 16-byte surface stores.
// (cudaMemcpy2DToArrayAsync of the 64 MiB map took ~270 us per call -- more than rasterising it.)
__global__ void depth_to_array(const uint32_t* __restrict__ depth, cudaSurfaceObject_t surf, int S) {
  const int x4 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x4 * 4 >= S || y >= S) return;
  // the array holds the texels as the floats every PCF tap would otherwise derive from the D24 integers
  // (`float(d24) * (1/16777215)`, the same expression, evaluated once per texel instead of 36 times per fragment)
  if (x4 * 4 + 3 < S && (S & 3) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(depth + (size_t)y * S + x4 * 4);
    const float4 f = make_float4((float)v.x * (1.0f / 16777215.0f), (float)v.y * (1.0f / 16777215.0f),
                                 (float)v.z * (1.0f / 16777215.0f), (float)v.w * (1.0f / 16777215.0f));
    surf2Dwrite(f, surf, x4 * 16, y);
  } else {
    for (int x = x4 * 4; x < min(x4 * 4 + 4, S); ++x)
      surf2Dwrite((float)depth[(size_t)y * S + x] * (1.0f / 16777215.0f), surf, x * 4, y);
  }
}

int launch_shadow(vct_context* c) {
  if (!c->nt) return set_error(c, VCT_ERR_STATE, "vct_draw_depth: no mesh uploaded");
  int rc = ensure_shadow(c); if (rc) return rc;
  rc = ensure_queues(c); if (rc) return rc;
  PassTimer timer(c, VCT_PASS_DEPTH);
  const size_t n = (size_t)c->P.S * c->P.S;
  uint32_t *sym_local = nullptr, *sym_mc = nullptr, *sym_peers = nullptr; size_t seg_words = 0; int world = 1, rank = 0;
  const bool sharded = c->shard_shadow && c->shared_world > 1 && comm_depth_views(c, &sym_local, &sym_mc, &sym_peers, &seg_words, &world, &rank);
  if (c->shard_shadow && c->shared_world > 1 && !sharded)
    return set_error(c, VCT_ERR_STATE, "ShardShadowMap: set it (and ShadowMapSize) before vct_comm_init");
  VCT_CUDA(c, reset_item_queue(c));
  const uint32_t nt = (uint32_t)c->nt;
  if (!sharded) {
    fill_u32<<<148 * 8, 256, 0, c->stream>>>(c->d_depth, n, 0xFFFFFFu);
    ShadowPass pass{c->P, c->d_verts, c->d_idx, c->d_depth, 0, 1, 0};
    raster_small<ShadowPass><<<(nt + 127) / 128, 128, 0, c->stream>>>(pass, 0, nt, c->d_items, (uint32_t)c->items_cap, c->d_counters);
    raster_tiles<ShadowPass><<<148 * 4, 256, 0, c->stream>>>(pass, c->d_items, (uint32_t)c->items_cap, c->d_counters);
    c->launches += 3;
  } else {
    // every rank clears ITS image, all wait, every rank rasterises its share of the triangles (blocks of 128 dealt
    // round-robin) into ALL images, all wait, every rank takes a private copy for the passes that sample the map
    fill_u32<<<148 * 8, 256, 0, c->stream>>>(sym_local, n, 0xFFFFFFu);
    int rc2 = comm_barrier(c, 3, c->stream); if (rc2) return rc2;
    ShadowPass pass{c->P, c->d_verts, c->d_idx, sym_mc ? sym_mc : sym_peers, sym_mc ? 1 : 2, world, seg_words};
    const uint32_t n_blocks = (nt + 127) / 128, own = n_blocks > (uint32_t)rank ? (n_blocks - rank + world - 1) / world : 0;
    if (own)
      raster_small<ShadowPass><<<own, 128, 0, c->stream>>>(pass, 0, nt, c->d_items, (uint32_t)c->items_cap, c->d_counters, (uint32_t)world, (uint32_t)rank);
    raster_tiles<ShadowPass><<<148 * 4, 256, 0, c->stream>>>(pass, c->d_items, (uint32_t)c->items_cap, c->d_counters);
    rc2 = comm_barrier(c, 3, c->stream); if (rc2) return rc2;
    VCT_CUDA(c, cudaMemcpyAsync(c->d_depth, sym_local, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    c->launches += 3;
  }
  VCT_CUDA(c, cudaGetLastError());
  // array copy for tex2Dgather (voxel shading)
  if (c->depth_surf) {
    dim3 b(32, 8), g((c->P.S / 4 + 32) / 32, (c->P.S + 7) / 8);
    depth_to_array<<<g, b, 0, c->stream>>>(c->d_depth, c->depth_surf, c->P.S);
    c->launches += 1;
    VCT_CUDA(c, cudaGetLastError());
  } else {
    return set_error(c, VCT_ERR_CUDA, "the shadow-map gather array has no surface object");
  }
  c->depth_valid = true;
  return VCT_OK;
}

}  // namespace vct
