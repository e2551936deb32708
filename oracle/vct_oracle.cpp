/*
 * vct_oracle.cpp -- CPU ORACLE: scalar restatement of the reference's three passes.
 *
 * TEST INFRASTRUCTURE ONLY (see vct_oracle.h).  PARITY: the programmable stages are pinned against
 * vectors produced by executing the reference's own shader files (tests/glsl_run.py,
 * tests/golden/reference_shader_vectors.npz, tests/test_reference_glsl.py: byte-exact frames); the
 * fixed-function stages between them are UNPINNED by the reference (it contains no code for them,
 * has no tests, cannot be built here and reads nothing back) and follow the GL 4.3 specification.
 * Further pins: the hand-derived known-answer tests (tests/test_oracle_kat.py), the committed golden
 * fixtures (tests/golden/) and a second, independent restatement of the same shader lines in float64
 * numpy / exact integers (tests/test_oracle_independent.py).
 *
 * Each function cites the reference file:line it restates; paths are relative to
 * /root/reference/Voxel_Cone_Tracing_Final/.  Where the reference leans on GL fixed function
 * (rasterisation, depth quantisation, texture filtering, glGenerateMipmap) the rule implemented is
 * the one written in DESIGN.md "Defined semantics", which both this file and the CUDA kernels follow
 * independently.  All float arithmetic that feeds an integer decision (coverage, depth slice,
 * shadow compare) is written as single IEEE-754 binary32 operations in a fixed order and this
 * file must be compiled with -ffp-contract=off so that nothing is fused.
 */
#include "vct_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V4 { float x, y, z, w; };
struct V3 { float x, y, z; };

/* mat4 (column-major, glm layout, Shader.h:414-417 uploads with transpose=GL_FALSE) times vec4.
 * Order of operations is part of the defined semantics: ((m0*x + m1*y) + m2*z) + m3*w. */
inline V4 mul_mat_vec(const float* m, float x, float y, float z, float w) {
  V4 r;
  r.x = ((m[0] * x + m[4] * y) + m[8] * z) + m[12] * w;
  r.y = ((m[1] * x + m[5] * y) + m[9] * z) + m[13] * w;
  r.z = ((m[2] * x + m[6] * y) + m[10] * z) + m[14] * w;
  r.w = ((m[3] * x + m[7] * y) + m[11] * z) + m[15] * w;
  return r;
}

inline V3 sub3(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add3(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 scale3(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross3(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3 normalize3(V3 a) {
  float l = std::sqrt(dot3(a, a));
  return {a.x / l, a.y / l, a.z / l};
}

struct Texture {
  int levels = 0;
  bool has_alpha = false;             /* any level-0 texel with alpha < 255 */
  std::vector<int> w, h;
  std::vector<std::vector<uint8_t>> px; /* RGBA8 per level */
};

struct Material { int diffuse = -1, specular = -1, height = -1; float shininess = 20.0f; };

template <class T>
inline void atomic_min(T* cell, T v) {
  T cur = __atomic_load_n(cell, __ATOMIC_RELAXED);
  while (v < cur && !__atomic_compare_exchange_n(cell, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}
inline void atomic_min_u32(uint32_t* c, uint32_t v) { atomic_min(c, v); }
inline void atomic_min_u64(uint64_t* c, uint64_t v) { atomic_min(c, v); }

/* IEEE binary16 <-> binary32, round to nearest even (what cvt.rn.f16.f32 / __float2half_rn do) */
inline uint16_t float_to_half(float f) {
  uint32_t x; std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t mant = x & 0x007FFFFFu;
  int32_t exp = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
  if (((x >> 23) & 0xFF) == 0xFF) return (uint16_t)(sign | 0x7C00u | (mant ? 0x200u : 0));
  if (exp >= 31) return (uint16_t)(sign | 0x7C00u);
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    mant |= 0x00800000u;
    uint32_t shift = (uint32_t)(14 - exp);
    uint32_t h = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)exp << 10) | (mant >> 13);
  uint32_t rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
  return (uint16_t)(sign | h);
}
inline float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1F, mant = h & 0x3FFu, x;
  if (exp == 0) {
    if (!mant) x = sign;
    else { int e = -1; do { ++e; mant <<= 1; } while (!(mant & 0x400u)); x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((mant & 0x3FFu) << 13); }
  } else if (exp == 31) x = sign | 0x7F800000u | (mant << 13);
  else x = sign | ((exp - 15 + 127) << 23) | (mant << 13);
  float f; std::memcpy(&f, &x, 4); return f;
}

constexpr int SUBPIX = 256;                 /* 8 sub-pixel bits */
constexpr float SNAP_LIMIT = 8388608.0f;    /* guard band: +-2^23 sub-pixel units */

}  // namespace

struct orc_ctx {
  orc_params p;
  std::vector<float> verts;        /* nv*14 */
  std::vector<uint32_t> idx;       /* nt*3 */
  std::vector<uint16_t> tri_mat;   /* nt */
  std::vector<Texture> textures;
  std::vector<Material> materials;

  std::vector<uint32_t> depth;     /* S*S d24 */
  std::vector<uint32_t> counts;    /* V^3 */
  std::vector<uint32_t> sums;      /* V^3*3 */
  std::vector<std::vector<uint8_t>> grid; /* levels, RGBA8 */
  std::vector<uint64_t> vis;       /* H*W: depth bits<<32 | tri */
  std::vector<uint8_t> frame;      /* H*W*4 */
  uint64_t cone_samples = 0;
  uint64_t fragments = 0;
  int gridV = 0;
  int gridFmt = -1;
  int bpt() const { return p.GridFormat == 1 ? 8 : 4; }   /* bytes per texel */
  /* channel ch of texel i of level l as the float the sampler sees */
  float texel(int l, size_t i, int ch) const {
    if (p.GridFormat == 1) { uint16_t h; std::memcpy(&h, &grid[l][i * 8 + ch * 2], 2); return half_to_float(h); }
    return grid[l][i * 4 + ch] * (1.0f / 255.0f);
  }
  void set_texel16(int l, size_t i, int ch, float v) { uint16_t h = float_to_half(v); std::memcpy(&grid[l][i * 8 + ch * 2], &h, 2); }
};

namespace {

/* ------------------------------------------------------------------ textures */

/* glGenerateMipmap(GL_TEXTURE_2D), Model.h:170: 2x2 box, round half up; odd sizes clamp the 2nd tap. */
void build_texture_mips(Texture& t) {
  while (t.w.back() > 1 || t.h.back() > 1) {
    int pw = t.w.back(), ph = t.h.back();
    int nw = std::max(1, pw / 2), nh = std::max(1, ph / 2);
    const std::vector<uint8_t>& src = t.px.back();
    std::vector<uint8_t> dst((size_t)nw * nh * 4);
    for (int y = 0; y < nh; ++y)
      for (int x = 0; x < nw; ++x) {
        int x0 = std::min(2 * x, pw - 1), x1 = std::min(2 * x + 1, pw - 1);
        int y0 = std::min(2 * y, ph - 1), y1 = std::min(2 * y + 1, ph - 1);
        for (int c = 0; c < 4; ++c) {
          int s = src[((size_t)y0 * pw + x0) * 4 + c] + src[((size_t)y0 * pw + x1) * 4 + c] +
                  src[((size_t)y1 * pw + x0) * 4 + c] + src[((size_t)y1 * pw + x1) * 4 + c];
          dst[((size_t)y * nw + x) * 4 + c] = (uint8_t)((s + 2) >> 2);
        }
      }
    t.w.push_back(nw);
    t.h.push_back(nh);
    t.px.push_back(std::move(dst));
  }
  t.levels = (int)t.w.size();
}

inline int wrap_repeat(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}

/* One bilinear tap of a mip level, GL_REPEAT (Model.h:172-173), GL_LINEAR. */
void sample_tex_level(const Texture& t, int l, float u, float v, float out[4]) {
  int W = t.w[l], H = t.h[l];
  float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
  float fx = std::floor(x), fy = std::floor(y);
  float a = x - fx, b = y - fy;
  int i0 = wrap_repeat((int)fx, W), i1 = wrap_repeat((int)fx + 1, W);
  int j0 = wrap_repeat((int)fy, H), j1 = wrap_repeat((int)fy + 1, H);
  const uint8_t* p = t.px[l].data();
  for (int c = 0; c < 4; ++c) {
    float t00 = p[((size_t)j0 * W + i0) * 4 + c] * (1.0f / 255.0f);
    float t10 = p[((size_t)j0 * W + i1) * 4 + c] * (1.0f / 255.0f);
    float t01 = p[((size_t)j1 * W + i0) * 4 + c] * (1.0f / 255.0f);
    float t11 = p[((size_t)j1 * W + i1) * 4 + c] * (1.0f / 255.0f);
    float top = t00 + a * (t10 - t00);
    float bot = t01 + a * (t11 - t01);
    out[c] = top + b * (bot - top);
  }
}

/* texture()/textureLod() on a 2D material texture: trilinear (Model.h:174), lod clamped to the chain. */
void sample_tex(const Texture& t, float u, float v, float lod, float out[4]) {
  if (!(u == u) || !(v == v) || std::isinf(u) || std::isinf(v)) { out[0] = out[1] = out[2] = 0; out[3] = 1; return; }
  float maxl = (float)(t.levels - 1);
  if (!(lod > 0.0f)) lod = 0.0f;
  if (lod > maxl) lod = maxl;
  int l0 = (int)std::floor(lod);
  float f = lod - (float)l0;
  float a[4];
  sample_tex_level(t, l0, u, v, a);
  if (f > 0.0f && l0 + 1 < t.levels) {
    float b[4];
    sample_tex_level(t, l0 + 1, u, v, b);
    for (int c = 0; c < 4; ++c) out[c] = a[c] + f * (b[c] - a[c]);
  } else {
    for (int c = 0; c < 4; ++c) out[c] = a[c];
  }
}

/* GL 4.3 section 8.14 scale factor: rho = max(|d(uv*size)/dx|, |d(uv*size)/dy|), lambda = log2(rho). */
inline float lod_from_derivs(float dudx, float dvdx, float dudy, float dvdy, int w, int h) {
  float ax = dudx * (float)w, bx = dvdx * (float)h;
  float ay = dudy * (float)w, by = dvdy * (float)h;
  float rx = std::sqrt(ax * ax + bx * bx);
  float ry = std::sqrt(ay * ay + by * by);
  float rho = std::max(rx, ry);
  if (!(rho > 0.0f)) return 0.0f;
  return std::log2(rho);
}

/* ------------------------------------------------------------------ shadow map sampling */

/* Depth texel as sampled: GL_DEPTH_COMPONENT24 read back as d24/(2^24-1); defined as float(d24)*k. */
inline float depth_texel(const orc_ctx* o, int i, int j) {
  const int S = o->p.ShadowMapSize;
  i = std::min(std::max(i, 0), S - 1);   /* GL_CLAMP_TO_EDGE, Voxel_Cone_Tracing.h:95-96 */
  j = std::min(std::max(j, 0), S - 1);
  return (float)o->depth[(size_t)j * S + i] * (1.0f / 16777215.0f);
}

/* texture(ShadowMap, uv).r with GL_LINEAR and no compare mode (Voxel_Cone_Tracing.h:93-94). */
inline float shadow_bilinear(const orc_ctx* o, float u, float v) {
  const float S = (float)o->p.ShadowMapSize;
  float x = u * S - 0.5f, y = v * S - 0.5f;
  float fx = std::floor(x), fy = std::floor(y);
  float a = x - fx, b = y - fy;
  /* keep the int conversion defined for wild coordinates */
  fx = std::min(std::max(fx, -2.0f), S + 1.0f);
  fy = std::min(std::max(fy, -2.0f), S + 1.0f);
  int i = (int)fx, j = (int)fy;
  float t00 = depth_texel(o, i, j), t10 = depth_texel(o, i + 1, j);
  float t01 = depth_texel(o, i, j + 1), t11 = depth_texel(o, i + 1, j + 1);
  float top = t00 + a * (t10 - t00);
  float bot = t01 + a * (t11 - t01);
  return top + b * (bot - top);
}

/* PCF_Shadow_Mapping, Voxelization.fs:18-52 and VoxelConeTracing.fs:132-163: returns the number of
 * lit taps; the caller applies /25 (voxelization) or *0.111 (cone trace). */
float pcf_lit_taps(const orc_ctx* o, float dcx, float dcy, float dcz, float dcw) {
  if (o->depth.empty()) return 0.0f;
  const int r = o->p.PcfRadius;
  float cur = dcz / dcw;
  float inv = 1.0f / (float)o->p.ShadowMapSize;
  float lit = 0.0f;
  for (int x = -r; x <= r; ++x)
    for (int y = -r; y <= r; ++y) {
      float ox = inv * (float)x, oy = inv * (float)y;
      float closest = shadow_bilinear(o, dcx + ox, dcy + oy);
      if (cur - o->p.ShadowBias <= closest) lit += 1.0f;
    }
  return lit;
}

/* ------------------------------------------------------------------ exact 2D rasteriser (ortho passes) */

inline bool snap(float v, int64_t* out) {
  float s = v * (float)SUBPIX;
  if (!(s == s)) return false;
  s = std::min(std::max(s, -SNAP_LIMIT), SNAP_LIMIT);
  *out = (int64_t)std::lrintf(s); /* round to nearest even */
  return true;
}

struct Edge {
  int64_t ax, ay, dx, dy;
  int64_t bias;  /* 0 for owner (left/top) edges, -1 otherwise */
  inline int64_t eval(int64_t px, int64_t py) const { return dx * (py - ay) - dy * (px - ax); }
};

inline Edge make_edge(int64_t ax, int64_t ay, int64_t bx, int64_t by) {
  Edge e;
  e.ax = ax; e.ay = ay; e.dx = bx - ax; e.dy = by - ay;
  bool owner = (e.dy < 0) || (e.dy == 0 && e.dx < 0);
  e.bias = owner ? 0 : -1;
  return e;
}

struct RasterTri {
  int64_t X[3], Y[3];
  int64_t area;     /* > 0 after orientation fix */
  Edge e01, e12, e20;
  bool flipped;     /* v1 and v2 were swapped to make the area positive */
};

/* returns false if nothing can be rasterised (degenerate / non finite) */
bool setup_raster(const float wx[3], const float wy[3], RasterTri* t) {
  for (int k = 0; k < 3; ++k)
    if (!snap(wx[k], &t->X[k]) || !snap(wy[k], &t->Y[k])) return false;
  int64_t area = (t->X[1] - t->X[0]) * (t->Y[2] - t->Y[0]) - (t->Y[1] - t->Y[0]) * (t->X[2] - t->X[0]);
  if (area == 0) return false;
  t->flipped = area < 0;
  if (t->flipped) {
    std::swap(t->X[1], t->X[2]);
    std::swap(t->Y[1], t->Y[2]);
    area = -area;
  }
  t->area = area;
  t->e01 = make_edge(t->X[0], t->Y[0], t->X[1], t->Y[1]);
  t->e12 = make_edge(t->X[1], t->Y[1], t->X[2], t->Y[2]);
  t->e20 = make_edge(t->X[2], t->Y[2], t->X[0], t->Y[0]);
  return true;
}

inline bool sample_inside(const RasterTri& t, int64_t sx, int64_t sy) {
  return t.e01.eval(sx, sy) + t.e01.bias >= 0 && t.e12.eval(sx, sy) + t.e12.bias >= 0 &&
         t.e20.eval(sx, sy) + t.e20.bias >= 0;
}

/* max of the edge function over the closed pixel square [x0,x0+256]x[y0,y0+256] */
inline int64_t edge_max_over_pixel(const Edge& e, int64_t x0, int64_t y0) {
  int64_t px = (-e.dy > 0) ? x0 + SUBPIX : x0;
  int64_t py = (e.dx > 0) ? y0 + SUBPIX : y0;
  return e.eval(px, py);
}

const int MSAA4_X[4] = {96, 224, 32, 160};
const int MSAA4_Y[4] = {32, 96, 160, 224};

inline bool pixel_covered(const RasterTri& t, int i, int j, int policy) {
  int64_t x0 = (int64_t)i * SUBPIX, y0 = (int64_t)j * SUBPIX;
  if (policy == 0) return sample_inside(t, x0 + 128, y0 + 128);
  if (policy == 1) {
    for (int s = 0; s < 4; ++s)
      if (sample_inside(t, x0 + MSAA4_X[s], y0 + MSAA4_Y[s])) return true;
    return false;
  }
  /* CONSERVATIVE: the open pixel square intersects the open triangle */
  return edge_max_over_pixel(t.e01, x0, y0) > 0 && edge_max_over_pixel(t.e12, x0, y0) > 0 &&
         edge_max_over_pixel(t.e20, x0, y0) > 0;
}

inline int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

/* pixel bounding box [i0,i1] x [j0,j1] clipped to the viewport; false if empty */
bool raster_bbox(const RasterTri& t, int policy, int W, int H, int* i0, int* i1, int* j0, int* j1) {
  int64_t minx = std::min(t.X[0], std::min(t.X[1], t.X[2])), maxx = std::max(t.X[0], std::max(t.X[1], t.X[2]));
  int64_t miny = std::min(t.Y[0], std::min(t.Y[1], t.Y[2])), maxy = std::max(t.Y[0], std::max(t.Y[1], t.Y[2]));
  int64_t a0, a1, b0, b1;
  if (policy == 2) {
    /* pixels whose open square overlaps (min,max) */
    a0 = floor_div(minx, SUBPIX); a1 = floor_div(maxx - 1, SUBPIX);
    b0 = floor_div(miny, SUBPIX); b1 = floor_div(maxy - 1, SUBPIX);
  } else {
    /* any sample position lies in [0,256) of its pixel */
    a0 = floor_div(minx - (SUBPIX - 1), SUBPIX); a1 = floor_div(maxx, SUBPIX);
    b0 = floor_div(miny - (SUBPIX - 1), SUBPIX); b1 = floor_div(maxy, SUBPIX);
  }
  a0 = std::max<int64_t>(a0, 0); b0 = std::max<int64_t>(b0, 0);
  a1 = std::min<int64_t>(a1, W - 1); b1 = std::min<int64_t>(b1, H - 1);
  if (a0 > a1 || b0 > b1) return false;
  *i0 = (int)a0; *i1 = (int)a1; *j0 = (int)b0; *j1 = (int)b1;
  return true;
}

/* barycentric weights of v1 and v2 at the pixel centre (may be outside [0,1] for MSAA/conservative) */
inline void pixel_lambdas(const RasterTri& t, int i, int j, float* l1, float* l2) {
  int64_t sx = (int64_t)i * SUBPIX + 128, sy = (int64_t)j * SUBPIX + 128;
  float fa = (float)t.area;
  *l1 = (float)t.e20.eval(sx, sy) / fa;
  *l2 = (float)t.e01.eval(sx, sy) / fa;
}

inline float interp(float a0, float a1, float a2, float l1, float l2) {
  return (a0 + l1 * (a1 - a0)) + l2 * (a2 - a0);
}

/* ------------------------------------------------------------------ S1: shadow map */

}  // namespace

extern "C" int orc_draw_depth(orc_ctx* o) {
  /* DrawDepthTexture, Voxel_Cone_Tracing.h:192-211 + Shadow.vs:7-10: cull back faces, depth LESS,
   * clear to 1.0, D24. */
  const int S = o->p.ShadowMapSize;
  o->depth.assign((size_t)S * S, 0xFFFFFFu);
  const size_t nt = o->idx.size() / 3;
  const float* M = o->p.DepthModelViewProjectionMatrix;
#pragma omp parallel for schedule(dynamic, 256)
  for (long long ti = 0; ti < (long long)nt; ++ti) {
    float wx[3], wy[3], wz[3];
    bool ok = true;
    for (int k = 0; k < 3; ++k) {
      const float* v = &o->verts[(size_t)o->idx[ti * 3 + k] * 14];
      V4 c = mul_mat_vec(M, v[0], v[1], v[2], 1.0f);
      if (!(c.w > 0.0f)) { ok = false; break; }
      float nx = c.x / c.w, ny = c.y / c.w, nz = c.z / c.w;
      wx[k] = (nx * 0.5f + 0.5f) * (float)S;
      wy[k] = (ny * 0.5f + 0.5f) * (float)S;
      wz[k] = nz * 0.5f + 0.5f;
    }
    if (!ok) continue;
    RasterTri t;
    if (!setup_raster(wx, wy, &t)) continue;
    if (t.flipped) continue; /* back face (clockwise in window space), main.cpp:57-58 / :194 */
    int i0, i1, j0, j1;
    if (!raster_bbox(t, 0, S, S, &i0, &i1, &j0, &j1)) continue;
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) {
        if (!pixel_covered(t, i, j, 0)) continue;
        float l1, l2;
        pixel_lambdas(t, i, j, &l1, &l2);
        float z = interp(wz[0], wz[1], wz[2], l1, l2);
        if (!(z >= 0.0f) || z > 1.0f) continue; /* near/far clip */
        uint32_t d = (uint32_t)std::lrintf(z * 16777215.0f);
        atomic_min_u32(&o->depth[(size_t)j * S + i], d);
      }
  }
  return 0;
}

namespace {

/* ------------------------------------------------------------------ V1..V4: voxelisation */

/* Voxelization.gs:25-39.  Decision taken on the un-normalised |cross|; all-zero or NaN -> axis 3 (the
 * reference's normalize() of a zero vector yields NaN and every comparison fails). */
int select_axis(V3 w0, V3 w1, V3 w2) {
  V3 e1 = sub3(w0, w1), e2 = sub3(w2, w0);
  V3 n = cross3(e1, e2);
  float nx = std::fabs(n.x), ny = std::fabs(n.y), nz = std::fabs(n.z);
  if (!(nx == nx) || !(ny == ny) || !(nz == nz)) return 3;
  if (nx == 0.0f && ny == 0.0f && nz == 0.0f) return 3;
  if (nx >= ny && nx >= nz) return 1;
  if (ny >= nx && ny >= nz) return 2;
  return 3;
}

struct Fragment { uint32_t voxel; uint8_t r, g, b; };

template <class Emit>
void voxelize_triangle(const orc_ctx* o, size_t ti, Emit&& emit) {
  const orc_params& p = o->p;
  const int V = p.VoxelDimensions;
  V4 world[3], dc[3];
  float uv[3][2];
  for (int k = 0; k < 3; ++k) {
    const float* v = &o->verts[(size_t)o->idx[ti * 3 + k] * 14];
    world[k] = mul_mat_vec(p.ModelMatrix, v[0], v[1], v[2], 1.0f);                /* Voxelization.vs:21 */
    dc[k] = mul_mat_vec(p.DepthModelViewProjectionMatrix, v[0], v[1], v[2], 1.0f); /* Voxelization.vs:18 */
    dc[k].x = dc[k].x * 0.5f + 0.5f; dc[k].y = dc[k].y * 0.5f + 0.5f; dc[k].z = dc[k].z * 0.5f + 0.5f; /* :19 */
    uv[k][0] = v[6]; uv[k][1] = v[7];
  }
  int axis = select_axis({world[0].x, world[0].y, world[0].z}, {world[1].x, world[1].y, world[1].z},
                         {world[2].x, world[2].y, world[2].z});
  const float* P = axis == 1 ? p.ProjX : axis == 2 ? p.ProjY : p.ProjZ;             /* Voxelization.gs:41 */
  float wx[3], wy[3], wz[3];
  for (int k = 0; k < 3; ++k) {
    V4 c = mul_mat_vec(P, world[k].x, world[k].y, world[k].z, world[k].w);
    if (!(c.w > 0.0f)) return;
    float nx = c.x / c.w, ny = c.y / c.w, nz = c.z / c.w;
    wx[k] = (nx * 0.5f + 0.5f) * (float)V;  /* glViewport(0,0,V,V), Voxel_Cone_Tracing.h:218 */
    wy[k] = (ny * 0.5f + 0.5f) * (float)V;
    wz[k] = nz * 0.5f + 0.5f;
  }
  RasterTri t;
  if (!setup_raster(wx, wy, &t)) return;
  int a = 1, b = 2;
  if (t.flipped) std::swap(a, b);  /* attribute order follows the vertex swap; culling is off (:215) */
  const float z0 = wz[0], z1 = wz[a], z2 = wz[b];
  const float zmin = std::min(z0, std::min(z1, z2)), zmax = std::max(z0, std::max(z1, z2));
  int i0, i1, j0, j1;
  if (!raster_bbox(t, p.CoveragePolicy, V, V, &i0, &i1, &j0, &j1)) return;

  /* implicit LOD of texture(DiffuseTexture, uv): uv is affine in window space for this triangle */
  const Material& m = o->materials[o->tri_mat.empty() ? 0 : o->tri_mat[ti]];
  const Texture* tex = (m.diffuse >= 0 && m.diffuse < (int)o->textures.size() && o->textures[m.diffuse].levels)
                           ? &o->textures[m.diffuse] : nullptr;
  float fa = (float)t.area;
  float dl1dx = (float)(-(t.e20.dy) * SUBPIX) / fa, dl1dy = (float)(t.e20.dx * SUBPIX) / fa;
  float dl2dx = (float)(-(t.e01.dy) * SUBPIX) / fa, dl2dy = (float)(t.e01.dx * SUBPIX) / fa;
  float du1 = uv[a][0] - uv[0][0], du2 = uv[b][0] - uv[0][0];
  float dv1 = uv[a][1] - uv[0][1], dv2 = uv[b][1] - uv[0][1];
  float lod = 0.0f;
  if (tex) {
    float dudx = dl1dx * du1 + dl2dx * du2, dvdx = dl1dx * dv1 + dl2dx * dv2;
    float dudy = dl1dy * du1 + dl2dy * du2, dvdy = dl1dy * dv1 + dl2dy * dv2;
    lod = lod_from_derivs(dudx, dvdx, dudy, dvdy, tex->w[0], tex->h[0]);
  }

  for (int j = j0; j <= j1; ++j)
    for (int i = i0; i <= i1; ++i) {
      if (!pixel_covered(t, i, j, p.CoveragePolicy)) continue;
      float l1, l2;
      pixel_lambdas(t, i, j, &l1, &l2);
      float z = interp(z0, z1, z2, l1, l2);
      if (p.CoveragePolicy == 2) z = std::min(std::max(z, zmin), zmax);
      float tz = (float)V * z;                         /* Voxelization.fs:58 */
      if (!(tz >= 0.0f) || !(tz < (float)V)) continue; /* clipped, or imageStore out of bounds */
      int cx = i, cy = j, cz = (int)tz;
      int vx, vy, vz;
      if (axis == 1) { vx = V - 1 - cz; vz = V - 1 - cx; vy = cy; }        /* Voxelization.fs:70-75 */
      else if (axis == 2) { vz = V - 1 - cy; vy = V - 1 - cz; vx = cx; }   /* :76-81 */
      else { vx = cx; vy = cy; vz = V - 1 - cz; }                          /* :82-86 */
      if (vx < 0 || vy < 0 || vz < 0 || vx >= V || vy >= V || vz >= V) continue;

      float u = interp(uv[0][0], uv[a][0], uv[b][0], l1, l2);
      float vv = interp(uv[0][1], uv[a][1], uv[b][1], l1, l2);
      float col[4] = {1, 1, 1, 1};
      if (tex) sample_tex(*tex, u, vv, lod, col);      /* Voxelization.fs:56 */
      float dx = interp(dc[0].x, dc[a].x, dc[b].x, l1, l2);
      float dy = interp(dc[0].y, dc[a].y, dc[b].y, l1, l2);
      float dz = interp(dc[0].z, dc[a].z, dc[b].z, l1, l2);
      float dw = interp(dc[0].w, dc[a].w, dc[b].w, l1, l2);
      int taps = (2 * p.PcfRadius + 1) * (2 * p.PcfRadius + 1);
      float shadow = pcf_lit_taps(o, dx, dy, dz, dw) / (float)taps; /* Voxelization.fs:46 */
      Fragment f;
      f.voxel = (uint32_t)(((size_t)vz * V + vy) * V + vx);
      float c3[3] = {col[0] * shadow, col[1] * shadow, col[2] * shadow};  /* Voxelization.fs:88 */
      uint8_t q[3];
      for (int c = 0; c < 3; ++c) {
        float x = std::min(std::max(c3[c], 0.0f), 1.0f);
        q[c] = (uint8_t)std::lrintf(x * 255.0f);       /* unorm8 conversion of imageStore */
      }
      f.r = q[0]; f.g = q[1]; f.b = q[2];
      emit(f);
    }
}

void ensure_grid(orc_ctx* o) {
  const int V = o->p.VoxelDimensions;
  if (o->gridV == V && o->gridFmt == o->p.GridFormat && !o->grid.empty()) return;
  o->gridV = V;
  o->gridFmt = o->p.GridFormat;
  o->grid.clear();
  for (int s = V; s >= 1; s >>= 1) o->grid.emplace_back((size_t)s * s * s * o->bpt(), 0);
  o->counts.assign((size_t)V * V * V, 0);
  o->sums.assign((size_t)V * V * V * 3, 0);
}

}  // namespace

extern "C" int orc_draw_voxels_range(orc_ctx* o, size_t tb, size_t te, int clear_first) {
  ensure_grid(o);
  const int V = o->p.VoxelDimensions;
  if (clear_first) {
    std::fill(o->counts.begin(), o->counts.end(), 0u);
    std::fill(o->sums.begin(), o->sums.end(), 0u);
    o->fragments = 0;
  }
  te = std::min(te, o->idx.size() / 3);
  if (o->p.VoxelStoreMode == 1) {
    /* last writer in primitive order (one legal outcome of the unordered imageStore, Voxelization.fs:88) */
    uint64_t nf = 0;
    for (size_t ti = tb; ti < te; ++ti)
      voxelize_triangle(o, ti, [&](const Fragment& f) {
        o->counts[f.voxel] = 1;
        o->sums[(size_t)f.voxel * 3 + 0] = f.r; o->sums[(size_t)f.voxel * 3 + 1] = f.g; o->sums[(size_t)f.voxel * 3 + 2] = f.b;
        ++nf;
      });
    o->fragments += nf;
    return 0;
  }
  uint64_t nf = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nf)
  for (long long ti = (long long)tb; ti < (long long)te; ++ti)
    voxelize_triangle(o, (size_t)ti, [&](const Fragment& f) {
      uint32_t* s = &o->sums[(size_t)f.voxel * 3];
      __atomic_fetch_add(&o->counts[f.voxel], 1u, __ATOMIC_RELAXED);
      __atomic_fetch_add(&s[0], (uint32_t)f.r, __ATOMIC_RELAXED);
      __atomic_fetch_add(&s[1], (uint32_t)f.g, __ATOMIC_RELAXED);
      __atomic_fetch_add(&s[2], (uint32_t)f.b, __ATOMIC_RELAXED);
      ++nf;
    });
  o->fragments += nf;
  (void)V;
  return 0;
}

extern "C" int orc_build_mips(orc_ctx* o) {
  /* glGenerateMipmap(GL_TEXTURE_3D), Voxel_Cone_Tracing.h:246-248: 2x2x2 box, round to nearest. */
  ensure_grid(o);
  int s = o->p.VoxelDimensions;
  if (o->p.GridFormat == 1) {
    /* RGBA16F: fp32 sum of the 8 parents in the fixed order ((a00 + a10) + a01) + a11, a_yz = the x pair, times
     * 0.125, rounded to half (nearest even) */
    for (size_t l = 1; l < o->grid.size(); ++l) {
      int ps = s;
      s >>= 1;
#pragma omp parallel for schedule(static) if (s >= 16)
      for (int z = 0; z < s; ++z)
        for (int y = 0; y < s; ++y)
          for (int x = 0; x < s; ++x)
            for (int c = 0; c < 4; ++c) {
              float acc = 0.0f;
              bool first = true;
              for (int dz = 0; dz < 2; ++dz)
                for (int dy = 0; dy < 2; ++dy) {
                  size_t i0 = (((size_t)(2 * z + dz)) * ps + (2 * y + dy)) * ps + 2 * x;
                  float a = o->texel((int)l - 1, i0, c) + o->texel((int)l - 1, i0 + 1, c);
                  acc = first ? a : acc + a;
                  first = false;
                }
              o->set_texel16((int)l, ((size_t)z * s + y) * s + x, c, acc * 0.125f);
            }
    }
    return 0;
  }
  for (size_t l = 1; l < o->grid.size(); ++l) {
    int ps = s;
    s >>= 1;
    const uint8_t* src = o->grid[l - 1].data();
    uint8_t* dst = o->grid[l].data();
#pragma omp parallel for schedule(static) if (s >= 16)
    for (int z = 0; z < s; ++z)
      for (int y = 0; y < s; ++y)
        for (int x = 0; x < s; ++x)
          for (int c = 0; c < 4; ++c) {
            int acc = 0;
            for (int dz = 0; dz < 2; ++dz)
              for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx)
                  acc += src[((((size_t)(2 * z + dz)) * ps + (2 * y + dy)) * ps + (2 * x + dx)) * 4 + c];
            dst[(((size_t)z * s + y) * s + x) * 4 + c] = (uint8_t)((acc + 4) >> 3);
          }
  }
  return 0;
}

namespace {

/* ------------------------------------------------------------------ C1: voxel texture sampling */

void sample_grid_level(const orc_ctx* o, int l, float u, float v, float w, float out[4]) {
  const int N = o->p.VoxelDimensions >> l;
  float x = u * (float)N - 0.5f, y = v * (float)N - 0.5f, z = w * (float)N - 0.5f;
  float fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
  float a = x - fx, b = y - fy, c = z - fz;
  if (o->p.FilterMode == 1) { /* 8-bit fixed-point filter weights, as texture hardware uses */
    a = std::floor(a * 256.0f + 0.5f) * (1.0f / 256.0f);
    b = std::floor(b * 256.0f + 0.5f) * (1.0f / 256.0f);
    c = std::floor(c * 256.0f + 0.5f) * (1.0f / 256.0f);
  }
  /* GL_REPEAT on s,t,r: the reference never sets a wrap mode (Voxel_Cone_Tracing.h:110-113) */
  int i0 = wrap_repeat((int)fx, N), i1 = wrap_repeat((int)fx + 1, N);
  int j0 = wrap_repeat((int)fy, N), j1 = wrap_repeat((int)fy + 1, N);
  int k0 = wrap_repeat((int)fz, N), k1 = wrap_repeat((int)fz + 1, N);
  auto at = [&](int i, int j, int k, int ch) { return o->texel(l, ((size_t)k * N + j) * N + i, ch); };
  for (int ch = 0; ch < 4; ++ch) {
    float c00 = at(i0, j0, k0, ch) + a * (at(i1, j0, k0, ch) - at(i0, j0, k0, ch));
    float c10 = at(i0, j1, k0, ch) + a * (at(i1, j1, k0, ch) - at(i0, j1, k0, ch));
    float c01 = at(i0, j0, k1, ch) + a * (at(i1, j0, k1, ch) - at(i0, j0, k1, ch));
    float c11 = at(i0, j1, k1, ch) + a * (at(i1, j1, k1, ch) - at(i0, j1, k1, ch));
    float c0 = c00 + b * (c10 - c00);
    float c1 = c01 + b * (c11 - c01);
    out[ch] = c0 + c * (c1 - c0);
  }
}

/* SampleVoxels, VoxelConeTracing.fs:59-66 + textureLod with LINEAR_MIPMAP_LINEAR (Voxel_Cone_Tracing.h:112) */
void sample_voxels(const orc_ctx* o, V3 pos, float lod, float out[4]) {
  const float half = o->p.VoxelGridWorldSize * 0.5f;
  float u = (pos.x / half) * 0.5f + 0.5f;
  float v = (pos.y / half) * 0.5f + 0.5f;
  float w = (pos.z / half) * 0.5f + 0.5f;
  if (!std::isfinite(u) || !std::isfinite(v) || !std::isfinite(w)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  /* keep (int) conversions in range for far away points: REPEAT makes the integer part irrelevant */
  u -= std::floor(u); v -= std::floor(v); w -= std::floor(w);
  const int maxl = (int)o->grid.size() - 1;
  if (!(lod > 0.0f)) lod = 0.0f;
  if (lod > (float)maxl) lod = (float)maxl;
  int l0 = (int)std::floor(lod);
  float f = lod - (float)l0;
  if (o->p.FilterMode == 1) f = std::floor(f * 256.0f) * (1.0f / 256.0f);   /* hardware truncates the LOD fraction */
  float a[4];
  sample_grid_level(o, l0, u, v, w, a);
  if (f > 0.0f && l0 < maxl) {
    float b[4];
    sample_grid_level(o, l0 + 1, u, v, w, b);
    for (int c = 0; c < 4; ++c) out[c] = a[c] + f * (b[c] - a[c]);
  } else {
    for (int c = 0; c < 4; ++c) out[c] = a[c];
  }
}

/* Voxel_Cone_Tracing(direction, tanHalfAngle), VoxelConeTracing.fs:82-107.  start = Position_world +
 * Normal_world * voxelWorldSize is computed by the caller (:92). */
void cone_trace(const orc_ctx* o, V3 start, V3 dir, float tanHalf, float out[4], uint64_t* samples) {
  const orc_params& p = o->p;
  float cr = 0, cg = 0, cb = 0, alpha = 0, occ = 0;
  float vws = p.VoxelGridWorldSize / (float)p.VoxelDimensions;
  float dist = vws;
  uint64_t n = 0;
  while (dist < p.MaxDistance && alpha < p.MaxAlpha) {
    float diameter = std::max(vws, 2.0f * tanHalf * dist);
    float lod = std::log2(diameter / vws);
    V3 pos = add3(start, scale3(dir, dist));
    float s[4];
    sample_voxels(o, pos, lod, s);
    float k = 1.0f - alpha;
    cr += k * s[0]; cg += k * s[1]; cb += k * s[2];
    occ += (k * s[3]) / (1.0f + 0.03f * diameter);
    alpha += k * s[3];
    dist += diameter * p.StepMultiplier;
    ++n;
  }
  out[0] = cr; out[1] = cg; out[2] = cb; out[3] = occ;
  *samples += n;
}

/* ------------------------------------------------------------------ S2: primary visibility */

struct HVert { float X, Y, w, zc; };  /* window-homogeneous x,y ; clip w ; clip z */

struct HEdge { float A, B, C; };

inline HEdge hcross(const HVert& a, const HVert& b) {
  HEdge e;
  e.A = a.Y * b.w - b.Y * a.w;
  e.B = b.X * a.w - a.X * b.w;
  e.C = a.X * b.Y - b.X * a.Y;
  return e;
}
inline float heval(const HEdge& e, float px, float py) { return (e.A * px + e.B * py) + e.C; }
inline bool hinside(const HEdge& e, float v) {
  return v > 0.0f || (v == 0.0f && (e.A > 0.0f || (e.A == 0.0f && e.B > 0.0f)));
}

struct HTri {
  HVert c[3];
  HEdge e[3];   /* e[i] is the edge opposite vertex i */
  int i0, i1, j0, j1;
  float ox, oy; /* origin the edge functions are expressed in: (0,0) = the window's (the defined semantics);
                 * RasterOrigin = 1 (experiment): a pixel corner next to the triangle */
};

/* VoxelConeTracing.vs:25 + viewport + back-face cull (main.cpp:57-58) as 2D homogeneous rasterisation
 * set-up; returns false if the triangle cannot produce fragments. */
bool setup_htri(const orc_ctx* o, size_t ti, HTri* t) {
  const orc_params& p = o->p;
  const float W = (float)p.screen_width, H = (float)p.screen_height;
  bool in_near[3];
  float cz[3];
  for (int k = 0; k < 3; ++k) {
    const float* v = &o->verts[(size_t)o->idx[ti * 3 + k] * 14];
    V4 e = mul_mat_vec(p.ModelViewMatrix, v[0], v[1], v[2], 1.0f);
    V4 c = mul_mat_vec(p.ProjectionMatrix, e.x, e.y, e.z, e.w);
    if (!std::isfinite(c.x) || !std::isfinite(c.y) || !std::isfinite(c.z) || !std::isfinite(c.w)) return false;
    t->c[k].X = (c.x + c.w) * (0.5f * W);
    t->c[k].Y = (c.y + c.w) * (0.5f * H);
    t->c[k].w = c.w;
    t->c[k].zc = c.z;
    cz[k] = c.z;
    in_near[k] = (c.z >= -c.w) && (c.w > 0.0f);
  }
  if (!in_near[0] && !in_near[1] && !in_near[2]) return false;
  t->ox = t->oy = 0.0f;
  HVert l[3] = {t->c[0], t->c[1], t->c[2]};
  if (p.RasterOrigin == 1) {
    /* same edge functions, written in a frame whose origin is the pixel corner below-left of the first vertex in front
     * of the eye: the products that cancel in hcross / heval are then a few pixels times w instead of the window size
     * times w, i.e. ~2.5 decimal digits more of the float32 mantissa go to the triangle itself */
    for (int k = 0; k < 3; ++k)
      if (t->c[k].w > 0.0f) { t->ox = std::floor(t->c[k].X / t->c[k].w); t->oy = std::floor(t->c[k].Y / t->c[k].w); break; }
    for (int k = 0; k < 3; ++k) { l[k].X = t->c[k].X - t->ox * t->c[k].w; l[k].Y = t->c[k].Y - t->oy * t->c[k].w; }
  }
  t->e[0] = hcross(l[1], l[2]);
  t->e[1] = hcross(l[2], l[0]);
  t->e[2] = hcross(l[0], l[1]);
  float det = (l[0].X * t->e[0].A + l[0].Y * t->e[0].B) + l[0].w * t->e[0].C;
  if (!(det > 0.0f)) return false;  /* back face or degenerate */

  /* conservative screen bounding box of the near-clipped polygon */
  float minx = std::numeric_limits<float>::infinity(), maxx = -minx, miny = minx, maxy = -minx;
  auto add_pt = [&](float X, float Y, float w) {
    float x = X / w, y = Y / w;
    minx = std::min(minx, x); maxx = std::max(maxx, x);
    miny = std::min(miny, y); maxy = std::max(maxy, y);
  };
  for (int k = 0; k < 3; ++k) {
    int n = (k + 1) % 3;
    if (in_near[k]) add_pt(t->c[k].X, t->c[k].Y, t->c[k].w);
    if (in_near[k] != in_near[n]) {
      float da = cz[k] + t->c[k].w, db = cz[n] + t->c[n].w;
      float s = da / (da - db);
      float X = t->c[k].X + s * (t->c[n].X - t->c[k].X);
      float Y = t->c[k].Y + s * (t->c[n].Y - t->c[k].Y);
      float w = t->c[k].w + s * (t->c[n].w - t->c[k].w);
      if (w > 0.0f) add_pt(X, Y, w);
      else { minx = miny = -1e30f; maxx = maxy = 1e30f; }
    }
  }
  if (!(minx <= maxx)) return false;
  float fx0 = std::floor(minx) - 1.0f, fx1 = std::ceil(maxx) + 1.0f;
  float fy0 = std::floor(miny) - 1.0f, fy1 = std::ceil(maxy) + 1.0f;
  fx0 = std::max(fx0, 0.0f); fy0 = std::max(fy0, 0.0f);
  fx1 = std::min(fx1, W - 1.0f); fy1 = std::min(fy1, H - 1.0f);
  if (!(fx0 <= fx1) || !(fy0 <= fy1)) return false;
  t->i0 = (int)fx0; t->i1 = (int)fx1; t->j0 = (int)fy0; t->j1 = (int)fy1;
  return true;
}

/* perspective-correct barycentrics at a pixel centre; false if outside */
inline bool hbary(const HTri& t, float px, float py, float b[3], bool test) {
  px -= t.ox; py -= t.oy;
  float e0 = heval(t.e[0], px, py), e1 = heval(t.e[1], px, py), e2 = heval(t.e[2], px, py);
  if (test && !(hinside(t.e[0], e0) && hinside(t.e[1], e1) && hinside(t.e[2], e2))) return false;
  float s = (e0 + e1) + e2;
  if (test && !(s > 0.0f)) return false;
  b[0] = e0 / s; b[1] = e1 / s; b[2] = e2 / s;
  return true;
}

inline float bary3(const float b[3], float a0, float a1, float a2) { return (b[0] * a0 + b[1] * a1) + b[2] * a2; }

struct PixelUV { float u, v, dudx, dvdx, dudy, dvdy; };

inline PixelUV pixel_uv(const orc_ctx* o, size_t ti, const HTri& t, float px, float py, const float b[3]) {
  const float* v0 = &o->verts[(size_t)o->idx[ti * 3 + 0] * 14];
  const float* v1 = &o->verts[(size_t)o->idx[ti * 3 + 1] * 14];
  const float* v2 = &o->verts[(size_t)o->idx[ti * 3 + 2] * 14];
  PixelUV r;
  r.u = bary3(b, v0[6], v1[6], v2[6]);
  r.v = bary3(b, v0[7], v1[7], v2[7]);
  float bx[3], by[3];
  hbary(t, px + 1.0f, py, bx, false);
  hbary(t, px, py + 1.0f, by, false);
  r.dudx = bary3(bx, v0[6], v1[6], v2[6]) - r.u;
  r.dvdx = bary3(bx, v0[7], v1[7], v2[7]) - r.v;
  r.dudy = bary3(by, v0[6], v1[6], v2[6]) - r.u;
  r.dvdy = bary3(by, v0[7], v1[7], v2[7]) - r.v;
  return r;
}

inline const Texture* get_tex(const orc_ctx* o, int id) {
  if (id < 0 || id >= (int)o->textures.size() || o->textures[id].levels == 0) return nullptr;
  return &o->textures[id];
}

inline void sample_material(const orc_ctx* o, int id, const PixelUV& q, float du, float dv, float out[4]) {
  const Texture* t = get_tex(o, id);
  if (!t) { out[0] = out[1] = out[2] = out[3] = 1.0f; return; }
  float lod = lod_from_derivs(q.dudx, q.dvdx, q.dudy, q.dvdy, t->w[0], t->h[0]);
  sample_tex(*t, q.u + du, q.v + dv, lod, out);
}

inline uint32_t float_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

}  // namespace

static void orc_visibility(orc_ctx* o, int y0, int y1) {
  const int W = o->p.screen_width, H = o->p.screen_height;
  o->vis.assign((size_t)W * H, ~0ull);
  const size_t nt = o->idx.size() / 3;
#pragma omp parallel for schedule(dynamic, 64)
  for (long long ti = 0; ti < (long long)nt; ++ti) {
    HTri t;
    if (!setup_htri(o, (size_t)ti, &t)) continue;
    const Material& m = o->materials[o->tri_mat.empty() ? 0 : o->tri_mat[ti]];
    const Texture* dt = get_tex(o, m.diffuse);
    const bool alpha_test = dt && dt->has_alpha;
    for (int j = std::max(t.j0, y0); j <= std::min(t.j1, y1 - 1); ++j)
      for (int i = t.i0; i <= t.i1; ++i) {
        float px = (float)i + 0.5f, py = (float)j + 0.5f;
        /* window depth = z_clip / w_clip at the pixel; both are linear in the (unnormalised) homogeneous edge
         * functions, so the common 1/sum cancels: one division per fragment */
        float lx = px - t.ox, ly = py - t.oy;
        float e0 = heval(t.e[0], lx, ly), e1 = heval(t.e[1], lx, ly), e2 = heval(t.e[2], lx, ly);
        if (!(hinside(t.e[0], e0) && hinside(t.e[1], e1) && hinside(t.e[2], e2))) continue;
        if (!((e0 + e1) + e2 > 0.0f)) continue;
        float zc = (e0 * t.c[0].zc + e1 * t.c[1].zc) + e2 * t.c[2].zc;
        float w = (e0 * t.c[0].w + e1 * t.c[1].w) + e2 * t.c[2].w;
        float zw = (zc / w) * 0.5f + 0.5f;
        if (!(zw >= 0.0f) || zw > 1.0f) continue;           /* near / far clip */
        float b[3];
        if (alpha_test) {                                   /* discard, VoxelConeTracing.fs:167-172 */
          hbary(t, px, py, b, false);
          PixelUV q = pixel_uv(o, (size_t)ti, t, px, py, b);
          float c[4];
          sample_material(o, m.diffuse, q, 0, 0, c);
          if (c[3] < 0.5f) continue;
        }
        uint64_t key = ((uint64_t)float_bits(zw) << 32) | (uint32_t)ti;  /* depth LESS, ties: lowest id */
        atomic_min_u64(&o->vis[(size_t)j * W + i], key);
      }
  }
}

/* optional per-pixel trace for diagnostics: 8 cones x (start3, dir3, tan1, result4) + shadow, N3 */
static float* g_debug = nullptr;

/* VoxelConeTracing.fs:165-229 for one visible pixel */
static void shade_pixel(const orc_ctx* o, size_t ti, int i, int j, uint8_t out[4], uint64_t* samples) {
  const orc_params& p = o->p;
  HTri t;
  setup_htri(o, ti, &t);
  float px = (float)i + 0.5f, py = (float)j + 0.5f;
  float b[3];
  hbary(t, px, py, b, false);
  const float* vv[3];
  for (int k = 0; k < 3; ++k) vv[k] = &o->verts[(size_t)o->idx[ti * 3 + k] * 14];
  /* vertex shader outputs, VoxelConeTracing.vs:27-34, then perspective-correct interpolation */
  V4 pw[3], pd[3], nw[3], tw[3], bw[3];
  for (int k = 0; k < 3; ++k) {
    pw[k] = mul_mat_vec(p.ModelMatrix, vv[k][0], vv[k][1], vv[k][2], 1.0f);
    pd[k] = mul_mat_vec(p.DepthModelViewProjectionMatrix, vv[k][0], vv[k][1], vv[k][2], 1.0f);
    pd[k].x = pd[k].x * 0.5f + 0.5f; pd[k].y = pd[k].y * 0.5f + 0.5f; pd[k].z = pd[k].z * 0.5f + 0.5f;
    nw[k] = mul_mat_vec(p.ModelMatrix, vv[k][3], vv[k][4], vv[k][5], 0.0f);
    tw[k] = mul_mat_vec(p.ModelMatrix, vv[k][8], vv[k][9], vv[k][10], 0.0f);
    bw[k] = mul_mat_vec(p.ModelMatrix, vv[k][11], vv[k][12], vv[k][13], 0.0f);
  }
  V3 Pw = {bary3(b, pw[0].x, pw[1].x, pw[2].x), bary3(b, pw[0].y, pw[1].y, pw[2].y), bary3(b, pw[0].z, pw[1].z, pw[2].z)};
  V4 Pd = {bary3(b, pd[0].x, pd[1].x, pd[2].x), bary3(b, pd[0].y, pd[1].y, pd[2].y),
           bary3(b, pd[0].z, pd[1].z, pd[2].z), bary3(b, pd[0].w, pd[1].w, pd[2].w)};
  V3 Nw = {bary3(b, nw[0].x, nw[1].x, nw[2].x), bary3(b, nw[0].y, nw[1].y, nw[2].y), bary3(b, nw[0].z, nw[1].z, nw[2].z)};
  V3 Tw = {bary3(b, tw[0].x, tw[1].x, tw[2].x), bary3(b, tw[0].y, tw[1].y, tw[2].y), bary3(b, tw[0].z, tw[1].z, tw[2].z)};
  V3 Bw = {bary3(b, bw[0].x, bw[1].x, bw[2].x), bary3(b, bw[0].y, bw[1].y, bw[2].y), bary3(b, bw[0].z, bw[1].z, bw[2].z)};
  V3 cam = {p.CameraPosition[0], p.CameraPosition[1], p.CameraPosition[2]};
  /* CameraDirection_world is CameraPosition - Position_world per vertex (:34); interpolation is linear */
  V3 Cd = sub3(cam, Pw);

  const Material& m = o->materials[o->tri_mat.empty() ? 0 : o->tri_mat[ti]];
  PixelUV q = pixel_uv(o, ti, t, px, py, b);
  float mat[4];
  sample_material(o, m.diffuse, q, 0, 0, mat);                         /* :167 */

  /* TBN = inverse(transpose(mat3(T,B,N))), :175.  transpose(mat3(T,B,N)) has ROWS T,B,N. */
  float r0[3] = {Tw.x, Tw.y, Tw.z}, r1[3] = {Bw.x, Bw.y, Bw.z}, r2[3] = {Nw.x, Nw.y, Nw.z};
  float c00 = r1[1] * r2[2] - r1[2] * r2[1];
  float c01 = r1[2] * r2[0] - r1[0] * r2[2];
  float c02 = r1[0] * r2[1] - r1[1] * r2[0];
  float det = (r0[0] * c00 + r0[1] * c01) + r0[2] * c02;
  float id = 1.0f / det;
  /* inverse = adjugate / det ; inv[r][c] */
  float inv[3][3];
  inv[0][0] = c00 * id;
  inv[1][0] = c01 * id;
  inv[2][0] = c02 * id;
  inv[0][1] = (r0[2] * r2[1] - r0[1] * r2[2]) * id;
  inv[1][1] = (r0[0] * r2[2] - r0[2] * r2[0]) * id;
  inv[2][1] = (r0[1] * r2[0] - r0[0] * r2[1]) * id;
  inv[0][2] = (r0[1] * r1[2] - r0[2] * r1[1]) * id;
  inv[1][2] = (r0[2] * r1[0] - r0[0] * r1[2]) * id;
  inv[2][2] = (r0[0] * r1[1] - r0[1] * r1[0]) * id;
  auto tbn_mul = [&](V3 v) {
    return V3{(inv[0][0] * v.x + inv[0][1] * v.y) + inv[0][2] * v.z,
              (inv[1][0] * v.x + inv[1][1] * v.y) + inv[1][2] * v.z,
              (inv[2][0] * v.x + inv[2][1] * v.y) + inv[2][2] * v.z};
  };

  /* CalcBumpNormal, :110-128 */
  const Texture* ht = get_tex(o, m.height);
  float hw = ht ? (float)ht->w[0] : 1.0f, hh = ht ? (float)ht->h[0] : 1.0f;
  float offx = 1.0f / hw, offy = 1.0f / hh;
  float h0[4], hx[4], hy[4];
  sample_material(o, m.height, q, 0, 0, h0);
  sample_material(o, m.height, q, offx, 0, hx);
  sample_material(o, m.height, q, 0, offy, hy);
  float ddx = hx[0] - h0[0], ddy = hy[0] - h0[0];
  V3 t1 = normalize3({1.0f, 0.0f, ddx});
  V3 t2 = normalize3({0.0f, 1.0f, ddy});
  V3 bump = normalize3(cross3(t1, t2));
  V3 N = normalize3(tbn_mul(bump));
  V3 L = normalize3({p.LightDirection[0], p.LightDirection[1], p.LightDirection[2]});  /* :179 */
  V3 E = normalize3(Cd);                                                                 /* :181 */

  /* :186, PCF with the *0.111 normalisation of :158 */
  float shadow = pcf_lit_taps(o, Pd.x, Pd.y, Pd.z, Pd.w) * 0.111f;
  float cos_theta = std::max(dot3(N, L), 0.0f);
  float directDiffuse = shadow * cos_theta;

  float vws = p.VoxelGridWorldSize / (float)p.VoxelDimensions;
  V3 start = add3(Pw, scale3(Nw, vws));                                /* :92 */
  float ind[4] = {0, 0, 0, 0};
  for (int c = 0; c < p.NumDiffuseCones; ++c) {                        /* :196-199 */
    V3 d = {p.ConeDirections[c * 3], p.ConeDirections[c * 3 + 1], p.ConeDirections[c * 3 + 2]};
    V3 dir = normalize3(tbn_mul(d));
    float r[4];
    cone_trace(o, start, dir, p.DiffuseTanHalfAngle, r, samples);
    if (g_debug && c < 7) {
      float* d = g_debug + c * 11;
      d[0] = start.x; d[1] = start.y; d[2] = start.z; d[3] = dir.x; d[4] = dir.y; d[5] = dir.z; d[6] = p.DiffuseTanHalfAngle;
      for (int k = 0; k < 4; ++k) d[7 + k] = r[k];
    }
    for (int k = 0; k < 4; ++k) ind[k] += p.ConeWeights[c] * r[k];
  }
  float occlusion = 1.0f - ind[3];                                     /* :201 */
  float diff[3];
  for (int k = 0; k < 3; ++k) diff[k] = (directDiffuse + occlusion * ind[k]) * mat[k];  /* :205 */

  float sc[4];
  sample_material(o, m.specular, q, 0, 0, sc);                         /* :209 */
  float lgb = std::sqrt(sc[1] * sc[1] + sc[2] * sc[2]);
  if (!(lgb > 0.0f)) { sc[1] = sc[0]; sc[2] = sc[0]; }                 /* .rrra, :210 */
  /* reflect(I,N) = I - 2*dot(N,I)*N */
  V3 negL = {-L.x, -L.y, -L.z};
  float dnl = dot3(N, negL);
  V3 R = normalize3(sub3(negL, scale3(N, 2.0f * dnl)));                /* :212 */
  float spec = std::pow(std::max(dot3(E, R), 0.0f), m.shininess);      /* :213 */
  float directSpec = spec * shadow;                                    /* :214 */
  V3 negE = {-E.x, -E.y, -E.z};
  float dne = dot3(N, negE);
  V3 refl = normalize3(sub3(negE, scale3(N, 2.0f * dne)));             /* :217 */
  float isp[4];
  cone_trace(o, start, refl, p.SpecularTanHalfAngle, isp, samples);    /* :218 */
  if (g_debug) {
    float* d = g_debug + 7 * 11;
    d[0] = start.x; d[1] = start.y; d[2] = start.z; d[3] = refl.x; d[4] = refl.y; d[5] = refl.z; d[6] = p.SpecularTanHalfAngle;
    for (int k = 0; k < 4; ++k) d[7 + k] = isp[k];
    d[11] = shadow; d[12] = N.x; d[13] = N.y; d[14] = N.z; d[15] = mat[0]; d[16] = mat[1]; d[17] = mat[2];
    d[18] = sc[0]; d[19] = sc[1]; d[20] = sc[2]; d[21] = spec;
  }
  float specOcc = 1.0f - isp[3];                                       /* :221 */
  float col[3];
  for (int k = 0; k < 3; ++k) {
    float specR = (isp[k] + specOcc * directSpec) * sc[k];             /* :223 */
    float amb = (p.ambientFactor * mat[k]) * occlusion;                /* :225 */
    col[k] = (amb + diff[k]) + specR;                                  /* :227 */
  }
  for (int k = 0; k < 3; ++k) {
    float x = std::min(std::max(col[k], 0.0f), 1.0f);
    if (!(x == x)) x = 0.0f;
    out[k] = (uint8_t)std::lrintf(x * 255.0f);
  }
  float a = std::min(std::max(mat[3], 0.0f), 1.0f);
  out[3] = (uint8_t)std::lrintf(a * 255.0f);
}

extern "C" int orc_render_rows(orc_ctx* o, int y0, int y1);
extern "C" int orc_render(orc_ctx* o) { return orc_render_rows(o, 0, o->p.screen_height); }

/* rows [y0, y1) only: what one rank does when the frame is sharded by row bands */
extern "C" int orc_render_rows(orc_ctx* o, int y0, int y1) {
  const int W = o->p.screen_width, H = o->p.screen_height;
  y0 = std::max(y0, 0); y1 = std::min(y1, H);
  ensure_grid(o);
  orc_visibility(o, y0, y1);
  o->frame.assign((size_t)W * H * 4, 0);
  /* glClearColor, Voxel_Cone_Tracing.h:156-159 */
  float cc = o->p.ambientFactor < 0.5f ? 0.5f : 1.0f;
  uint8_t bg = (uint8_t)std::lrintf(cc * 255.0f);
  uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
  for (int j = y0; j < y1; ++j)
    for (int i = 0; i < W; ++i) {
      uint64_t key = o->vis[(size_t)j * W + i];
      uint8_t* px = &o->frame[((size_t)j * W + i) * 4];
      if (key == ~0ull) { px[0] = px[1] = px[2] = bg; px[3] = 255; continue; }
      uint64_t n = 0;
      shade_pixel(o, (size_t)(uint32_t)key, i, j, px, &n);
      total += n;
    }
  o->cone_samples = total;
  return 0;
}

/* ------------------------------------------------------------------ resolve + bounce extension */

static void resolve_level0(orc_ctx* o) {
  const size_t n = o->counts.size();
  uint8_t* g = o->grid[0].data();
  if (o->p.GridFormat == 1) {
    /* RGBA16F: rgb = half(sum / (count * 255)), alpha = 1 */
    for (size_t i = 0; i < n; ++i) {
      uint32_t c = o->counts[i];
      for (int k = 0; k < 4; ++k) o->set_texel16(0, i, k, 0.0f);
      if (c == 0) continue;
      for (int k = 0; k < 3; ++k) o->set_texel16(0, i, k, (float)o->sums[i * 3 + k] / ((float)c * 255.0f));
      o->set_texel16(0, i, 3, 1.0f);
    }
    return;
  }
  for (size_t i = 0; i < n; ++i) {
    uint32_t c = o->counts[i];
    if (c == 0) { g[i * 4 + 0] = g[i * 4 + 1] = g[i * 4 + 2] = g[i * 4 + 3] = 0; continue; }
    for (int k = 0; k < 3; ++k) g[i * 4 + k] = (uint8_t)((o->sums[i * 3 + k] + (c >> 1)) / c);
    g[i * 4 + 3] = 255;  /* alpha written as 1.0, Voxelization.fs:88 */
  }
}

/* Extension with no reference counterpart (README.md:14 only claims it): for Bounces >= 3 each extra
 * bounce gathers indirect light at every occupied voxel.  A voxel has no stored normal, so the gather
 * is isotropic: one diffuse-aperture cone along each of +/-X, +/-Y, +/-Z, averaged, modulated by the
 * voxel's own stored colour and added to it.  Defined in DESIGN.md "Bounces". */
static void reinject_bounce(orc_ctx* o) {
  const orc_params& p = o->p;
  const int V = p.VoxelDimensions;
  const float vws = p.VoxelGridWorldSize / (float)V;
  std::vector<uint8_t> next(o->grid[0]);
  static const float AX[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
#pragma omp parallel for schedule(dynamic, 1)
  for (int z = 0; z < V; ++z)
    for (int y = 0; y < V; ++y)
      for (int x = 0; x < V; ++x) {
        size_t i = ((size_t)z * V + y) * V + x;
        if (o->texel(0, i, 3) == 0.0f) continue;
        V3 c = {((float)x + 0.5f) * vws - 0.5f * p.VoxelGridWorldSize,
                ((float)y + 0.5f) * vws - 0.5f * p.VoxelGridWorldSize,
                ((float)z + 0.5f) * vws - 0.5f * p.VoxelGridWorldSize};
        float acc[3] = {0, 0, 0};
        uint64_t dummy = 0;
        for (int a = 0; a < 6; ++a) {
          V3 d = {AX[a][0], AX[a][1], AX[a][2]};
          V3 start = add3(c, scale3(d, vws));
          float r[4];
          cone_trace(o, start, d, p.DiffuseTanHalfAngle, r, &dummy);
          for (int k = 0; k < 3; ++k) acc[k] += r[k] * (1.0f / 6.0f);
        }
        for (int k = 0; k < 3; ++k) {
          float base = o->texel(0, i, k);
          float v = std::min(base + acc[k] * base, 1.0f);
          if (o->p.GridFormat == 1) { uint16_t h = float_to_half(v); std::memcpy(&next[i * 8 + k * 2], &h, 2); }
          else next[i * 4 + k] = (uint8_t)std::lrintf(v * 255.0f);
        }
      }
  o->grid[0].swap(next);
}

extern "C" int orc_resolve_and_mip(orc_ctx* o) {
  ensure_grid(o);
  resolve_level0(o);
  orc_build_mips(o);
  for (int b = 3; b <= o->p.Bounces; ++b) {
    reinject_bounce(o);
    orc_build_mips(o);
  }
  return 0;
}

extern "C" int orc_draw_voxels(orc_ctx* o) {
  orc_draw_voxels_range(o, 0, o->idx.size() / 3, 1);
  return orc_resolve_and_mip(o);
}

/* ------------------------------------------------------------------ plumbing */

extern "C" void orc_default_params(orc_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->VoxelDimensions = 128;
  p->VoxelGridWorldSize = 150.0f;
  p->ShadowMapSize = 4096;
  p->screen_width = 1280;
  p->screen_height = 720;
  static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float* mats[] = {p->ModelMatrix, p->ModelViewMatrix, p->ProjectionMatrix, p->DepthModelViewProjectionMatrix,
                   p->ProjX, p->ProjY, p->ProjZ};
  for (float* m : mats) std::memcpy(m, I, sizeof(I));
  p->CameraPosition[1] = 4.0f;
  p->LightDirection[1] = 1.0f; p->LightDirection[2] = 0.25f;
  p->ambientFactor = 0.1f;
  p->NumDiffuseCones = 6;
  static const float dirs[18] = {0, 0, 1, 0, 0.866025f, 0.5f, 0.823639f, 0.267617f, 0.5f,
                                 0.509037f, -0.700629f, 0.5f, -0.509037f, -0.700629f, 0.5f,
                                 -0.823639f, 0.267617f, 0.5f};
  static const float wts[6] = {0.25f, 0.15f, 0.15f, 0.15f, 0.15f, 0.15f};
  std::memcpy(p->ConeDirections, dirs, sizeof(dirs));
  std::memcpy(p->ConeWeights, wts, sizeof(wts));
  p->DiffuseTanHalfAngle = 0.577f;
  p->SpecularTanHalfAngle = 0.07f;
  p->StepMultiplier = 1.0f;
  p->MaxDistance = 75.0f;
  p->MaxAlpha = 0.95f;
  p->PcfRadius = 2;
  p->ShadowBias = 0.002f;
  p->CoveragePolicy = 1;
  p->VoxelStoreMode = 0;
  p->Bounces = 2;
  p->GridFormat = 0;
  p->FilterMode = 1;   /* 8-bit fixed-point filter weights: what texture hardware and llvmpipe's RGBA8 path use */
}

extern "C" orc_ctx* orc_create(void) {
  orc_ctx* o = new orc_ctx();
  orc_default_params(&o->p);
  o->materials.resize(1);
  return o;
}
extern "C" void orc_destroy(orc_ctx* o) { delete o; }

extern "C" int orc_set_params(orc_ctx* o, const orc_params* p) {
  if (p->VoxelDimensions < 1 || (p->VoxelDimensions & (p->VoxelDimensions - 1))) return -1;
  if (p->NumDiffuseCones < 0 || p->NumDiffuseCones > ORC_MAX_CONES) return -1;
  bool regrid = p->VoxelDimensions != o->p.VoxelDimensions || p->GridFormat != o->p.GridFormat;
  bool reshadow = p->ShadowMapSize != o->p.ShadowMapSize;
  o->p = *p;
  if (regrid) { o->grid.clear(); o->gridV = 0; }
  if (reshadow) o->depth.clear();
  return 0;
}

extern "C" int orc_upload_texture(orc_ctx* o, int id, int w, int h, int channels, const uint8_t* pix) {
  if (id < 0 || w < 1 || h < 1 || (channels != 1 && channels != 3 && channels != 4)) return -1;
  if ((int)o->textures.size() <= id) o->textures.resize(id + 1);
  Texture t;
  t.w.push_back(w); t.h.push_back(h);
  std::vector<uint8_t> l0((size_t)w * h * 4);
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    /* Model.h:159-169: 1 -> GL_RED (r,0,0,1), 3 -> GL_RGB (r,g,b,1), 4 -> GL_RGBA */
    l0[i * 4 + 0] = pix[i * channels];
    l0[i * 4 + 1] = channels >= 3 ? pix[i * channels + 1] : 0;
    l0[i * 4 + 2] = channels >= 3 ? pix[i * channels + 2] : 0;
    l0[i * 4 + 3] = channels == 4 ? pix[i * channels + 3] : 255;
    if (l0[i * 4 + 3] != 255) t.has_alpha = true;
  }
  t.px.push_back(std::move(l0));
  build_texture_mips(t);
  o->textures[id] = std::move(t);
  return 0;
}

extern "C" int orc_set_material(orc_ctx* o, int mat, int d, int s, int h, float shininess) {
  if (mat < 0 || mat > 65535) return -1;
  if ((int)o->materials.size() <= mat) o->materials.resize(mat + 1);
  o->materials[mat].diffuse = d; o->materials[mat].specular = s; o->materials[mat].height = h;
  o->materials[mat].shininess = shininess;
  return 0;
}

extern "C" int orc_upload_mesh(orc_ctx* o, const float* v, size_t nv, const uint32_t* idx, size_t nt,
                               const uint16_t* tm) {
  for (size_t i = 0; i < nt * 3; ++i) if (idx[i] >= nv) return -1;
  o->verts.assign(v, v + nv * 14);
  o->idx.assign(idx, idx + nt * 3);
  if (tm) {
    o->tri_mat.assign(tm, tm + nt);
    uint16_t mx = 0;
    for (size_t i = 0; i < nt; ++i) mx = std::max(mx, tm[i]);
    if (o->materials.size() <= mx) o->materials.resize((size_t)mx + 1);
  } else o->tri_mat.clear();
  return 0;
}

extern "C" int orc_get_depth(orc_ctx* o, uint32_t* d) {
  if (o->depth.empty()) return -1;
  std::memcpy(d, o->depth.data(), o->depth.size() * 4);
  return 0;
}
extern "C" int orc_get_counts(orc_ctx* o, uint32_t* c) {
  ensure_grid(o);
  std::memcpy(c, o->counts.data(), o->counts.size() * 4);
  return 0;
}
extern "C" int orc_get_sums(orc_ctx* o, uint32_t* s) {
  ensure_grid(o);
  std::memcpy(s, o->sums.data(), o->sums.size() * 4);
  return 0;
}
extern "C" int orc_set_accum(orc_ctx* o, const uint32_t* c, const uint32_t* s) {
  ensure_grid(o);
  std::memcpy(o->counts.data(), c, o->counts.size() * 4);
  std::memcpy(o->sums.data(), s, o->sums.size() * 4);
  return 0;
}
extern "C" int orc_get_grid(orc_ctx* o, int level, uint8_t* rgba) {
  ensure_grid(o);
  if (level < 0 || level >= (int)o->grid.size()) return -1;
  std::memcpy(rgba, o->grid[level].data(), o->grid[level].size());
  return 0;
}
extern "C" int orc_set_grid_level0(orc_ctx* o, const uint8_t* rgba) {
  ensure_grid(o);
  std::memcpy(o->grid[0].data(), rgba, o->grid[0].size());
  return 0;
}
extern "C" int orc_get_visibility(orc_ctx* o, uint32_t* tri) {
  if (o->vis.empty()) return -1;
  for (size_t i = 0; i < o->vis.size(); ++i) tri[i] = o->vis[i] == ~0ull ? 0xFFFFFFFFu : (uint32_t)o->vis[i];
  return 0;
}
extern "C" int orc_get_frame(orc_ctx* o, uint8_t* rgba) {
  if (o->frame.empty()) return -1;
  std::memcpy(rgba, o->frame.data(), o->frame.size());
  return 0;
}
/* diagnostics: re-shade pixel (i,j) of the last render and return 8 x 11 cone records + 11 scalars */
extern "C" int orc_debug_pixel(orc_ctx* o, int i, int j, float* out110, uint8_t* rgba) {
  const int W = o->p.screen_width;
  if (o->vis.empty()) return -1;
  uint64_t key = o->vis[(size_t)j * W + i];
  if (key == ~0ull) return 1;
  for (int k = 0; k < 110; ++k) out110[k] = 0;
  g_debug = out110;
  uint64_t n = 0;
  shade_pixel(o, (size_t)(uint32_t)key, i, j, rgba, &n);
  g_debug = nullptr;
  return 0;
}

extern "C" uint64_t orc_cone_samples(orc_ctx* o) { return o->cone_samples; }
extern "C" uint64_t orc_fragment_count(orc_ctx* o) { return o->fragments; }

extern "C" int orc_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}

extern "C" void orc_sample_voxels(orc_ctx* o, const float pos[3], float lod, float out[4]) {
  ensure_grid(o);
  sample_voxels(o, {pos[0], pos[1], pos[2]}, lod, out);
}
extern "C" void orc_cone(orc_ctx* o, const float start[3], const float dir[3], float th, float out[4], int* n) {
  ensure_grid(o);
  uint64_t s = 0;
  cone_trace(o, {start[0], start[1], start[2]}, {dir[0], dir[1], dir[2]}, th, out, &s);
  if (n) *n = (int)s;
}
extern "C" int orc_select_axis(const float a[3], const float b[3], const float c[3]) {
  return select_axis({a[0], a[1], a[2]}, {b[0], b[1], b[2]}, {c[0], c[1], c[2]});
}
extern "C" void orc_sample_texture(orc_ctx* o, int tex, float u, float v, float lod, float out[4]) {
  const Texture* t = get_tex(o, tex);
  if (!t) { out[0] = out[1] = out[2] = out[3] = 1; return; }
  sample_tex(*t, u, v, lod, out);
}
extern "C" float orc_pcf(orc_ctx* o, const float dc[4]) {
  int taps = (2 * o->p.PcfRadius + 1) * (2 * o->p.PcfRadius + 1);
  return pcf_lit_taps(o, dc[0], dc[1], dc[2], dc[3]) / (float)taps;
}
