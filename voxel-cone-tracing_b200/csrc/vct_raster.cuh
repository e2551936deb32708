// vct_raster.cuh -- triangle-parallel rasterisation skeleton shared by the shadow, voxel-coverage and
// visibility passes.
//
// The GL pipeline the reference relies on (glDrawElements -> fixed-function rasteriser) is replaced by
// two kernels per pass:
//   raster_small : one thread per triangle.  Set-up; bounding boxes of <= SMALL_AREA pixels are
//                  rasterised in the thread; larger ones are cut into 8x8-pixel tiles and queued as
//                  work items of up to ITEM_TILES tiles.
//   raster_tiles : persistent warps; one warp per work item, one lane per pixel (2 steps per tile).
// A Pass supplies:  struct Setup;  bool setup_full(tri, Setup&, i0,i1,j0,j1)  (raster_small: from the vertex cache)
//                   bool setup(tri, Setup&, i0,i1,j0,j1)       (raster_tiles: may reload a stored record)
//                   bool tile_may_cover(Setup&, x0,y0,x1,y1)   (exact or conservative reject)
//                   unsigned small(Setup&, tri, i0,i1,j0,j1)    (thread-serial path)
//                   void pixel(Setup&, tri, i, j, bool in_bbox) (warp path, called by all 32 lanes)
//                   void tile_rows(int& ty0, int& ty1, int& step)  (which 8-pixel tile rows of [ty0, ty1] to walk; default all)
//                   static constexpr bool kAppends; if true also covered(), reserve(n), emit(tri,i,j,pos)
//                   static constexpr bool kWarpMedium: boxes up to MEDIUM_MAX^2 pixels are rasterised by the whole warp inside
//                   raster_small through pixel() (passes whose pixel() has no warp-collective operation)
#pragma once

#include "vct_internal.h"

namespace vct {

constexpr int SMALL_AREA = 16;
constexpr int MEDIUM_MAX = 16;
constexpr int TILE = 8;
constexpr int ITEM_TILES = 16;

// Every lane receives lane `src`'s copy of a POD (word by word; fully unrolled, stays in registers)
template <class T>
__device__ __forceinline__ T warp_broadcast(const T& v, int src) {
  static_assert(sizeof(T) % 4 == 0, "POD of 32-bit words");
  T out;
  const uint32_t* a = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* b = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(T) / 4); ++k) b[k] = __shfl_sync(0xffffffffu, a[k], src);
  return out;
}

template <class Pass>
__global__ void __launch_bounds__(128, 5) raster_small(Pass pass, uint32_t tri_begin, uint32_t tri_end,
                                                    TileItem* __restrict__ items, uint32_t items_cap,
                                                    Counters* __restrict__ ctr, uint32_t interleave = 1,
                                                    uint32_t phase = 0) {
  // interleave > 1: this launch owns every interleave-th block of 128 triangles of [tri_begin, tri_end), starting at
  // block `phase` (triangle sharding across ranks: contiguous ranges of a mesh differ a lot in fragments per triangle)
  uint32_t tri = tri_begin + (blockIdx.x * interleave + phase) * blockDim.x + threadIdx.x;
  bool live = tri < tri_end;
  typename Pass::Setup s;
  int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
  if (live) live = pass.setup_full(tri, s, i0, i1, j0, j1);   // may also write a per-triangle record for later stages
  int w = i1 - i0 + 1, h = j1 - j0 + 1;
  bool small_tri = live && (w * h <= SMALL_AREA);
  // every lane of the warp calls small(): passes that append to a queue aggregate across the warp
  pass.small(s, tri, small_tri, i0, i1, j0, j1);
  if constexpr (Pass::kWarpMedium) {
    // Medium triangles (bounding box up to MEDIUM_MAX pixels a side, too many pixels for one lane): the warp takes them
    // one at a time, the owner lane broadcasts its set-up and 32 lanes test an 8x4 block of the box per step.  No queue entry,
    // no second set-up in raster_tiles, no walk over the up to four 8x8 screen tiles such a box straddles -- meshes of
    // ~1-pixel triangles (BASELINE config 4) are almost entirely of this kind in the shadow and visibility passes.
    const bool medium = live && !small_tri && w <= MEDIUM_MAX && h <= MEDIUM_MAX;
    unsigned todo = __ballot_sync(0xffffffffu, medium);
    const int lane = (int)(threadIdx.x & 31);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const typename Pass::Setup bs = warp_broadcast(s, src);
      const int bi0 = __shfl_sync(0xffffffffu, i0, src), bi1 = __shfl_sync(0xffffffffu, i1, src);
      const int bj0 = __shfl_sync(0xffffffffu, j0, src), bj1 = __shfl_sync(0xffffffffu, j1, src);
      const uint32_t btri = __shfl_sync(0xffffffffu, tri, src);
      for (int by = bj0; by <= bj1; by += 4)            // warp-uniform loops over the 8x4 blocks of the box
        for (int bx = bi0; bx <= bi1; bx += 8) {
          const int i = bx + (lane & 7), j = by + (lane >> 3);
          pass.pixel(bs, btri, i, j, i <= bi1 && j <= bj1);
        }
    }
    if (medium) live = false;
  }
  if (live && !small_tri) {
    int tx0 = i0 / TILE, tx1 = i1 / TILE, ty0 = j0 / TILE, ty1 = j1 / TILE;
    int ty_step = 1;
    pass.tile_rows(ty0, ty1, ty_step);           // the tile rows this pass wants of [ty0, ty1]: first, last, stride
    uint32_t ntiles = ty0 <= ty1 ? (uint32_t)(tx1 - tx0 + 1) * (uint32_t)((ty1 - ty0) / ty_step + 1) : 0u;
    uint32_t nitems = (ntiles + ITEM_TILES - 1) / ITEM_TILES;
    uint32_t base = nitems ? atomicAdd(&ctr->n_items, nitems) : 0u;
    if (base + nitems > items_cap) {
      // Nothing of this triangle is queued and n_items now overstates what was written: raster_tiles must not
      // touch the queue at all (slots past the last complete write hold stale or uninitialised items).
      ctr->overflow = 1;
      ctr->items_overflow = 1;
    } else {
      for (uint32_t k = 0; k < nitems; ++k) {
        TileItem it;
        it.tri = tri;
        it.origin = k * ITEM_TILES;
        items[base + k] = it;
      }
    }
  }
}

template <class Pass>
__global__ void __launch_bounds__(256) raster_tiles(Pass pass, const TileItem* __restrict__ items,
                                                    uint32_t items_cap, Counters* __restrict__ ctr) {
  const uint32_t n_items = ctr->items_overflow ? 0u : min(ctr->n_items, items_cap);
  const uint32_t lane = threadIdx.x & 31;
  // dynamic distribution: item costs vary by orders of magnitude (slivers vs. screen-filling triangles), so
  // each warp takes the next item from a ticket counter instead of a static stride
  while (true) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(&ctr->next_item, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= n_items) break;
    TileItem item = items[it];
    typename Pass::Setup s;
    int i0, i1, j0, j1;
    if (!pass.setup(item.tri, s, i0, i1, j0, j1)) continue;   // warp-uniform
    int tx0 = i0 / TILE, tx1 = i1 / TILE, ty0 = j0 / TILE, ty1 = j1 / TILE;
    int ty_step = 1;
    pass.tile_rows(ty0, ty1, ty_step);
    if (ty0 > ty1) continue;
    uint32_t tw = (uint32_t)(tx1 - tx0 + 1);
    uint32_t ntiles = tw * (uint32_t)((ty1 - ty0) / ty_step + 1);
    uint32_t t_end = min(item.origin + (uint32_t)ITEM_TILES, ntiles);
    if constexpr (Pass::kAppends) {
      // Passes that append to a queue: evaluate the coverage of the item's (up to 16 tiles x 2 half-tiles =) 32 steps
      // ONCE, each lane remembering its own result as one bit per step; reserve the whole range with ONE atomic per item;
      // then replay the bits (one ballot per step instead of a second coverage test).  The fragment order inside the
      // item stays tile by tile.
      static_assert(ITEM_TILES * 2 <= 32, "one coverage bit per half-tile step");
      uint32_t total = 0, mine = 0;
      {
        uint32_t k = 0;
        for (uint32_t t = item.origin, tcol = item.origin % tw, trow = item.origin / tw; t < t_end;
             ++t, k += 2, trow += (tcol + 1 == tw), tcol = (tcol + 1 == tw) ? 0u : tcol + 1) {     // one division per item, not per tile
          int px0 = (tx0 + (int)tcol) * TILE, py0 = (ty0 + (int)trow * ty_step) * TILE;
          if (!pass.tile_may_cover(s, px0, py0, px0 + TILE, py0 + TILE)) continue;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            int i = px0 + (int)(lane & 7), j = py0 + half * 4 + (int)(lane >> 3);
            bool cov = i >= i0 && i <= i1 && j >= j0 && j <= j1 && pass.covered(s, i, j);
            mine |= (cov ? 1u : 0u) << (k + half);
            total += __popc(__ballot_sync(0xffffffffu, cov));
          }
        }
      }
      if (!total) continue;
      uint32_t base = 0;
      if (lane == 0) base = pass.reserve(total);
      base = __shfl_sync(0xffffffffu, base, 0);
      {
        uint32_t k = 0;
        for (uint32_t t = item.origin, tcol = item.origin % tw, trow = item.origin / tw; t < t_end;
             ++t, k += 2, trow += (tcol + 1 == tw), tcol = (tcol + 1 == tw) ? 0u : tcol + 1) {
          int px0 = (tx0 + (int)tcol) * TILE, py0 = (ty0 + (int)trow * ty_step) * TILE;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const bool cov = (mine >> (k + half)) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, cov);
            if (cov) pass.emit(item.tri, px0 + (int)(lane & 7), py0 + half * 4 + (int)(lane >> 3), base + __popc(m & ((1u << lane) - 1)));
            base += __popc(m);
          }
        }
      }
    } else {
      for (uint32_t t = item.origin, tcol = item.origin % tw, trow = item.origin / tw; t < t_end;
           ++t, trow += (tcol + 1 == tw), tcol = (tcol + 1 == tw) ? 0u : tcol + 1) {
        int tx = tx0 + (int)tcol, ty = ty0 + (int)trow * ty_step;
        int px0 = tx * TILE, py0 = ty * TILE;
        if (!pass.tile_may_cover(s, px0, py0, px0 + TILE, py0 + TILE)) continue;  // warp-uniform
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          int i = px0 + (int)(lane & 7), j = py0 + half * 4 + (int)(lane >> 3);
          bool in_bbox = i >= i0 && i <= i1 && j >= j0 && j <= j1;
          pass.pixel(s, item.tri, i, j, in_bbox);
        }
      }
    }
  }
}

// Exact tile reject for the integer passes: the samples of every pixel in the tile lie inside the
// closed box [x0*256, x1*256] x [y0*256, y1*256]; an edge whose maximum over the box (plus its
// fill-rule bias) is negative cannot be satisfied.
__device__ __forceinline__ bool tile_may_cover_exact(const RasterTri& t, int x0, int y0, int x1, int y1) {
  auto emax_box = [&](int ax, int ay, int bx, int by) {
    int dx = bx - ax, dy = by - ay;
    int px = (-dy > 0) ? x1 * SUBPIX : x0 * SUBPIX;
    int py = (dx > 0) ? y1 * SUBPIX : y0 * SUBPIX;
    return RasterTri::ev(ax, ay, bx, by, px, py);
  };
  return emax_box(t.X0, t.Y0, t.X1, t.Y1) >= 0 && emax_box(t.X1, t.Y1, t.X2, t.Y2) >= 0 &&
         emax_box(t.X2, t.Y2, t.X0, t.Y0) >= 0;
}

}  // namespace vct
