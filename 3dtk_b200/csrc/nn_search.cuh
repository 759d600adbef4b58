// nn_search.cuh -- exact nearest neighbour of a query in a scan's uniform grid.
//
// Replaces the per-query work of KDtree::FindClosest / KDTreeImpl::_FindClosest
// (reference src/slam6d/kd.cc:78-87, include/slam6d/kdTreeImpl.h:345-383): closest model point with
// squared distance STRICTLY below maxdist2, else none.
//
// Two stages, both inside the calling kernel (one launch):
//   stage 1  one thread per query: the 3x3x3 cell stencil, read as 9 contiguous x-runs of the
//            cell-sorted fp32x4 point array.  Exact whenever the hit is closer than one cell edge.
//   stage 2  one warp per still-open query: rows (dy,dz) are visited ring by ring, one row per lane,
//            each row clipped to the x-extent of the current search sphere; after every 32 rows a
//            warp-shuffle arg-min merges the lanes and the loop stops as soon as the best distance
//            is covered by the completed rings.
// Precision: candidates are screened in fp32 on origin-relative coordinates against a bound that is
// provably above the fp64 distance of any candidate able to beat the current best; survivors are
// re-evaluated in fp64 with the reference's Dist2 rounding (no FMA), so EXACT=true returns the fp64
// arg-min.  EXACT=false takes all decisions in fp32 (fast mode of the fused match).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr unsigned kNoIdx = 0xFFFFFFFFu;

struct Best {
  double d2;      // exact (EXACT) or fp32 (fast) squared distance of the current best, init maxdist2
  float thr;      // fp32 screening bound
  int j;          // sorted position of the best point, -1 = none
  unsigned oidx;  // its original row
};

struct SearchSmem {
  double sx[kBlock], sy[kBlock], sz[kBlock];
  double bd2[kBlock];
  int bj[kBlock];
  unsigned boidx[kBlock];
  int list[kBlock];
  int warp_cnt[2][kWarps];  // double-buffered by call parity (a fast warp may enter the next tile)
};

template <bool EXACT>
__device__ __forceinline__ float filter_bound(double b, float e) {
  if (!EXACT) return __double2float_rn(b);
  // |d2_fp32 - d2_exact| <= 2*sqrt(3)*e*d + 3e^2 + 3u*d2 with e the per-axis difference error;
  // the bound below dominates it.
  const double ed = (double)e;
  double t = (b + ed * (4.0 * sqrt(b) + 4.0 * ed)) * 1.000001;
  return __double2float_ru(t);
}

__device__ __forceinline__ int cell_coord(double v, double g0, double inv_h) {
  double f = floor((v - g0) * inv_h);
  f = fmin(fmax(f, -1.0e9), 1.0e9);
  return (int)f;
}

template <bool EXACT>
__device__ __forceinline__ void scan_range(const GridDev& g, unsigned beg, unsigned end, float qx,
                                           float qy, float qz, double sx, double sy, double sz,
                                           float e, Best& b) {
#pragma unroll 2
  for (unsigned j = beg; j < end; ++j) {
    const float4 p = __ldg(g.p32 + j);
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    if (d2 < b.thr) {
      const unsigned oi = __float_as_uint(p.w);
      if (EXACT) {
        const double2 pa = __ldg(reinterpret_cast<const double2*>(g.p64 + j));
        const double pz = __ldg(reinterpret_cast<const double*>(g.p64 + j) + 2);
        // Dist2(query, point), globals.icc:237-245: (x2-x1)^2 summed left to right, no contraction
        const double ex = __dsub_rn(pa.x, sx), ey = __dsub_rn(pa.y, sy), ez = __dsub_rn(pz, sz);
        const double d2e =
            __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
        if (d2e < b.d2 || (d2e == b.d2 && b.j >= 0 && oi < b.oidx)) {
          b.d2 = d2e;
          b.j = (int)j;
          b.oidx = oi;
          b.thr = filter_bound<true>(d2e, e);
        }
      } else {
        b.d2 = (double)d2;
        b.thr = d2;
        b.j = (int)j;
        b.oidx = oi;
      }
    }
  }
}

// stage 1: 27-cell stencil as 9 x-runs
template <bool EXACT>
__device__ __forceinline__ void stencil_search(const GridDev& g, int cx, int cy, int cz, float qx,
                                               float qy, float qz, double sx, double sy, double sz,
                                               float e, Best& b) {
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
  if (x0 > x1) return;
  unsigned rb[9], re[9];
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int y = cy + (r % 3) - 1, z = cz + (r / 3) - 1;
    if ((unsigned)y < (unsigned)g.ny && (unsigned)z < (unsigned)g.nz) {
      const size_t row = ((size_t)z * g.ny + y) * g.nx;
      rb[r] = __ldg(g.cell_start + row + x0);
      re[r] = __ldg(g.cell_start + row + x1 + 1);
    } else {
      rb[r] = re[r] = 0;
    }
  }
#pragma unroll
  for (int r = 0; r < 9; ++r) scan_range<EXACT>(g, rb[r], re[r], qx, qy, qz, sx, sy, sz, e, b);
}

__device__ __forceinline__ void best_merge(Best& b, double od2, int oj, unsigned ooidx) {
  if (od2 < b.d2 || (od2 == b.d2 && ooidx < b.oidx)) {
    b.d2 = od2;
    b.j = oj;
    b.oidx = ooidx;
  }
}

// ring index r of flattened row f: f == 0 -> 0, else (2r-1)^2 <= f < (2r+1)^2
__device__ __forceinline__ int ring_of(long long f) {
  int r = (int)ceilf((sqrtf((float)(f + 1)) - 1.0f) * 0.5f);
  if (r < 0) r = 0;
  while (r > 0 && (2LL * r - 1) * (2LL * r - 1) > f) --r;
  while ((2LL * r + 1) * (2LL * r + 1) <= f) ++r;
  return r;
}

// stage 2: whole warp works on one query.  `b` enters identical on all lanes and leaves identical.
template <bool EXACT>
__device__ __forceinline__ void ring_search_warp(const GridDev& g, int cx, int cy, int cz, float qx,
                                                 float qy, float qz, double sx, double sy,
                                                 double sz, float e, Best& b) {
  const int lane = threadIdx.x & 31;
  const double fy = (sy - g.g0[1]) - (double)cy * g.h;  // offset of the query inside its cell
  const double fz = (sz - g.g0[2]) - (double)cz * g.h;
  const long long kgrid =
      max(max((long long)cy, (long long)g.ny - 1 - cy), max((long long)cz, (long long)g.nz - 1 - cz));
  for (long long base = 0;; base += 32) {
    const double R = sqrt(b.d2);
    long long k = (long long)fmin(ceil(R * g.inv_h), 1.0e9);
    if (k > kgrid) k = kgrid;
    if (k < 0) k = 0;
    const long long total = (2 * k + 1) * (2 * k + 1);
    if (base >= total) break;
    const long long f = base + lane;
    if (f < total) {
      const int r = ring_of(f);
      int dy = 0, dz = 0;
      if (r > 0) {
        const long long eidx = f - (2LL * r - 1) * (2LL * r - 1);
        const int side = (int)(eidx / (2 * r)), off = (int)(eidx % (2 * r));
        if (side == 0) { dy = -r + off; dz = -r; }
        else if (side == 1) { dy = r; dz = -r + off; }
        else if (side == 2) { dy = r - off; dz = r; }
        else { dy = -r; dz = r - off; }
      }
      const int y = cy + dy, z = cz + dz;
      if ((unsigned)y < (unsigned)g.ny && (unsigned)z < (unsigned)g.nz) {
        // exact lower bound of the (y,z)-distance between the query and any point of this row
        const double ddy = dy > 0 ? (double)dy * g.h - fy : (dy < 0 ? fy + (double)(-dy - 1) * g.h : 0.0);
        const double ddz = dz > 0 ? (double)dz * g.h - fz : (dz < 0 ? fz + (double)(-dz - 1) * g.h : 0.0);
        const double dyz2 = (ddy > 0 ? ddy * ddy : 0.0) + (ddz > 0 ? ddz * ddz : 0.0);
        if (dyz2 * (1.0 - 1e-9) < b.d2) {
          const double w = sqrt(fmax(b.d2 - dyz2 * (1.0 - 1e-9), 0.0)) * (1.0 + 1e-9) + 1e-300;
          const int x0 = max(cell_coord(sx - w, g.g0[0], g.inv_h), 0);
          const int x1 = min(cell_coord(sx + w, g.g0[0], g.inv_h), g.nx - 1);
          if (x0 <= x1) {
            const size_t row = ((size_t)z * g.ny + y) * g.nx;
            const unsigned beg = __ldg(g.cell_start + row + x0);
            const unsigned end = __ldg(g.cell_start + row + x1 + 1);
            scan_range<EXACT>(g, beg, end, qx, qy, qz, sx, sy, sz, e, b);
          }
        }
      }
    }
    // warp arg-min (distance, then original row)
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const double od2 = __shfl_xor_sync(0xffffffffu, b.d2, m);
      const int oj = __shfl_xor_sync(0xffffffffu, b.j, m);
      const unsigned oo = __shfl_xor_sync(0xffffffffu, b.oidx, m);
      best_merge(b, od2, oj, oo);
    }
    b.thr = filter_bound<EXACT>(b.d2, e);
    if (b.j >= 0) {
      // rings 0..rc are complete after base+32 rows
      const long long rows_done = base + 32;
      long long rc = (long long)floorf((sqrtf((float)rows_done) - 1.0f) * 0.5f);
      if (rc < 0) rc = 0;
      while ((2 * (rc + 1) + 1) * (2 * (rc + 1) + 1) <= rows_done) ++rc;
      while (rc > 0 && (2 * rc + 1) * (2 * rc + 1) > rows_done) --rc;
      const double lim = (double)rc * g.h;
      if (b.d2 <= lim * lim * (1.0 - 1e-9)) break;
    }
  }
}

// Block-wide search: every thread of a kBlock-thread block calls this once per tile (it contains
// __syncthreads).  `active` threads carry a query s (in the grid's frame).
template <bool EXACT>
__device__ __forceinline__ void nn_block_search(const GridDev& g, bool active, double sx, double sy,
                                                double sz, double maxdist2, SearchSmem& sm,
                                                int parity, int& out_j, double& out_d2,
                                                unsigned& out_oidx, unsigned& stage2_count) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Best b;
  b.d2 = maxdist2;
  b.j = -1;
  b.oidx = kNoIdx;
  b.thr = 0.f;
  bool need2 = false;
  if (active) {
    const double ox = fmax(fmax(g.bbox_lo[0] - sx, sx - g.bbox_hi[0]), 0.0);
    const double oy = fmax(fmax(g.bbox_lo[1] - sy, sy - g.bbox_hi[1]), 0.0);
    const double oz = fmax(fmax(g.bbox_lo[2] - sz, sz - g.bbox_hi[2]), 0.0);
    const double dbox2 = ox * ox + oy * oy + oz * oz;
    if (dbox2 < maxdist2) {  // false for NaN queries as well
      const double rx = sx - g.c[0], ry = sy - g.c[1], rz = sz - g.c[2];
      const float qx = (float)rx, qy = (float)ry, qz = (float)rz;
      const float qm = fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz));
      const float e = 1.25e-7f * (qm + g.bmax) + 1e-30f;  // > 2.02 * 2^-24 * (|q|+|p|)
      b.thr = filter_bound<EXACT>(maxdist2, e);
      const int cx = cell_coord(sx, g.g0[0], g.inv_h);
      const int cy = cell_coord(sy, g.g0[1], g.inv_h);
      const int cz = cell_coord(sz, g.g0[2], g.inv_h);
      stencil_search<EXACT>(g, cx, cy, cz, qx, qy, qz, sx, sy, sz, e, b);
      const double rg2 = g.h * g.h * (1.0 - 1e-9);
      need2 = !((b.j >= 0 && b.d2 <= rg2) || (maxdist2 <= rg2));
    }
  }
  const unsigned ball = __ballot_sync(0xffffffffu, need2);
  if (lane == 0) sm.warp_cnt[parity & 1][warp] = __popc(ball);
  __syncthreads();
  int offset = 0, nlist = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const int c = sm.warp_cnt[parity & 1][w];
    if (w < warp) offset += c;
    nlist += c;
  }
  if (nlist > 0) {  // block-uniform
    if (need2) {
      const int slot = offset + __popc(ball & ((1u << lane) - 1u));
      sm.list[slot] = tid;
      sm.sx[tid] = sx; sm.sy[tid] = sy; sm.sz[tid] = sz;
      sm.bd2[tid] = b.d2; sm.bj[tid] = b.j; sm.boidx[tid] = b.oidx;
    }
    __syncthreads();
    for (int li = warp; li < nlist; li += kWarps) {
      const int t = sm.list[li];
      const double qsx = sm.sx[t], qsy = sm.sy[t], qsz = sm.sz[t];
      Best wb;
      wb.d2 = sm.bd2[t]; wb.j = sm.bj[t]; wb.oidx = sm.boidx[t];
      const float qx = (float)(qsx - g.c[0]), qy = (float)(qsy - g.c[1]), qz = (float)(qsz - g.c[2]);
      const float qm = fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz));
      const float e = 1.25e-7f * (qm + g.bmax) + 1e-30f;
      wb.thr = filter_bound<EXACT>(wb.d2, e);
      const int cx = cell_coord(qsx, g.g0[0], g.inv_h);
      const int cy = cell_coord(qsy, g.g0[1], g.inv_h);
      const int cz = cell_coord(qsz, g.g0[2], g.inv_h);
      ring_search_warp<EXACT>(g, cx, cy, cz, qx, qy, qz, qsx, qsy, qsz, e, wb);
      if (lane == 0) { sm.bd2[t] = wb.d2; sm.bj[t] = wb.j; sm.boidx[t] = wb.oidx; }
    }
    __syncthreads();
    if (need2) { b.d2 = sm.bd2[tid]; b.j = sm.bj[tid]; b.oidx = sm.boidx[tid]; }
    if (tid == 0) stage2_count += (unsigned)nlist;
  }
  out_j = b.j;
  out_d2 = b.d2;
  out_oidx = b.oidx;
}

}  // namespace b200
