// nn_search.cuh -- exact nearest neighbour of a query in a scan's uniform grid.
//
// Replaces the per-query work of KDtree::FindClosest / KDTreeImpl::_FindClosest
// (reference src/slam6d/kd.cc:78-87, include/slam6d/kdTreeImpl.h:345-383): closest model point with
// squared distance STRICTLY below maxdist2, else none.
//
// Structure (everything is per thread or per warp -- no block barrier in the query loop):
//   seed     the neighbour found in the previous ICP iteration gives an exact upper bound r on the
//            NN distance (the idea of the reference's "cached k-d tree", doc/papers/3dim2007.pdf,
//            applied to a grid); no seed -> r = maxdist.
//   stage 1  one thread per query: the cells of the 3x3x3 stencil that the ball of radius r touches,
//            read as contiguous x-runs of the cell-sorted fp32x4 point array, in ONE flattened loop
//            (a per-thread range table in shared memory keeps the lanes of a warp busy).
//            Exact whenever the hit is closer than one cell edge.
//   stage 1b one thread per query, for balls wider than a cell (early ICP iterations, no seed): the rows
//            of cells the ball touches are enumerated near-to-far, eight at a time (their ranges are
//            fetched with independent loads), scanned in the same flattened loop, and the ball shrinks
//            as soon as a candidate is found.  Covers up to kBallRings rings of cells.
//   stage 2  (rare: radius beyond kBallRings cells, or an fp32-ambiguous winner) the query is finished by
//            its own warp: rows (dy,dz) are visited ring by ring, one row per lane, clipped to the
//            x-extent of the current search sphere; after every 32 rows a warp-shuffle arg-min merges
//            the lanes and the loop stops as soon as the best distance is covered by the finished rings.
//   skip     every full search also certifies a MOTION BUDGET: it scans a ball 2*delta wider than needed and
//            records half the gap between the winner and the nearest thing that could replace it (the
//            runner-up, or the edge of the scanned ball).  While the query has moved less than that since
//            the search, its neighbour provably cannot have changed (triangle inequality, strict), so the
//            caller skips the search and only re-evaluates the exact distance to the cached neighbour.
//            In the converged phase of ICP almost every query is skipped.
// Precision: the inner loops run in fp32 on origin-relative coordinates and track the best and the
// second-best distance.  If the two are separated by more than the provable fp32 error the fp32 winner
// IS the fp64 arg-min and only its distance is re-evaluated in fp64 (reference Dist2 rounding, no FMA);
// otherwise (rare) the same cells are rescanned with every contender evaluated in fp64.  EXACT=false
// skips the fp64 work and takes all decisions in fp32 (fast mode of the fused match).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr unsigned kNoIdx = 0xFFFFFFFFu;
constexpr int kMaxRows = 9;
#ifndef B200_SCAN_BATCH
#define B200_SCAN_BATCH 3   // (2: 7.20, 3: 7.12 ms per match on the final kernels; 4 spills)
#endif
constexpr int kScanBatch = B200_SCAN_BATCH;   // candidate loads kept in flight per lane
constexpr int kBallRings = 8;
// (tried: prefetch.global.L1 of every range's cache lines once the ranges are known, so that the walk below hits --
//  slower in every phase, 8.66 vs 8.24 ms per 1M/1M match: the kernel is issue-bound enough that the extra
//  instructions cost more than the misses they hide)

struct Best {
  double d2;      // exact (EXACT) or fp32 (fast) squared distance of the current best, init maxdist2
  float thr;      // fp32 screening bound (exact rescans only)
  int j;          // sorted position of the best point, -1 = none
  unsigned oidx;  // its original row
};

struct Cand {     // fp32 tracking state of the inner loops
  float d1;       // smallest fp32 distance seen
  float d2nd;     // second smallest
  int j1;         // position of the smallest
};

// thread-private columns: rng[k][tid].  The 32 columns of a warp double as that warp's candidate TILE (see
// "tile" in the header comment): row k of the warp's slice holds 8 float4 points (128 B), so the tile never
// touches another warp's columns and no block barrier is needed between the two uses.
#ifndef B200_RNG_ROWS
#define B200_RNG_ROWS 18
#endif
constexpr int kRngRows = B200_RNG_ROWS > 2 * kMaxRows ? B200_RNG_ROWS : 2 * kMaxRows;
// Off by default: measured slower on the bench pair (profiles/r01_tile_search_experiment.md) -- the union of 32
// neighbouring stencils is ~270 points, so evaluating all of it from shared memory costs as many instructions
// as the divergent per-lane walks it replaces, and the staging adds a serial chain of shuffles and loads.
#ifndef B200_TILE
#define B200_TILE 0
#endif
#ifndef B200_TILE_MIN_LANES
#define B200_TILE_MIN_LANES 12
#endif
template <int ROWS>
struct SearchSmemT {
  static constexpr int kTileCap = ROWS * 8;    // candidate points a warp can stage
  unsigned rng[ROWS][kBlock];
};
using SearchSmem = SearchSmemT<kRngRows>;            // iteration kernels (dynamic shared memory)
using SearchSmemSmall = SearchSmemT<2 * kMaxRows>;   // kernels with a static table (API batch search, LUM link)

// optional counters of the tile path (build with -DB200_TILE_STATS; read with b200icp_debug_tile_stats)
__device__ unsigned long long g_tile_stats[8];
__device__ __forceinline__ void tile_stat(int k, unsigned long long v) {
#ifdef B200_TILE_STATS
  if ((threadIdx.x & 31) == 0) atomicAdd(&g_tile_stats[k], v);
#endif
}

template <class SM>
__device__ __forceinline__ float4* tile_row(SM& sm, int k) {
  return reinterpret_cast<float4*>(&sm.rng[k][threadIdx.x & ~31]);
}
template <class SM>
__device__ __forceinline__ float4* tile_slot(SM& sm, int s) { return tile_row(sm, s >> 3) + (s & 7); }

template <bool EXACT>
__device__ __forceinline__ float filter_bound(double b, float e) {
  if (!EXACT) return __double2float_rn(b);
  // |d2_fp32 - d2_exact| <= 2*sqrt(3)*e*d + 3e^2 + 3u*d2 with e the per-axis difference error;
  // the bound below dominates it.
  const double ed = (double)e;
  double t = (b + ed * (4.0 * sqrt(b) + 4.0 * ed)) * 1.000001;
  return __double2float_ru(t);
}

// fp32 version of the same error bound: tol(d) >= |d2_fp32 - d2_exact| for a candidate at fp32 distance d
__device__ __forceinline__ float fp32_tol(float d, float e) {
  return (e * (4.0f * sqrtf(d) + 4.0f * e)) * 1.01f + 2e-6f * d + 1e-37f;
}

__device__ __forceinline__ float query_err(float qx, float qy, float qz, float bmax) {
  const float qm = fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz));
  return 1.25e-7f * (qm + bmax) + 1e-30f;  // > 2.02 * 2^-24 * (|q|+|p|)
}

__device__ __forceinline__ int cell_coord(double v, double g0, double inv_h) {
  double f = floor((v - g0) * inv_h);
  f = fmin(fmax(f, -1.0e9), 1.0e9);
  return (int)f;
}

// Dist2(query, point), globals.icc:237-245: (x2-x1)^2 summed left to right, no contraction
__device__ __forceinline__ double exact_d2(const GridDev& g, int j, double sx, double sy, double sz) {
  const double2 pa = __ldg(reinterpret_cast<const double2*>(g.p64 + j));
  const double pz = __ldg(reinterpret_cast<const double*>(g.p64 + j) + 2);
  const double ex = __dsub_rn(pa.x, sx), ey = __dsub_rn(pa.y, sy), ez = __dsub_rn(pz, sz);
  return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
}

__device__ __forceinline__ float dist32(const float4 p, float qx, float qy, float qz) {
  const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
  return fmaf(dx, dx, fmaf(dy, dy, dz * dz));
}

__device__ __forceinline__ void cand_update(Cand& c, float d, int j) {
  c.d2nd = fminf(c.d2nd, fmaxf(d, c.d1));
  if (d < c.d1) { c.d1 = d; c.j1 = j; }
}

// exact (slow) scan of one range: every candidate that passes the fp32 screen is evaluated in fp64
__device__ __forceinline__ void scan_range_exact(const GridDev& g, unsigned beg, unsigned end, float qx,
                                                 float qy, float qz, double sx, double sy, double sz,
                                                 float e, Best& b) {
  for (unsigned j = beg; j < end; ++j) {
    const float4 p = __ldg(g.p32 + j);
    const float d2 = dist32(p, qx, qy, qz);
    if (d2 < b.thr) {
      const unsigned oi = __float_as_uint(p.w);
      const double d2e = exact_d2(g, (int)j, sx, sy, sz);
      if (d2e < b.d2 || (d2e == b.d2 && b.j >= 0 && oi < b.oidx)) {
        b.d2 = d2e;
        b.j = (int)j;
        b.oidx = oi;
        b.thr = filter_bound<true>(d2e, e);
      }
    }
  }
}

__device__ __forceinline__ void best_merge(Best& b, double od2, int oj, unsigned ooidx) {
  if (od2 < b.d2 || (od2 == b.d2 && ooidx < b.oidx)) {
    b.d2 = od2;
    b.j = oj;
    b.oidx = ooidx;
  }
}

// ---- stage 1 -------------------------------------------------------------------------------------
// flattened fp32 scan of the first `nrows` ranges of this thread's table.  Candidates are taken four at
// a time: their indices come from the (cheap) range walk, then four independent 16-byte loads are in
// flight before the first distance is needed -- the loop is bound by load latency, not arithmetic.
template <int SB = kScanBatch, class SM>
__device__ __forceinline__ void scan_rows(const GridDev& g, const SM& sm, int nrows, float qx,
                                          float qy, float qz, Cand& c) {
  const int tid = threadIdx.x;
  int r = 0;
  unsigned j = 0, end = 0;
  for (;;) {
    int idx[SB];
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      while (j >= end && r < nrows) {
        j = sm.rng[2 * r][tid];
        end = sm.rng[2 * r + 1][tid];
        ++r;
      }
      idx[u] = j < end ? (int)j++ : -1;
    }
    if (idx[0] < 0) break;
    float4 p[SB];
#pragma unroll
    for (int u = 0; u < SB; ++u)
      if (idx[u] >= 0) p[u] = __ldg(g.p32 + idx[u]);
#pragma unroll
    for (int u = 0; u < SB; ++u)
      if (idx[u] >= 0 && idx[u] != c.j1)   // a point met twice must not become its own runner-up
        cand_update(c, dist32(p[u], qx, qy, qz), idx[u]);
  }
}

// Turns the fp32 tracking state into the exact result.  Returns false when the winner is ambiguous in
// fp32 (caller must settle it with an exact rescan).
template <bool EXACT>
__device__ __forceinline__ bool finalize_cand(const GridDev& g, const Cand& c, float e, double sx, double sy,
                                              double sz, Best& b) {
  if (c.j1 < 0) return true;  // nothing in reach (b keeps its seed / {maxdist2,-1})
  if (!EXACT) {
    if (c.j1 != b.j && (double)c.d1 < b.d2) {
      b.d2 = (double)c.d1; b.j = c.j1; b.oidx = __float_as_uint(__ldg(g.p32 + c.j1).w);
    }
    return true;
  }
  if (c.d2nd <= c.d1 + 2.5f * fp32_tol(c.d1, e)) return false;
  if (c.j1 != b.j) {  // a new winner: it is the fp64 arg-min of everything scanned plus the seed
    const double d2e = exact_d2(g, c.j1, sx, sy, sz);
    if (d2e < b.d2) { b.d2 = d2e; b.j = c.j1; b.oidx = __float_as_uint(__ldg(g.p32 + c.j1).w); }
  }
  return true;
}

__device__ __forceinline__ void cand_seed(const GridDev& g, const Best& b, float qx, float qy, float qz, Cand& c) {
  c.d1 = 3.0e38f; c.d2nd = 3.0e38f; c.j1 = -1;
  // the seed point itself, measured the same way as every candidate
  if (b.j >= 0) { c.d1 = dist32(__ldg(g.p32 + b.j), qx, qy, qz); c.j1 = b.j; }
}

// The part of the 3x3x3 stencil that the ball of squared radius r2 (already inflated for fp32 rounding)
// touches, as up to nine x-runs [va[r], vb[r]) of the cell-sorted point array (va == vb: row not needed or
// empty; all range loads are issued together).  rc2 = squared distance from the query to the nearest cell
// that is NOT covered (at most h^2: the edge of the stencil) -- everything closer than that is covered, which
// certifies a radius that is usually much larger than the ball that was asked for.
__device__ __forceinline__ void stencil_ranges(const GridDev& g, int cx, int cy, int cz, float fx, float fy,
                                               float fz, float r2, unsigned (&va)[9], unsigned (&vb)[9],
                                               float& rc2) {
  const float h = (float)g.h;
  const float lo_y = fy * fy, hi_y = (h - fy) * (h - fy);
  const float lo_z = fz * fz, hi_z = (h - fz) * (h - fz);
  const float lo_x = fx * fx, hi_x = (h - fx) * (h - fx);
  rc2 = h * h;
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int dy = (r % 3) - 1, dz = (r / 3) - 1;
    const int y = cy + dy, z = cz + dz;
    const float dyz2 = (dy < 0 ? lo_y : (dy > 0 ? hi_y : 0.f)) + (dz < 0 ? lo_z : (dz > 0 ? hi_z : 0.f));
    const float rem = r2 - dyz2;
    const bool left = lo_x <= rem, right = hi_x <= rem;
    va[r] = vb[r] = 0;
    if (rem < 0.f) { rc2 = fminf(rc2, dyz2); continue; }          // whole row left out
    if (!left) rc2 = fminf(rc2, dyz2 + lo_x);                      // its -x / +x ends left out
    if (!right) rc2 = fminf(rc2, dyz2 + hi_x);
    const int x0 = max(cx - (left ? 1 : 0), 0);
    const int x1 = min(cx + (right ? 1 : 0), g.nx - 1);
    if ((unsigned)y < (unsigned)g.ny && (unsigned)z < (unsigned)g.nz && x0 <= x1) {
      const size_t row = ((size_t)z * g.ny + y) * g.nx;
      va[r] = __ldg(g.cell_start + row + x0);
      vb[r] = __ldg(g.cell_start + row + x1 + 1);
    }
  }
}

// fp32 scan of those runs by the query's own thread.  Leaves the non-empty ranges in the thread's table
// (`nrows` of them) and the tracking state in `c` (seeded with b's point).
// (tried for the leftover batches: prefetch.global.L1 of the runs' cache lines as soon as the runs are known --
//  7.65 vs 7.60 ms per match, slower)
template <int SB = kScanBatch, class SM>
__device__ __forceinline__ int stencil_scan(const GridDev& g, SM& sm, int cx, int cy, int cz,
                                            float fx, float fy, float fz, float qx, float qy, float qz,
                                            float r2, const Best& b, Cand& c, float& rc2) {
  const int tid = threadIdx.x;
  unsigned va[9], vb[9];
  stencil_ranges(g, cx, cy, cz, fx, fy, fz, r2, va, vb, rc2);
  int nrows = 0;
#pragma unroll
  for (int r = 0; r < 9; ++r)
    if (va[r] < vb[r]) {
      sm.rng[2 * nrows][tid] = va[r];
      sm.rng[2 * nrows + 1][tid] = vb[r];
      ++nrows;
    }
  cand_seed(g, b, qx, qy, qz, c);
  scan_rows<SB>(g, sm, nrows, qx, qy, qz, c);
  return nrows;
}

// ---- tile (warp-cooperative form of stage 1) --------------------------------------------------------
// When the 32 queries of a batch are neighbours in the data scan's cell order, their stencils overlap almost
// completely: each model point is wanted by ~10 lanes.  Instead of every lane walking its own runs through
// L1/L2, the warp stages the UNION of the runs once in its slice of shared memory (coalesced-ish 16-byte
// loads, each point fetched once) and every lane then evaluates every staged point from shared memory with
// broadcast reads: no divergence, no dependent global loads in the loop.  A lane sees candidates outside its own
// runs as well; they lie outside its ball, so they can only tighten the runner-up bound (conservative) or
// raise an fp32 ambiguity that the exact rescan of the lane's own runs settles -- the result is unchanged.
//
// Union: per stencil slot r, lane l stages [max(a_l, M_l), b_l) where M_l = max b over the earlier lanes,
// provided a_l >= a_k of the lane k that set M_l (then [a_l, M_l) lies inside k's run, which is staged by
// induction); otherwise it stages its whole run.  A point staged twice (irregular order, or the same model
// row reached through different slots by lanes of different cell rows) is harmless: a repeat of the current
// best is skipped by index, a repeat of anything else leaves (d1, d2nd) unchanged.
// Returns false (nothing consumed) when the union does not fit the tile.
__device__ __forceinline__ unsigned long long warp_excl_max_u64(unsigned long long v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d && o > v) v = o;
  }
  const unsigned long long e = __shfl_up_sync(0xffffffffu, v, 1);
  return lane == 0 ? 0ull : e;
}

template <class SM>
__device__ __forceinline__ bool tile_stage(const GridDev& g, SM& sm, const unsigned (&va)[9],
                                           const unsigned (&vb)[9], int& T_out) {
  const int lane = threadIdx.x & 31;
  int base = 0;
  __syncwarp();   // the lanes' table columns are about to be overwritten by other lanes
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const bool has = va[r] < vb[r];
    const unsigned a = has ? va[r] : 0xFFFFFFFFu, b = has ? vb[r] : 0u;
    const unsigned long long pk = warp_excl_max_u64(((unsigned long long)b << 32) | (unsigned)(~a), lane);
    const unsigned M = (unsigned)(pk >> 32), aK = ~(unsigned)pk;
    unsigned from = a;
    if (has && M != 0u && a >= aK) from = max(a, M);
    const int n = has && b > from ? (int)(b - from) : 0;
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (base + total > SM::kTileCap) { __syncwarp(); return false; }   // warp-uniform
    int off = base + incl - n;
    for (unsigned j = from; j < from + (unsigned)n; ++j, ++off) {
      float4 p = __ldg(g.p32 + j);
      p.w = __int_as_float((int)j);
      *tile_slot(sm, off) = p;
    }
    base += total;
  }
  // pad to whole rows of 8 with points at infinity (never the best, never the runner-up)
  const int T8 = (base + 7) & ~7;
  if (base + lane < T8) *tile_slot(sm, base + lane) = make_float4(__int_as_float(0x7f800000), __int_as_float(0x7f800000),
                                                                   __int_as_float(0x7f800000), __int_as_float(-1));
  __syncwarp();
  T_out = T8;
  return true;
}

// every lane evaluates every staged point (c enters seeded)
template <class SM>
__device__ __forceinline__ void tile_scan(SM& sm, int T8, float qx, float qy, float qz, Cand& c) {
  const float4* row = tile_row(sm, 0);
  for (int k = 0; k < T8; k += 8, row += kBlock / 4) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 p = row[u];
      const float d = dist32(p, qx, qy, qz);
      const int j = __float_as_int(p.w);
      if (j != c.j1) cand_update(c, d, j);
    }
  }
}

// ---- stage 1b ------------------------------------------------------------------------------------
__device__ __forceinline__ int zigzag(int t) { return (t & 1) ? -((t + 1) >> 1) : (t >> 1); }  // 0,-1,1,-2,2
// distance from a query at offset f inside its cell to the slab of cells d cells away (d != 0)
__device__ __forceinline__ float slab_dist(int d, float f, float h) {
  return d > 0 ? (float)d * h - f : (d < 0 ? f + (float)(-d - 1) * h : 0.f);
}

// Thread-level fp32 scan of a ball wider than a cell: covers the cells within `kmax` (<= kBallRings) of
// the query's cell.  r2 = initial squared search radius (inflated); the ball shrinks to
// (dist(best) + 2*delta)^2 as candidates appear.  On return r2 is the final (inflated) radius.
template <int SB = kScanBatch, class SM>
__device__ __forceinline__ void ball_scan(const GridDev& g, SM& sm, int cx, int cy, int cz, float fx,
                                          float fy, float fz, float qx, float qy, float qz, float e,
                                          float delta, int kmax, const Best& b, float& r2, Cand& c) {
  const int tid = threadIdx.x;
  const float h = (float)g.h, inv_h = (float)g.inv_h;
  const float hh = h * h;
  cand_seed(g, b, qx, qy, qz, c);
  // Row enumeration state.  Slabs are visited 0,-1,+1,-2,+2,... in z, and inside a z slab the same way in y.
  // On either axis the slab distance grows monotonically on each side, so two consecutive misses (one per
  // side) end that axis -- the walk costs ~(2r/h+3)^2 tests for the CURRENT radius, not (2kmax+1)^2.
  int tz = 0, ty = 0, zmiss = 0, ymiss = 0;
  float remz = 0.f;
  bool zfresh = true;       // remz must be (re)computed: new slab, or the ball shrank since the last chunk
  bool exhausted = false;
  while (!exhausted) {
    // ---- gather the cell-table indices of up to 8 rows the ball touches, nearest slabs first
    int n = 0;
    zfresh = true;
    while (n < 8) {
      if (zfresh) {
        if (tz > 2 * kmax || zmiss >= 2) { exhausted = true; break; }
        const float ddz = slab_dist(zigzag(tz), fz, h);
        remz = r2 - ddz * ddz;
        zfresh = false;
        if (remz < 0.f) { ++zmiss; ++tz; ty = 0; ymiss = 0; zfresh = true; continue; }
        zmiss = 0;
        if ((unsigned)(cz + zigzag(tz)) >= (unsigned)g.nz) { ++tz; ty = 0; ymiss = 0; zfresh = true; continue; }
      }
      if (ty > 2 * kmax || ymiss >= 2) { ++tz; ty = 0; ymiss = 0; zfresh = true; continue; }
      const int dy = zigzag(ty);
      ++ty;
      const float ddy = slab_dist(dy, fy, h);
      const float rem = remz - ddy * ddy;
      if (rem < 0.f) { ++ymiss; continue; }
      ymiss = 0;
      const int y = cy + dy, z = cz + zigzag(tz);
      if ((unsigned)y >= (unsigned)g.ny) continue;
      const float w = sqrtf(rem) * 1.00001f + 1e-6f * h;
      const int x0 = max(cx + (int)floorf((fx - w) * inv_h), 0);
      const int x1 = min(cx + (int)floorf((fx + w) * inv_h), g.nx - 1);
      if (x0 > x1) continue;
      const unsigned row = (unsigned)(((size_t)z * g.ny + y) * g.nx);
      sm.rng[2 * n][tid] = row + (unsigned)x0;
      sm.rng[2 * n + 1][tid] = row + (unsigned)x1 + 1u;
      ++n;
    }
    // ---- resolve them with independent loads
    unsigned va[8], vb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      va[k] = 0; vb[k] = 0;
      if (k < n) {
        va[k] = __ldg(g.cell_start + sm.rng[2 * k][tid]);
        vb[k] = __ldg(g.cell_start + sm.rng[2 * k + 1][tid]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < n) { sm.rng[2 * k][tid] = va[k]; sm.rng[2 * k + 1][tid] = vb[k]; }
    scan_rows<SB>(g, sm, n, qx, qy, qz, c);
    // ---- shrink the ball to the (error-inflated) fp32 distance of the current best, plus the margin
    if (c.j1 >= 0) {
      const float rr = sqrtf(c.d1 + fp32_tol(c.d1, e)) * 1.00001f + 2.0f * delta;
      r2 = fminf(r2, rr * rr * 1.00001f + 1e-6f * hh);
    }
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

// ---- stage 2 -------------------------------------------------------------------------------------
// Whole warp works on one query.  `b` enters identical on all lanes and leaves identical.
template <bool EXACT>
__device__ __forceinline__ void ring_search_warp(const GridDev& g, int cx, int cy, int cz, float qx,
                                                 float qy, float qz, double sx, double sy,
                                                 double sz, float e, Best& b) {
  const int lane = threadIdx.x & 31;
  // rings are laid around the in-grid cell nearest to the query, so a query outside the grid does not
  // walk rings of non-existent rows; row distances are measured from the true query position.
  const int ccy = min(max(cy, 0), g.ny - 1), ccz = min(max(cz, 0), g.nz - 1);
  const long long kgrid = max(max(ccy, g.ny - 1 - ccy), max(ccz, g.nz - 1 - ccz));
  for (long long base = 0;; base += 32) {
    const double R = sqrt(b.d2);
    long long k = (long long)fmin(ceil(R * g.inv_h) + (double)(abs(cy - ccy) + abs(cz - ccz) > 0 ? 1 : 0), 1.0e9);
    if (k > kgrid) k = kgrid;
    if (k < 0) k = 0;
    const long long total = (2 * k + 1) * (2 * k + 1);
    if (base >= total) break;
    const long long f = base + lane;
    unsigned beg = 0, end = 0;
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
      const int y = ccy + dy, z = ccz + dz;
      if ((unsigned)y < (unsigned)g.ny && (unsigned)z < (unsigned)g.nz) {
        // lower bound of the (y,z)-distance between the query and the slab of cells of this row
        const double ylo = g.g0[1] + (double)y * g.h, zlo = g.g0[2] + (double)z * g.h;
        const double ddy = fmax(fmax(ylo - sy, sy - (ylo + g.h)), 0.0);
        const double ddz = fmax(fmax(zlo - sz, sz - (zlo + g.h)), 0.0);
        const double dyz2 = ddy * ddy + ddz * ddz;
        if (dyz2 * (1.0 - 1e-9) < b.d2) {
          const double w = sqrt(fmax(b.d2 - dyz2 * (1.0 - 1e-9), 0.0)) * (1.0 + 1e-9) + 1e-300;
          const int x0 = max(cell_coord(sx - w, g.g0[0], g.inv_h), 0);
          const int x1 = min(cell_coord(sx + w, g.g0[0], g.inv_h), g.nx - 1);
          if (x0 <= x1) {
            const size_t row = ((size_t)z * g.ny + y) * g.nx;
            beg = __ldg(g.cell_start + row + x0);
            end = __ldg(g.cell_start + row + x1 + 1);
          }
        }
      }
    }
    // fp32 pass over this lane's row
    Cand c;
    c.d1 = 3.0e38f; c.d2nd = 3.0e38f; c.j1 = -1;
    for (unsigned j = beg; j < end; ++j) cand_update(c, dist32(__ldg(g.p32 + j), qx, qy, qz), (int)j);
    // warp arg-min of d1 and the runner-up
    float m1 = c.d1;
    int mj = c.j1, ml = lane;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, m1, m);
      const int oj = __shfl_xor_sync(0xffffffffu, mj, m);
      const int ol = __shfl_xor_sync(0xffffffffu, ml, m);
      if (od < m1 || (od == m1 && ol < ml)) { m1 = od; mj = oj; ml = ol; }
    }
    if (mj >= 0 && m1 < filter_bound<EXACT>(b.d2, e)) {  // the batch may improve the best
      if (!EXACT) {
        if ((double)m1 < b.d2) { b.d2 = (double)m1; b.j = mj; b.oidx = __float_as_uint(__ldg(g.p32 + mj).w); }
      } else {
        float m2 = lane == ml ? c.d2nd : c.d1;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) m2 = fminf(m2, __shfl_xor_sync(0xffffffffu, m2, m));
        if (m2 > m1 + 2.5f * fp32_tol(m1, e)) {
          const double d2e = exact_d2(g, mj, sx, sy, sz);  // same address on all lanes
          const unsigned oi = __float_as_uint(__ldg(g.p32 + mj).w);
          if (d2e < b.d2 || (d2e == b.d2 && b.j >= 0 && oi < b.oidx)) { b.d2 = d2e; b.j = mj; b.oidx = oi; }
        } else {
          // contenders within the fp32 error: every lane rescans its row in fp64, then merge
          Best lb = b;
          lb.thr = filter_bound<true>(b.d2, e);
          scan_range_exact(g, beg, end, qx, qy, qz, sx, sy, sz, e, lb);
          if (lb.j < 0) lb.oidx = kNoIdx;
#pragma unroll
          for (int m = 16; m > 0; m >>= 1) {
            const double od2 = __shfl_xor_sync(0xffffffffu, lb.d2, m);
            const int oj = __shfl_xor_sync(0xffffffffu, lb.j, m);
            const unsigned oo = __shfl_xor_sync(0xffffffffu, lb.oidx, m);
            best_merge(lb, od2, oj, oo);
          }
          b.d2 = lb.d2; b.j = lb.j; b.oidx = lb.j >= 0 ? lb.oidx : kNoIdx;
        }
      }
    }
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

// Motion budget certified by a finished fp32 scan: the query may move this far (strictly less) before
// its nearest neighbour -- or the fact that nothing is within maxdist -- can change.
//   rcert2  squared radius of the ball that was completely scanned (not inflated)
__device__ __forceinline__ float motion_budget(const Best& b, const Cand& c, float e, float rcert2,
                                               double maxdist2) {
  if (b.j >= 0) {
    // everything except the winner is at least D2 away: scanned points by their fp32 distance minus the
    // fp32 error, unscanned points by the radius of the scanned ball
    const float second = fminf(c.d2nd - fp32_tol(c.d2nd, e), rcert2);
    const double D1 = sqrt(b.d2) * (1.0 + 1e-12);
    const double D2 = sqrt(fmax((double)second, 0.0)) * (1.0 - 1e-7);
    const double bud = 0.5 * (D2 - D1) - 1e-12;
    return bud > 0.0 ? __double2float_rd(bud) : 0.f;
  }
  // no pair: the nearest point of all is at least `nearest` away (a scanned point just outside maxdist,
  // or the edge of the scanned ball); pairs stay impossible while the query moves less than the slack
  const float nearest = fminf(c.d1 - fp32_tol(c.d1, e), rcert2);
  const double bud = (sqrt(fmax((double)nearest, 0.0)) * (1.0 - 1e-7) - sqrt(maxdist2) * (1.0 + 1e-12)) - 1e-12;
  return bud > 0.0 ? __double2float_rd(bud) : 0.f;
}

#if B200_TILE
// ---- variant with the cooperative tile form of stage 1 (experiment, see B200_TILE above)
// Warp-synchronous search: every lane of every warp calls this once per tile (it contains warp
// collectives, no block barrier).  `active` lanes carry a query s (in the grid's frame); seed_j is the
// sorted position of the neighbour found for this query last time (-1: none); delta >= 0 asks for a
// motion budget (see header).  budget_out = certified budget (0 when none could be certified).
// `dense`: (warp-uniform hint) the batch's queries are close together in the data scan's cell order, so the
// cooperative tile form of stage 1 is worth trying.
template <bool EXACT, class SM>
__device__ __forceinline__ void nn_warp_search(const GridDev& g, SM& sm, bool active, double sx,
                                               double sy, double sz, double maxdist2, int seed_j,
                                               float delta, int& out_j, double& out_d2,
                                               unsigned& out_oidx, float& budget_out,
                                               unsigned& stage2_count, bool dense = false) {
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  Best b;
  b.d2 = maxdist2;
  b.j = -1;
  b.oidx = kNoIdx;
  b.thr = 0.f;
  bool need2 = false;
  float qx = 0.f, qy = 0.f, qz = 0.f, e = 0.f;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  int cx = 0, cy = 0, cz = 0;
  int cls = 0;          // 0: nothing to scan, 1: ball inside the stencil's guaranteed radius, 2: wider ball
  double rs = 0.0, rs2 = 0.0;
  const float hh = (float)(g.h * g.h);
  budget_out = 0.f;
  if (active) {
    const double ox = fmax(fmax(g.bbox_lo[0] - sx, sx - g.bbox_hi[0]), 0.0);
    const double oy = fmax(fmax(g.bbox_lo[1] - sy, sy - g.bbox_hi[1]), 0.0);
    const double oz = fmax(fmax(g.bbox_lo[2] - sz, sz - g.bbox_hi[2]), 0.0);
    const double dbox2 = ox * ox + oy * oy + oz * oz;
    if (dbox2 < maxdist2) {  // false for NaN queries as well
      qx = (float)(sx - g.c[0]); qy = (float)(sy - g.c[1]); qz = (float)(sz - g.c[2]);
      e = query_err(qx, qy, qz, g.bmax);
      cx = cell_coord(sx, g.g0[0], g.inv_h);
      cy = cell_coord(sy, g.g0[1], g.inv_h);
      cz = cell_coord(sz, g.g0[2], g.inv_h);
      // offsets of the query inside its cell (meaningful also for queries outside the grid)
      fx = (float)((sx - g.g0[0]) - (double)cx * g.h);
      fy = (float)((sy - g.g0[1]) - (double)cy * g.h);
      fz = (float)((sz - g.g0[2]) - (double)cz * g.h);
      if (seed_j >= 0 && (unsigned)seed_j < g.n) {
        const double ds = EXACT ? exact_d2(g, seed_j, sx, sy, sz)
                                : (double)dist32(__ldg(g.p32 + seed_j), qx, qy, qz);
        if (ds < maxdist2) { b.d2 = ds; b.j = seed_j; b.oidx = __float_as_uint(__ldg(g.p32 + seed_j).w); }
      }
      const double rg2 = g.h * g.h * (1.0 - 1e-9);
      rs = sqrt(b.d2) + 2.0 * (double)delta;     // radius to certify
      rs2 = rs * rs;
      cls = rs2 <= rg2 ? 1 : 2;
    } else {
      // farther than maxdist from the model's bounding box: nothing can pair until it comes closer
      const double slack = sqrt(dbox2) * (1.0 - 1e-9) - sqrt(maxdist2) * (1.0 + 1e-12);
      if (slack > 0.0 && slack < 1.0e30) budget_out = __double2float_rd(slack);
    }
  }
  const float r2 = fminf(__double2float_ru(rs2), 3.0e38f) * 1.00001f + 1e-6f * hh;   // inflated for fp32 rounding
  Cand c;
  c.d1 = 3.0e38f; c.d2nd = 3.0e38f; c.j1 = -1;
  float rc2 = 0.f;
  bool tiled = false;
#if B200_TILE
  if (dense && __popc(__ballot_sync(0xffffffffu, cls == 1)) >= B200_TILE_MIN_LANES) {
    unsigned va[9], vb[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) va[r] = vb[r] = 0;
    if (cls == 1) stencil_ranges(g, cx, cy, cz, fx, fy, fz, r2, va, vb, rc2);
    int T8 = 0;
    tiled = tile_stage(g, sm, va, vb, T8);
    if (tiled) {
      if (cls == 1) cand_seed(g, b, qx, qy, qz, c);
      tile_scan(sm, T8, qx, qy, qz, c);
      __syncwarp();   // the tile is dead from here on; lanes may reuse their table columns
      tile_stat(0, 1); tile_stat(1, (unsigned long long)T8);
    } else {
      tile_stat(2, 1);
    }
  }
#endif
  if (cls == 1) {
    // the ball fits inside the stencil's guaranteed radius: stage 1 alone is exact
    int nrows = -1;
    if (!tiled) nrows = stencil_scan(g, sm, cx, cy, cz, fx, fy, fz, qx, qy, qz, r2, b, c, rc2);
    // certified radius: the requested ball, or better the (de-inflated) reach of the scanned cells
    const float rcert2 = fmaxf(__double2float_rd(rs2), rc2 * 0.99999f - 2e-6f * hh);
    const bool settled = finalize_cand<EXACT>(g, c, e, sx, sy, sz, b);
    if (!settled) {  // rare: two contenders closer than the fp32 error -> settle in fp64 over the lane's own runs
      b.thr = filter_bound<true>(b.d2, e);
      if (nrows >= 0) {
        for (int k = 0; k < nrows; ++k)
          scan_range_exact(g, sm.rng[2 * k][tid], sm.rng[2 * k + 1][tid], qx, qy, qz, sx, sy, sz, e, b);
      } else {
        unsigned va[9], vb[9];
        float unused;
        stencil_ranges(g, cx, cy, cz, fx, fy, fz, r2, va, vb, unused);
#pragma unroll
        for (int r = 0; r < 9; ++r) scan_range_exact(g, va[r], vb[r], qx, qy, qz, sx, sy, sz, e, b);
      }
    }
    if (settled) budget_out = motion_budget(b, c, e, rcert2, maxdist2);
  } else if (cls == 2) {
    // (tried: a first pass with a 1.5-cell ball for unseeded queries -- slower, 1.22 vs 1.01 ms at iteration 0)
    const double kneed = ceil(rs * g.inv_h);
    const int kmax = (int)fmin(kneed, (double)kBallRings);
    float rb2 = r2;
    const Best seed = b;
    ball_scan(g, sm, cx, cy, cz, fx, fy, fz, qx, qy, qz, e, delta, kmax, b, rb2, c);
    const bool settled = finalize_cand<EXACT>(g, c, e, sx, sy, sz, b);
    // certified radius: the final ball, de-inflated, but never more than the scanned rings cover
    const float cover = (float)kmax * (float)g.h;
    const float rcert2 = fminf(fmaxf(rb2 - 1e-6f * hh, 0.f) * 0.99997f, cover * cover * 0.99999f);
    if (!settled) { b = seed; need2 = true; }                       // fp32-ambiguous winner
    else if (kneed > (double)kBallRings) {                          // ball wider than the scan covers
      const double cv = (double)kmax * g.h;
      need2 = !(b.j >= 0 && b.d2 <= cv * cv * (1.0 - 1e-9));
    }
    if (settled && !need2) budget_out = motion_budget(b, c, e, rcert2, maxdist2);
  }
  unsigned todo = __ballot_sync(0xffffffffu, need2);
  stage2_count += lane == 0 ? __popc(todo) : 0;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    Best wb;
    wb.d2 = __shfl_sync(0xffffffffu, b.d2, src);
    wb.j = __shfl_sync(0xffffffffu, b.j, src);
    wb.oidx = __shfl_sync(0xffffffffu, b.oidx, src);
    wb.thr = 0.f;
    const double wsx = __shfl_sync(0xffffffffu, sx, src), wsy = __shfl_sync(0xffffffffu, sy, src),
                 wsz = __shfl_sync(0xffffffffu, sz, src);
    const float wqx = __shfl_sync(0xffffffffu, qx, src), wqy = __shfl_sync(0xffffffffu, qy, src),
                wqz = __shfl_sync(0xffffffffu, qz, src), we = __shfl_sync(0xffffffffu, e, src);
    const int wcx = __shfl_sync(0xffffffffu, cx, src), wcy = __shfl_sync(0xffffffffu, cy, src),
              wcz = __shfl_sync(0xffffffffu, cz, src);
    ring_search_warp<EXACT>(g, wcx, wcy, wcz, wqx, wqy, wqz, wsx, wsy, wsz, we, wb);
    if (lane == src) { b.d2 = wb.d2; b.j = wb.j; b.oidx = wb.oidx; }
  }
  out_j = b.j;
  out_d2 = b.d2;
  out_oidx = b.j >= 0 ? b.oidx : kNoIdx;
}

#else
// Warp-synchronous search: every lane of every warp calls this once per tile (it contains warp
// collectives, no block barrier).  `active` lanes carry a query s (in the grid's frame); seed_j is the
// sorted position of the neighbour found for this query last time (-1: none); delta >= 0 asks for a
// motion budget (see header).  budget_out = certified budget (0 when none could be certified).
// SB: candidate loads kept in flight per lane
template <bool EXACT, int SB = kScanBatch, class SM>
__device__ __forceinline__ void nn_warp_search(const GridDev& g, SM& sm, bool active, double sx,
                                               double sy, double sz, double maxdist2, int seed_j,
                                               float delta, int& out_j, double& out_d2,
                                               unsigned& out_oidx, float& budget_out,
                                               unsigned& stage2_count, bool /*dense*/ = false) {
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  Best b;
  b.d2 = maxdist2;
  b.j = -1;
  b.oidx = kNoIdx;
  b.thr = 0.f;
  bool need2 = false;
  float qx = 0.f, qy = 0.f, qz = 0.f, e = 0.f;
  int cx = 0, cy = 0, cz = 0;
  budget_out = 0.f;
  if (active) {
    const double ox = fmax(fmax(g.bbox_lo[0] - sx, sx - g.bbox_hi[0]), 0.0);
    const double oy = fmax(fmax(g.bbox_lo[1] - sy, sy - g.bbox_hi[1]), 0.0);
    const double oz = fmax(fmax(g.bbox_lo[2] - sz, sz - g.bbox_hi[2]), 0.0);
    const double dbox2 = ox * ox + oy * oy + oz * oz;
    if (dbox2 < maxdist2) {  // false for NaN queries as well
      qx = (float)(sx - g.c[0]); qy = (float)(sy - g.c[1]); qz = (float)(sz - g.c[2]);
      e = query_err(qx, qy, qz, g.bmax);
      cx = cell_coord(sx, g.g0[0], g.inv_h);
      cy = cell_coord(sy, g.g0[1], g.inv_h);
      cz = cell_coord(sz, g.g0[2], g.inv_h);
      // offsets of the query inside its cell (meaningful also for queries outside the grid)
      const float fx = (float)((sx - g.g0[0]) - (double)cx * g.h);
      const float fy = (float)((sy - g.g0[1]) - (double)cy * g.h);
      const float fz = (float)((sz - g.g0[2]) - (double)cz * g.h);
      if (seed_j >= 0 && (unsigned)seed_j < g.n) {
        const double ds = EXACT ? exact_d2(g, seed_j, sx, sy, sz)
                                : (double)dist32(__ldg(g.p32 + seed_j), qx, qy, qz);
        if (ds < maxdist2) { b.d2 = ds; b.j = seed_j; b.oidx = __float_as_uint(__ldg(g.p32 + seed_j).w); }
      }
      const float hh = (float)(g.h * g.h);
      const double rg2 = g.h * g.h * (1.0 - 1e-9);
      const double rs = sqrt(b.d2) + 2.0 * (double)delta;     // radius to certify
      const double rs2 = rs * rs;
      Cand c;
      float rcert2;
      bool settled;
      if (rs2 <= rg2) {
        // the ball fits inside the stencil's guaranteed radius: stage 1 alone is exact
        const float r2 = fminf(__double2float_ru(rs2), 3.0e38f) * 1.00001f + 1e-6f * hh;
        float rc2;
        const int nrows = stencil_scan<SB>(g, sm, cx, cy, cz, fx, fy, fz, qx, qy, qz, r2, b, c, rc2);
        // certified radius: the requested ball, or better the (de-inflated) reach of the scanned cells
        rcert2 = fmaxf(__double2float_rd(rs2), rc2 * 0.99999f - 2e-6f * hh);
        settled = finalize_cand<EXACT>(g, c, e, sx, sy, sz, b);
#ifdef B200_COUNT_UNSETTLED   // debug counters, tools/unsettled_stats.py
        if (!settled) atomicAdd(&g_tile_stats[0], 1ull);
#endif
        if (!settled) {  // rare: two contenders closer than the fp32 error -> settle in fp64
          b.thr = filter_bound<true>(b.d2, e);
          for (int k = 0; k < nrows; ++k)
            scan_range_exact(g, sm.rng[2 * k][tid], sm.rng[2 * k + 1][tid], qx, qy, qz, sx, sy, sz, e, b);
        }
      } else {
        // (tried: a first pass with a 1.5-cell ball for unseeded queries -- slower, 1.22 vs 1.01 ms at iteration 0)
        const double kneed = ceil(rs * g.inv_h);
        const int kmax = (int)fmin(kneed, (double)kBallRings);
        float r2 = fminf(__double2float_ru(rs2), 3.0e38f) * 1.00001f + 1e-6f * hh;
        const Best seed = b;
        ball_scan<SB>(g, sm, cx, cy, cz, fx, fy, fz, qx, qy, qz, e, delta, kmax, b, r2, c);
        settled = finalize_cand<EXACT>(g, c, e, sx, sy, sz, b);
        // certified radius: the final ball, de-inflated, but never more than the scanned rings cover
        const float cover = (float)kmax * (float)g.h;
        rcert2 = fminf(fmaxf(r2 - 1e-6f * hh, 0.f) * 0.99997f, cover * cover * 0.99999f);
        if (!settled) { b = seed; need2 = true; }                       // fp32-ambiguous winner
        else if (kneed > (double)kBallRings) {                          // ball wider than the scan covers
          const double cv = (double)kmax * g.h;
          need2 = !(b.j >= 0 && b.d2 <= cv * cv * (1.0 - 1e-9));
        }
#ifdef B200_COUNT_UNSETTLED
        if (!settled) atomicAdd(&g_tile_stats[1], 1ull);
#endif
      }
      if (settled && !need2) budget_out = motion_budget(b, c, e, rcert2, maxdist2);
#ifdef B200_COUNT_UNSETTLED
      if (settled && !need2 && budget_out < 2e-3f) atomicAdd(&g_tile_stats[2], 1ull);
      if (settled && !need2 && budget_out < 2e-4f) atomicAdd(&g_tile_stats[3], 1ull);
#endif
    } else {
      // farther than maxdist from the model's bounding box: nothing can pair until it comes closer
      const double slack = sqrt(dbox2) * (1.0 - 1e-9) - sqrt(maxdist2) * (1.0 + 1e-12);
      if (slack > 0.0 && slack < 1.0e30) budget_out = __double2float_rd(slack);
    }
  }
  unsigned todo = __ballot_sync(0xffffffffu, need2);
  stage2_count += lane == 0 ? __popc(todo) : 0;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    Best wb;
    wb.d2 = __shfl_sync(0xffffffffu, b.d2, src);
    wb.j = __shfl_sync(0xffffffffu, b.j, src);
    wb.oidx = __shfl_sync(0xffffffffu, b.oidx, src);
    wb.thr = 0.f;
    const double wsx = __shfl_sync(0xffffffffu, sx, src), wsy = __shfl_sync(0xffffffffu, sy, src),
                 wsz = __shfl_sync(0xffffffffu, sz, src);
    const float wqx = __shfl_sync(0xffffffffu, qx, src), wqy = __shfl_sync(0xffffffffu, qy, src),
                wqz = __shfl_sync(0xffffffffu, qz, src), we = __shfl_sync(0xffffffffu, e, src);
    const int wcx = __shfl_sync(0xffffffffu, cx, src), wcy = __shfl_sync(0xffffffffu, cy, src),
              wcz = __shfl_sync(0xffffffffu, cz, src);
    ring_search_warp<EXACT>(g, wcx, wcy, wcz, wqx, wqy, wqz, wsx, wsy, wsz, we, wb);
    if (lane == src) { b.d2 = wb.d2; b.j = wb.j; b.oidx = wb.oidx; }
  }
  out_j = b.j;
  out_d2 = b.d2;
  out_oidx = b.j >= 0 ? b.oidx : kNoIdx;
}

#endif

}  // namespace b200
