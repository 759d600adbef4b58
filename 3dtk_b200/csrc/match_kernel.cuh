// match_kernel.cuh -- the fused match: ONE persistent cooperative launch runs every iteration of icp6D::match.
//
// Replaces the loop of icp6D::match (reference src/slam6d/icp6D.cc:104-285) with everything it calls per
// iteration: Scan::getPtPairs (scan.cc:1220-1260), SearchTree::getPtPairs (searchTree.cc:92-188) incl. the
// CLOSEST_PLANE_SIMPLE projection, the pair walk of icp6Dminimizer::Align (icp6Dquat.cc:57-71 ...), the 6-DoF solve,
// Scan::transform (scan.cc:851-898: applied on load, no point is moved) and the convergence test (icp6D.cc:266-279).
//
// Per iteration every block walks its CONTIGUOUS share of the (cell-sorted) data scan:
//   stream   per data point  t = X d0,  s = Sinv t.  The point keeps a CANDIDATE LIST from its last full search --
//            kListK model points (fp32 coordinates + position, SoA: list[k][i], 16 B each, coalesced) and the
//            certificate radius R of nn_search.cuh, decremented by the motion of every iteration.  While the nearest
//            listed candidate is closer than what is left of R, it IS the exact nearest neighbour: no search, one
//            gather of its fp64 coordinates, exact distance (reference rounding), rejection, moments.
//   search   points whose certificate is spent are queued per warp and searched 32 at a time (nn_warp_search, all
//            lanes busy), seeded with their nearest listed candidate; the search leaves a new list and radius.
//            Early iterations (large motion) keep only best + runner-up; once the per-iteration motion is below
//            kWideGate cells the scan covers kWideRmin cells and keeps all kListK candidates -- from then on most
//            points are never searched again.
//   solve    per-block moments -> the last block to arrive reduces them in a fixed order (+ the NVLink mailbox
//            all-reduce of a query-sharded match), runs solve.h, updates the pose and the loop state and releases
//            the next iteration through a generation word; the other blocks spin on it (grid barrier with the
//            solve inside).  No host round trip, no launch per iteration.
// Determinism: static point -> thread assignment, thread-private accumulator columns, fixed-shape reductions:
// reruns are bit-identical.
#pragma once
#include <cooperative_groups.h>
#include "icp_kernels.cuh"

namespace b200 {

constexpr int kListK = 4;            // candidates kept per data point
#ifndef B200_WIDE_GATE
#define B200_WIDE_GATE 0.2f         // keep all candidates once a point moves less than this many cells per iteration
#endif
#ifndef B200_WIDE_RMIN
#define B200_WIDE_RMIN 0.55f         // ... and scan at least this many cells around it
#endif
#ifndef B200_CONTIG_TILES
#define B200_CONTIG_TILES 0
#endif
#ifndef B200_ITER_MINBLOCKS
#define B200_ITER_MINBLOCKS 3
#endif
constexpr int kQueueCap = 64;        // per-warp queue of points waiting for a full search
#ifndef B200_COOP_MAX
#define B200_COOP_MAX 48
#endif
constexpr int kCoopMax = B200_COOP_MAX;   // at most this many leftover points of a block are searched one per warp

// device workspace of one match (owned by the context)
struct MatchWork {
  float4* list;                      // [kListK][stride]
  float* rrem;                       // [nd] what is left of the certificate radius
  uint32_t stride;
  double* partials;                  // [grid][NS_MAX]
  unsigned* sync;                    // [0] arrivals (monotonic)  [1] released generation  [2] stage-2 count  [3] searches
  double* rms_log;
  unsigned long long* npairs_log;
  unsigned* stage2_log;              // [max_iter][2]
  unsigned long long* clock_log;     // [max_iter + 1] globaltimer at kernel start and at the end of every iteration
  float wide_gate, wide_rmin;        // in cells (defaults B200_WIDE_GATE / B200_WIDE_RMIN)
  unsigned long long* dbg;           // optional [grid][8] per-block phase stamps of the last iteration (NULL: off)
};

template <int NS>
struct MatchSmem {
  SearchSmem search;
  int queue[kWarps][kQueueCap];
  int leftover[kWarps * 32];         // end-of-iteration merge of the warps' partial queues
  int left_count[kWarps];
  double acc[NS][kBlock];
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool NAPX, bool PLANE, bool EXACT>
__global__ void __launch_bounds__(kBlock, B200_ITER_MINBLOCKS)
icp_match_kernel(const __grid_constant__ GridDev model, const double4* __restrict__ dq,
                 const double4* __restrict__ dn, uint32_t nd, IterState* __restrict__ st, double maxdist2, int rnd,
                 const __grid_constant__ MatchWork mw, const __grid_constant__ CommDev comm) {
  constexpr int NS = NAPX ? (int)NS_NAPX : (int)NS_P2P;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  MatchSmem<NS>& sm = *reinterpret_cast<MatchSmem<NS>*>(dyn_smem);
  __shared__ XfSmem xf;
  __shared__ int s_done, s_iter, s_iters_run, s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ntiles = (nd + kBlock - 1) / kBlock;
#if B200_CONTIG_TILES
  // contiguous share of the scan: consecutive trips of a warp stay in one neighbourhood of the model
  const uint32_t tile0 = (uint32_t)((unsigned long long)ntiles * blockIdx.x / gridDim.x);
  const uint32_t tile1 = (uint32_t)((unsigned long long)ntiles * (blockIdx.x + 1) / gridDim.x);
#endif
  const float hcell = (float)model.h;
  const float dmax = 0.1f * hcell, dmin = 1e-3f * hcell;
  const float wide_gate = mw.wide_gate * hcell, wide_rmin = mw.wide_rmin * hcell;
  const float maxdist2_up = __double2float_ru(maxdist2) * 1.000002f;
  const float maxdist_up = sqrtf(maxdist2_up) * 1.000001f;
  float4* const l0 = mw.list;
  SmemAcc acc{&sm.acc[0][tid]};
  PairCtx pc{&model, dn, &xf};
  int* queue = sm.queue[warp];
  if (blockIdx.x == 0 && tid == 0) mw.clock_log[0] = global_timer();

  for (unsigned epoch = 0;; ++epoch) {
    // ---- state of this iteration (written by the block that ran the previous solve)
    load_xf(xf, st);
    if (tid == 0) {
      s_done = __ldcg(&st->done);
      s_iter = __ldcg(&st->iter);
      s_iters_run = __ldcg(&st->iters_run);
    }
    __syncthreads();
    if (s_done) break;
    auto mark = [&](int k) { if (mw.dbg && tid == 0) mw.dbg[(size_t)blockIdx.x * 8 + k] = global_timer(); };
    mark(0);
    const unsigned iter_salt = (unsigned)s_iter * 0x9E3779B9u;
    // a point may skip its search while its certificate holds; needs the previous pose (not the first
    // iteration) and every point visited every iteration (no subsampling)
    const bool can_skip = s_iters_run > 0 && rnd <= 1;
#pragma unroll 4
    for (int k = 0; k < NS; ++k) acc[k] = 0.0;
    unsigned stage2 = 0, searches = 0;
    int qcount = 0;            // warp-uniform number of queued points

    // full search of up to 32 queued points (list[0..nb)), one per lane
    auto run_batch = [&](const int* list, int nb) {
      const bool on = lane < nb;
      uint32_t i = 0;
      double tx = 0, ty = 0, tz = 0, sx = 0, sy = 0, sz = 0;
      float step = 0.f, delta = 0.f;
      int seed = -1;
      bool wide = false;
      if (on) {
        i = (uint32_t)list[lane];
        if (s_iters_run > 0) seed = __float_as_int(__ldcs(l0 + i).w);
        query_state(xf, dq, i, can_skip, dmin, dmax, tx, ty, tz, sx, sy, sz, step, delta);
        wide = can_skip && step <= wide_gate;
      }
      // one mode per batch (warp-uniform), so that only one of the two scan bodies is hot at a time: all
      // candidates are kept once every point of the batch moves little
      const bool wide_batch = __all_sync(0xffffffffu, wide || !on) && can_skip;
      SearchOut<kListK> so;
      if (wide_batch)
        nn_warp_search<EXACT, kListK, true>(model, sm.search, on, sx, sy, sz, maxdist2, seed, delta, wide_rmin, so, stage2);
      else
        nn_warp_search<EXACT, kListK, false>(model, sm.search, on, sx, sy, sz, maxdist2, seed, delta, 0.f, so, stage2);
      if (on) {
        // (best + runner-up mode keeps only the head: the other entries stay what they were -- real model points of
        //  an older list, harmless as extra candidates -- except in the first iteration, which initialises them)
#pragma unroll
        for (int k = 0; k < kListK; ++k) {
          if (k == 0 || wide_batch || !can_skip) {
            float4 p = make_float4(3.0e38f, 3.0e38f, 3.0e38f, __int_as_float(-1));   // empty entry: far away
            if (so.lj[k] >= 0) { p = __ldg(model.p32 + so.lj[k]); p.w = __int_as_float(so.lj[k]); }
            __stcs(l0 + (size_t)k * mw.stride + i, p);
          }
        }
        __stcs(mw.rrem + i, so.R);
        if (so.j >= 0) accumulate_pair<NAPX, PLANE>(pc, acc, i, so.j, tx, ty, tz);
      }
      __syncwarp();
    };

    // streaming part of one point: certificate test inline, otherwise report "needs a search"
    auto stream_one = [&](uint32_t i, bool active) -> bool {
      if (rnd > 1 && active) active = (hash32(i ^ iter_salt) % (unsigned)rnd) == 0u;
      if (!active) return false;
      if (!can_skip) return true;
      double tx, ty, tz, sx, sy, sz;
      float step, delta;
      query_state(xf, dq, i, true, dmin, dmax, tx, ty, tz, sx, sy, sz, step, delta);
      const float rleft = __ldcs(mw.rrem + i) - step;
      __stcs(mw.rrem + i, rleft);
      if (!(rleft > 0.f)) return true;     // certificate spent: the list is not even read
      float4 L[kListK];
#pragma unroll
      for (int k = 0; k < kListK; ++k) L[k] = __ldcs(l0 + (size_t)k * mw.stride + i);
      const float qx = (float)(sx - model.c[0]), qy = (float)(sy - model.c[1]), qz = (float)(sz - model.c[2]);
      const float e = query_err(qx, qy, qz, model.bmax);
      // nearest and second nearest listed candidate (empty entries are far away)
      float dl[kListK];
      float d1 = 3.0e38f, d2 = 3.0e38f;
      int k1 = 0;
#pragma unroll
      for (int k = 0; k < kListK; ++k) {
        dl[k] = dist32(L[k], qx, qy, qz);
        d2 = fminf(d2, fmaxf(dl[k], d1));
        if (dl[k] < d1) { d1 = dl[k]; k1 = k; }
      }
      bool search = true;
      {
        const float tol1 = fp32_tol(d1, e);
        const float r2 = __fmul_rd(__fmul_rd(rleft, rleft), 0.999998f);
        const float band = d1 + 2.5f * tol1;
        if (band < r2) {
          // every listed candidate inside the band is closer than anything that is not listed: the exact nearest
          // neighbour is one of them -- the fp32 winner when it stands alone (the usual case), otherwise (two
          // near-equidistant candidates) the fp64 arg-min of the band
          search = false;
          double d2e, px, py, pz;
          if (!EXACT || d2 > band) {
            float4 w = L[0];
#pragma unroll
            for (int k = 1; k < kListK; ++k) if (k1 == k) w = L[k];
            const int j = __float_as_int(w.w);
            const double2 pa = __ldg(reinterpret_cast<const double2*>(model.p64 + j));
            px = pa.x; py = pa.y;
            pz = __ldg(reinterpret_cast<const double*>(model.p64 + j) + 2);
            if (EXACT) {
              const double ex = __dsub_rn(px, sx), ey = __dsub_rn(py, sy), ez = __dsub_rn(pz, sz);
              d2e = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
            } else {
              d2e = (double)d1;
            }
          } else {
            exact_pick<kListK>(model, L, dl, band, sx, sy, sz, d2e, px, py, pz);
          }
          if (d2e < maxdist2) accumulate_pair_pt<NAPX, PLANE>(pc, acc, i, px, py, pz, tx, ty, tz);
        } else if (rleft >= maxdist_up && d1 >= maxdist2_up + tol1) {
          search = false;   // nothing within maxdist, listed or not
        }
      }
      if (search && k1 != 0 && d1 < 3.0e38f) {
        // the nearest listed candidate seeds the search
        float4 w = L[1];
#pragma unroll
        for (int k = 2; k < kListK; ++k) if (k1 == k) w = L[k];
        __stcs(l0 + i, w);
      }
      return search;
    };
    // queue the points that need a full search; run them 32 at a time so every lane works
    auto enqueue = [&](uint32_t i, bool search) {
      const unsigned smask = __ballot_sync(0xffffffffu, search);
      if (smask) {
        if (search) queue[qcount + __popc(smask & ((1u << lane) - 1u))] = (int)i;
        qcount += __popc(smask);
        if (lane == 0) searches += __popc(smask);
        __syncwarp();
        if (qcount >= 32) { qcount -= 32; run_batch(queue + qcount, 32); }
      }
    };
#if B200_CONTIG_TILES
    for (uint32_t tile = tile0; tile < tile1; ++tile) {
#else
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#endif
      const uint32_t i = tile * kBlock + tid;
      enqueue(i, stream_one(i, i < nd));
    }
    // leftovers: merge the warps' partial queues so the remaining searches run in full batches
    if (lane == 0) sm.left_count[warp] = qcount;
    __syncthreads();
    mark(1);
    {
      int off = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const int n = sm.left_count[w];
        if (w < warp) off += n;
        total += n;
      }
      if (lane < qcount) sm.leftover[off + lane] = queue[lane];
      __syncthreads();
      if (can_skip && total <= kCoopMax) {
        // few stragglers: one point per warp at a time, the whole warp on its search (nn_coop_search); what that
        // form cannot settle goes to the per-lane batch below
        int nfall = 0;
        for (int q = warp; q < total; q += kWarps) {
          const uint32_t i = (uint32_t)sm.leftover[q];
          double tx, ty, tz, sx, sy, sz;
          float step, delta;
          query_state(xf, dq, i, true, dmin, dmax, tx, ty, tz, sx, sy, sz, step, delta);
          const int seed = __float_as_int(__ldcs(l0 + i).w);
          SearchOut<kListK> so;
          const bool ok = nn_coop_search<EXACT, kListK>(model, sx, sy, sz, maxdist2, seed, delta, wide_rmin, so);
          if (ok) {
            if (lane < kListK) {
              float4 p = make_float4(3.0e38f, 3.0e38f, 3.0e38f, __int_as_float(-1));
              int lj = so.lj[0];
#pragma unroll
              for (int k = 1; k < kListK; ++k) if (lane == k) lj = so.lj[k];
              if (lj >= 0) { p = __ldg(model.p32 + lj); p.w = __int_as_float(lj); }
              __stcs(l0 + (size_t)lane * mw.stride + i, p);
            }
            if (lane == 0) {
              __stcs(mw.rrem + i, so.R);
              if (so.j >= 0) accumulate_pair<NAPX, PLANE>(pc, acc, i, so.j, tx, ty, tz);
            }
          } else {
            if (lane == 0) queue[nfall] = (int)i;   // the warp's queue is free by now
            ++nfall;
          }
        }
        __syncwarp();
        if (lane == 0) sm.left_count[warp] = nfall;
        __syncthreads();
        int off2 = 0, total2 = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          const int n = sm.left_count[w];
          if (w < warp) off2 += n;
          total2 += n;
        }
        __syncthreads();                        // everyone has read leftover[] and left_count[]
        if (lane < nfall) sm.leftover[off2 + lane] = queue[lane];
        total = total2;
        __syncthreads();
      }
      for (int base = warp * 32; base < total; base += kWarps * 32)
        run_batch(sm.leftover + base, min(32, total - base));
    }

    mark(2);
    block_reduce_store<NS>(acc, mw.partials + (size_t)blockIdx.x * NS_MAX);
    mark(3);
    if (lane == 0 && stage2) atomicAdd(mw.sync + 2, stage2);
    if (lane == 0 && searches) atomicAdd(mw.sync + 3, searches);
    // ---- arrive; the last block reduces all partials, runs the solve and releases the next iteration
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned ticket = atomicAdd(mw.sync + 0, 1u);
      s_last = ticket == (epoch + 1u) * gridDim.x - 1u;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      solve_step(st, mw.partials, (int)gridDim.x, mw.rms_log, mw.npairs_log, mw.stage2_log, mw.sync + 2,
                 &sm.acc[0][0], comm);
      __threadfence();
      __syncthreads();
      mark(5);
      if (tid == 0) {
        mw.clock_log[epoch + 1u] = global_timer();
        st_release_u32(mw.sync + 1, epoch + 1u);
      }
    } else if (tid == 0) {
      while (ld_acquire_u32(mw.sync + 1) < epoch + 1u) __nanosleep(40);
    }
    __syncthreads();
    mark(4);
  }
}

}  // namespace b200
