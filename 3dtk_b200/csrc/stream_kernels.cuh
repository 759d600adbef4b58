// stream_kernels.cuh -- the two-kernel form of one ICP iteration (point-to-point moments, every point visited).
//
// One iteration of icp6D::match (reference src/slam6d/icp6D.cc:124-279) is split by WHAT BOUNDS the work:
//
//   icp_stream_kernel   every data point, pure streaming.  A point whose motion budget (nn_search.cuh) still
//                       covers this iteration's motion keeps its neighbour: the pair (t = X d0, cached model
//                       point) is re-evaluated exactly and accumulated.  Everything it needs is contiguous per
//                       tile of 256 points -- d0 (32 B), the cached neighbour's coordinates `pm` (32 B, written by
//                       the search when it found the neighbour, so there is no gather) and the budget (4 B) --
//                       and is brought into shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier,
//                       4-stage ring), so the kernel runs at HBM/L2 streaming rate with registers left for the
//                       23 fp64 moment accumulators.  Points whose budget is spent are appended, in a fixed
//                       order, to the block's segment of the search queue.
//   icp_search_kernel   the queued points only, 32 per warp (all lanes busy): exact grid search
//                       (nn_warp_search), new neighbour / budget / pm, pair accumulation; the last block to
//                       finish reduces the partial moments of BOTH kernels in a fixed order and runs the solve,
//                       pose update and convergence test (solve_step) -- and, for a query-sharded match, the
//                       fused NVLink all-reduce.
//
// Determinism: a stream block owns a contiguous range of tiles and writes its queue segment in (tile, warp,
// lane) order; search batches are assigned to warps statically; all reductions have a fixed shape.  Reruns are
// bit-identical.
#pragma once
#include "icp_kernels.cuh"

namespace b200 {

// ---- mbarrier / TMA (1-D bulk copy) primitives ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy by the TMA engine; `bytes` multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// programmatic dependent launch: the next kernel of the stream may be scheduled while this one drains
// (launch_dependents), and must not touch its predecessor's results before pdl_wait()
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifndef B200_EVICT_FIRST
#define B200_EVICT_FIRST 1
#endif
// L2 policy for the streamed arrays: they are read once per iteration and must not push the model grid (which the
// latency-bound searches live on) out of the L2
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                                 unsigned long long pol) {
#if B200_EVICT_FIRST
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol) : "memory");
#else
  tma_load_1d(dst, src, bytes, bar);
#endif
}

constexpr int kStreamStages = 4;
constexpr int kMaxSegments = 512;          // stream-kernel grid (= queue segments) upper bound

struct StreamStage {
  double4 dq[kBlock];    // data points of the tile, original frame
  double4 pm[kBlock];    // cached neighbour (model frame); w != 0: the point is paired
  float bud[kBlock];     // motion budget left
};
constexpr unsigned kStageBytes = (unsigned)sizeof(StreamStage);

struct StreamSmem {
  StreamStage st[kStreamStages];
  unsigned long long full[kStreamStages];
  unsigned wcount[2][kWarps];
};

// grid = number of queue segments; block b owns tiles [b*tiles_per_seg, (b+1)*tiles_per_seg)
template <bool PLANE>
__global__ void __launch_bounds__(kBlock, 2)
icp_stream_kernel(GridDev model, const double4* __restrict__ dq, const double4* __restrict__ dn, uint32_t nd,
                  const IterState* __restrict__ st, double maxdist2, const double4* __restrict__ pm,
                  float* __restrict__ nn_budget, int* __restrict__ queue, unsigned* __restrict__ seg_count,
                  uint32_t tiles_per_seg, double* __restrict__ partials) {
  if (blockIdx.x == 0) tl_mark(16);
  pdl_wait();                 // the previous iteration's solve has published the pose
  if (blockIdx.x == 0) tl_mark(17);
  pdl_launch_dependents();    // the search kernel's blocks may take SM slots as they free up (they wait, too)
  if (st->done) return;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  StreamSmem& sm = *reinterpret_cast<StreamSmem*>(dyn_smem);
  __shared__ XfSmem xf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ntiles = (nd + kBlock - 1) / kBlock;
  const uint32_t t0 = min(blockIdx.x * tiles_per_seg, ntiles), t1 = min(t0 + tiles_per_seg, ntiles);
  int* const qseg = queue + (size_t)t0 * kBlock;
  const bool can_skip = st->iters_run > 0;
  double acc[NS_P2P];
#pragma unroll
  for (int k = 0; k < NS_P2P; ++k) acc[k] = 0.0;

  if (!can_skip) {
    // first iteration: nothing is cached, every point of the range goes to the search kernel
    const uint32_t lo = t0 * kBlock, hi = min(t1 * kBlock, nd);
    for (uint32_t i = lo + tid; i < hi; i += kBlock) qseg[i - lo] = (int)i;
    if (tid == 0) seg_count[blockIdx.x] = hi > lo ? hi - lo : 0u;
    block_reduce_store_regs<NS_P2P>(acc, partials + (size_t)blockIdx.x * NS_MAX);
    return;
  }

  load_xf(xf, st);
  if (tid == 0) {
    for (int s = 0; s < kStreamStages; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const unsigned long long pol = l2_evict_first_policy();
  auto issue = [&](uint32_t tile, int stage) {   // one thread: arm the barrier, start the three bulk copies
    const size_t base = (size_t)tile * kBlock;
    mbar_arrive_expect_tx(&sm.full[stage], kStageBytes);
    tma_load_1d_hint(sm.st[stage].dq, dq + base, (unsigned)sizeof(sm.st[stage].dq), &sm.full[stage], pol);
    tma_load_1d_hint(sm.st[stage].pm, pm + base, (unsigned)sizeof(sm.st[stage].pm), &sm.full[stage], pol);
    tma_load_1d_hint(sm.st[stage].bud, nn_budget + base, (unsigned)sizeof(sm.st[stage].bud), &sm.full[stage], pol);
  };
  auto is_full = [&](uint32_t tile) { return (size_t)(tile + 1) * kBlock <= (size_t)nd; };
  if (tid == 0)
    for (int s = 0; s < kStreamStages; ++s)
      if (t0 + s < t1 && is_full(t0 + s)) issue(t0 + s, s);

  PairCtx pc{&model, dn, &xf};
  const float dmax = 0.1f * (float)model.h, dmin = 1e-3f * (float)model.h;
  unsigned qn = 0;   // block-uniform: entries written to the segment so far
  for (uint32_t tile = t0, k = 0; tile < t1; ++tile, ++k) {
    const int stage = (int)(k % kStreamStages);
    const uint32_t i = tile * kBlock + tid;
    const bool active = i < nd;
    double4 d0 = make_double4(0, 0, 0, 0), pmv = make_double4(0, 0, 0, 0);
    float bud = 0.f;
    if (is_full(tile)) {
      mbar_wait(&sm.full[stage], (k / kStreamStages) & 1u);
      d0 = sm.st[stage].dq[tid];
      pmv = sm.st[stage].pm[tid];
      bud = sm.st[stage].bud[tid];
    } else if (active) {   // ragged last tile: plain loads
      d0 = dq[i];
      pmv = pm[i];
      bud = nn_budget[i];
    }
    bool search = false;
    if (active) {
      double tx, ty, tz, sx, sy, sz;
      float step, delta;
      query_state_pt(xf, Pt3{d0.x, d0.y}, d0.z, true, dmin, dmax, tx, ty, tz, sx, sy, sz, step, delta);
      const float left = bud - step;
      search = !(left > 0.f);
      if (!search) {
        // the cached neighbour is still THE nearest neighbour; only the pair is re-evaluated
        nn_budget[i] = left;
        if (pmv.w != 0.0) {
          const double ex = __dsub_rn(pmv.x, sx), ey = __dsub_rn(pmv.y, sy), ez = __dsub_rn(pmv.z, sz);
          const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
          if (d2 < maxdist2) accumulate_pair_pt<false, PLANE>(pc, acc, i, pmv.x, pmv.y, pmv.z, tx, ty, tz);
        }
      }
    }
    const unsigned smask = __ballot_sync(0xffffffffu, search);
    if (lane == 0) sm.wcount[k & 1][warp] = __popc(smask);
    __syncthreads();   // every thread has read its slot of the stage; warp counts are visible
    if (tid == 0 && tile + kStreamStages < t1 && is_full(tile + kStreamStages)) issue(tile + kStreamStages, stage);
    unsigned before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const unsigned c = sm.wcount[k & 1][w];
      before += w < warp ? c : 0u;
      total += c;
    }
    if (search) qseg[qn + before + __popc(smask & ((1u << lane) - 1u))] = (int)i;
    qn += total;
  }
  if (tid == 0) seg_count[blockIdx.x] = qn;
  if (blockIdx.x == 0) tl_mark(18);
  block_reduce_store_regs<NS_P2P>(acc, partials + (size_t)blockIdx.x * NS_MAX);
  if (blockIdx.x == 0) tl_mark(19);
}

// dynamic shared memory of icp_search_kernel
struct SearchKSmem {
  SearchSmem search;
  unsigned seg_off[kMaxSegments + 1];
  double acc[NS_P2P][kBlock];
};

template <bool PLANE, bool EXACT>
__global__ void __launch_bounds__(kBlock, 3)
icp_search_kernel(GridDev model, const double4* __restrict__ dq, const double4* __restrict__ dn, uint32_t nd,
                  IterState* __restrict__ st, double maxdist2, int* __restrict__ nn_cache,
                  float* __restrict__ nn_budget, double4* __restrict__ pm, const int* __restrict__ queue,
                  const unsigned* __restrict__ seg_count, int nseg, uint32_t tiles_per_seg,
                  double* __restrict__ partials, unsigned* __restrict__ stage2_counter,
                  double* __restrict__ rms_log, unsigned long long* __restrict__ npairs_log,
                  unsigned* __restrict__ stage2_log, const __grid_constant__ CommDev comm) {
  if (blockIdx.x == 0) tl_mark(0);
  pdl_wait();                 // queue segments, budgets and stream partials are complete
  if (blockIdx.x == 0) tl_mark(1);
  if (st->done) return;
  constexpr int NS = (int)NS_P2P;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  SearchKSmem& sm = *reinterpret_cast<SearchKSmem*>(dyn_smem);
  __shared__ XfSmem xf;
  __shared__ unsigned scan_tmp[kWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_xf(xf, st);
  // exclusive prefix of the segment counts (nseg <= 512: two entries per thread)
  {
    const unsigned c0 = 2 * tid < nseg ? seg_count[2 * tid] : 0u;
    const unsigned c1 = 2 * tid + 1 < nseg ? seg_count[2 * tid + 1] : 0u;
    unsigned v = c0 + c1;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, v, m);
      if (lane >= m) v += o;
    }
    if (lane == 31) scan_tmp[warp] = v;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) wbase += w < warp ? scan_tmp[w] : 0u;
    const unsigned excl = wbase + v - (c0 + c1);
    sm.seg_off[2 * tid] = excl;
    sm.seg_off[2 * tid + 1] = excl + c0;
    if (tid == kBlock - 1) sm.seg_off[kMaxSegments] = excl + c0 + c1;
  }
  const bool can_skip = st->iters_run > 0;
  const float dmax = 0.1f * (float)model.h, dmin = 1e-3f * (float)model.h;
  SmemAcc acc{&sm.acc[0][tid]};
#pragma unroll 4
  for (int k = 0; k < NS; ++k) acc[k] = 0.0;
  __syncthreads();
  const unsigned total = sm.seg_off[kMaxSegments];
  const unsigned seg_stride = tiles_per_seg * kBlock;
  if (blockIdx.x == 0) tl_mark(2);

  PairCtx pc{&model, dn, &xf};
  unsigned stage2 = 0;
  const unsigned nbatch = (total + 31u) / 32u;
  for (unsigned b = blockIdx.x * kWarps + warp; b < nbatch; b += gridDim.x * kWarps) {
    const unsigned e = b * 32u + lane;
    const bool on = e < total;
    uint32_t i = 0;
    double tx = 0, ty = 0, tz = 0, sx = 0, sy = 0, sz = 0;
    float step = 0.f, delta = 0.f;
    int seed = -1;
    if (on) {
      // segment of entry e: last s with seg_off[s] <= e
      int lo = 0, hi = nseg - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sm.seg_off[mid] <= e) lo = mid; else hi = mid - 1;
      }
      i = (uint32_t)queue[(size_t)lo * seg_stride + (e - sm.seg_off[lo])];
      seed = nn_cache[i];
      query_state(xf, dq, i, can_skip, dmin, dmax, tx, ty, tz, sx, sy, sz, step, delta);
    }
    int bj;
    double bd2;
    unsigned boidx;
    float newbud;
    // queue entries are in the data scan's cell order: a full batch with a short index span is "dense"
    const unsigned imax = __reduce_max_sync(0xffffffffu, on ? i : 0u);
    const unsigned imin = __reduce_min_sync(0xffffffffu, on ? i : 0xffffffffu);
    const bool dense = __all_sync(0xffffffffu, on) && imax - imin <= (unsigned)kTileSpan;
    nn_warp_search<EXACT>(model, sm.search, on, sx, sy, sz, maxdist2, seed, delta, bj, bd2, boidx, newbud, stage2,
                          dense);
    if (on) {
      nn_cache[i] = bj;
      nn_budget[i] = newbud;
      double4 pv = make_double4(0.0, 0.0, 0.0, 0.0);
      if (bj >= 0) {
        const double2 pa = __ldg(reinterpret_cast<const double2*>(model.p64 + bj));
        const double pz = __ldg(reinterpret_cast<const double*>(model.p64 + bj) + 2);
        pv = make_double4(pa.x, pa.y, pz, 1.0);
        accumulate_pair_pt<false, PLANE>(pc, acc, i, pa.x, pa.y, pz, tx, ty, tz);
      }
      pm[i] = pv;
    }
    __syncwarp();
  }

  if (blockIdx.x == 0) tl_mark(3);
  // blocks beyond the last batch accumulated nothing: their rows are left out of the reduction
  const unsigned active_blocks = min((nbatch + kWarps - 1u) / kWarps, gridDim.x);
  if (blockIdx.x < active_blocks) block_reduce_store<NS>(acc, partials + (size_t)(nseg + blockIdx.x) * NS_MAX);
  if (lane == 0 && stage2) atomicAdd(stage2_counter, stage2);
  if (blockIdx.x == 0 && tid == 0 && total) atomicAdd(stage2_counter + 1, total);
  if (blockIdx.x == 0) tl_mark(4);
  // ---- the last block to get here reduces all partials (stream + search) and runs the solve
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(stage2_counter + 2, 1u);
    is_last = ticket == gridDim.x - 1;
    if (is_last) stage2_counter[2] = 0;   // ready for the next launch
  }
  __syncthreads();
  // every block is past its searches: the next iteration's stream kernel may be scheduled (it waits for the
  // solve below through griddepcontrol.wait, which only returns when this whole grid has finished)
  pdl_launch_dependents();
  if (is_last) {
    tl_mark(7);
    __threadfence();
    solve_step(st, partials, nseg + (int)active_blocks, false, rms_log, npairs_log, stage2_log, stage2_counter, &sm.acc[0][0], comm);
  }
}

}  // namespace b200
