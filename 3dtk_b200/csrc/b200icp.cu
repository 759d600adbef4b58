// b200icp.cu -- C ABI (include/b200icp.h) over the sm_100a kernels.  One translation unit.
//
// Host orchestration only: buffer ownership, grid-build driver, the launch loop of the fused match.
// All numerics run in the kernels of grid_build.cuh / nn_search.cuh / icp_kernels.cuh (device) or
// solve.h (the O(1) solve, shared by device and host).
#include "../../include/b200icp.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "grid_build.cuh"
#ifndef B200_PDL
#define B200_PDL 1
#endif
#include "icp_kernels.cuh"
#include "stream_kernels.cuh"
#include "normals.cuh"
#include "reduce.cuh"

using namespace b200;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(_e == cudaErrorMemoryAllocation ? B200ICP_ENOMEM : B200ICP_ECUDA,          \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
  } while (0)

// Device buffer.  With a stream it lives in the device's stream-ordered memory pool (cudaMallocAsync /
// cudaFreeAsync, release threshold raised in b200icp_create): building a scan allocates ~10 buffers, and the
// e2e path builds two scans per match -- pooled allocation keeps cudaMalloc/cudaFree (and the device-wide
// synchronisation cudaFree implies) off that path.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t count = 0;
  cudaStream_t stream = nullptr;
  bool pooled = false;
  cudaError_t alloc(size_t n) {
    release();
    count = n;
    if (n == 0) return cudaSuccess;
    return cudaMalloc((void**)&p, n * sizeof(T));
  }
  cudaError_t alloc_async(size_t n, cudaStream_t st) {
    release();
    count = n;
    if (n == 0) return cudaSuccess;
    stream = st;
    pooled = true;
    return cudaMallocAsync((void**)&p, n * sizeof(T), st);
  }
  cudaError_t ensure(size_t n) { return n <= count && p ? cudaSuccess : alloc(n); }
  cudaError_t ensure_async(size_t n, cudaStream_t st) { return n <= count && p ? cudaSuccess : alloc_async(n, st); }
  void release() {
    if (p) {
      if (pooled) cudaFreeAsync(p, stream);
      else cudaFree(p);
    }
    p = nullptr;
    count = 0;
    pooled = false;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

constexpr int kMaxProfileEvents = 4096;
constexpr int kMaxBlocksPerSm = 8;  // partial-sum workspace is sized sm_count * kMaxBlocksPerSm

}  // namespace

// LUM seed cache key: the scans' ids (process-wide, monotonically increasing, never reused), not their addresses --
// a scan allocated where a destroyed one lived can never pick up the old link's neighbours.
struct LinkKey {
  uint64_t first;
  uint64_t second;
  bool operator<(const LinkKey& o) const { return first != o.first ? first < o.first : second < o.second; }
};

struct b200icp_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t own = nullptr;
  // workspaces
  DevBuf<double> partials;          // [max_blocks][NS_MAX]
  DevBuf<IterState> d_state;
  IterState* h_state = nullptr;     // pinned, 2 slots (speculative chunk polling)
  DevBuf<double> rms_log;
  // neighbour caches of LUM links (first, second) -> int[second->n]: seeds for the link's next evaluation
  std::map<LinkKey, DevBuf<int>> lum_seeds;
  size_t lum_seed_bytes = 0, lum_seed_limit = (size_t)4 << 30;
  DevBuf<double> pose_log;          // [max_iter][16], see IterState::pose_log
  int pose_log_count = 0;           // iterations of the last match
  DevBuf<unsigned long long> npairs_log;
  DevBuf<unsigned> stage2_counter;
  DevBuf<unsigned> stage2_log;
  DevBuf<int> nn_cache;             // per data point: neighbour of the previous iteration
  DevBuf<float> nn_budget;          // per data point: motion budget left before it must search again
  DevBuf<double4> nn_pm;            // per data point: coordinates of the cached neighbour (w != 0: paired)
  DevBuf<int> queue;                // per data point slot: search queue, one segment per stream block
  DevBuf<unsigned> seg_count;       // [kMaxSegments] entries in each queue segment
  int split_ready = 0;              // kernel attributes of the two-kernel iteration are set
  int search_blocks_per_sm[4] = {0, 0, 0, 0};
  std::vector<double> prof_nn_ms, prof_solve_ms;   // last profiled match, per iteration
  std::vector<unsigned> prof_stage2;
  DevBuf<double> d_small;           // 64 doubles of scratch (transforms for API kernels)
  double* h_small = nullptr;        // pinned 64 doubles
  std::vector<cudaEvent_t> events;
  cudaEvent_t poll_ev[2] = {nullptr, nullptr};
  // query-sharded match: peer mailboxes over NVLink
  Mailbox* mailbox = nullptr;       // own, cudaMalloc'ed (IPC-exportable)
  CommDev comm = {0, 1, 0, {nullptr}};
  bool comm_ipc = false;
  unsigned long long comm_matches = 0;
  int blocks_per_sm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int blocks_per_sm_batch = 0;
};

static std::atomic<uint64_t> g_next_scan_id{1};

struct b200icp_scan {
  uint64_t id = g_next_scan_id.fetch_add(1);
  size_t n = 0;
  bool has_normals = false;
  GridDev g;                     // device pointers into the buffers below
  DevBuf<uint32_t> cell_start;
  DevBuf<float4> p32;
  DevBuf<double4> p64;
  DevBuf<double4> nrm;
  DevBuf<uint32_t> perm;
  uint64_t n_cells = 0, n_occupied = 0;
  double transMat[16];
  double dalignxf[16];
  double nmat[9];                // cumulative normal map (see IterState::Nm)
  // frees go to `st` (a live stream) instead of the stream the buffers were built on, which may be gone
  void retarget(cudaStream_t st) {
    cell_start.stream = st; p32.stream = st; p64.stream = st; nrm.stream = st; perm.stream = st;
  }
};

namespace {

double cells_for(const double ext[3], double h, int dims[3]) {
  double total = 1.0;
  for (int k = 0; k < 3; ++k) {
    double d = std::floor(ext[k] / h) + 1.0;
    if (d > 2.0e9) d = 2.0e9;
    dims[k] = (int)d;
    total *= d;
  }
  return total;
}

template <bool NAPX, bool PLANE, bool EXACT>
cudaError_t launch_iter(b200icp_ctx* ctx, int variant, const b200icp_scan* model, const b200icp_scan* data,
                        double maxdist2, int rnd, const CommDev& comm, int* grid_out) {
  auto kern = icp_iter_kernel<NAPX, PLANE, EXACT>;
  constexpr int NS = NAPX ? (int)NS_NAPX : (int)NS_P2P;
  const size_t smem = sizeof(IterSmem<NS>);
  if (ctx->blocks_per_sm[variant] == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kBlock, smem);
    if (e != cudaSuccess) return e;
    ctx->blocks_per_sm[variant] = std::max(b, 1);
  }
  const uint32_t nd = (uint32_t)data->n;
  const uint32_t ntiles = (nd + kBlock - 1) / kBlock;
  const int grid = (int)std::min<uint32_t>(
      ntiles, (uint32_t)(ctx->sm_count * std::min(ctx->blocks_per_sm[variant], kMaxBlocksPerSm)));
  *grid_out = grid;
#if B200_PDL
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, model->g, (const double4*)data->g.p64, (const double4*)data->g.nrm, nd,
                            ctx->d_state.p, maxdist2, rnd, ctx->nn_cache.p, ctx->nn_budget.p, ctx->partials.p,
                            ctx->stage2_counter.p, ctx->rms_log.p, ctx->npairs_log.p, ctx->stage2_log.p, comm);
#else
  kern<<<grid, kBlock, smem, ctx->stream>>>(model->g, data->g.p64, data->g.nrm, nd, ctx->d_state.p,
                                            maxdist2, rnd, ctx->nn_cache.p, ctx->nn_budget.p, ctx->partials.p,
                                            ctx->stage2_counter.p, ctx->rms_log.p, ctx->npairs_log.p,
                                            ctx->stage2_log.p, comm);
  return cudaSuccess;
#endif
}

int launch_iter_dispatch(b200icp_ctx* ctx, bool napx, bool plane, bool exact,
                         const b200icp_scan* model, const b200icp_scan* data, double maxdist2,
                         int rnd, const CommDev& comm, int* grid_out) {
  const int v = (napx ? 4 : 0) | (plane ? 2 : 0) | (exact ? 1 : 0);
  cudaError_t e;
  switch (v) {
    case 0: e = launch_iter<false, false, false>(ctx, v, model, data, maxdist2, rnd, comm, grid_out); break;
    case 1: e = launch_iter<false, false, true>(ctx, v, model, data, maxdist2, rnd, comm, grid_out); break;
    case 2: e = launch_iter<false, true, false>(ctx, v, model, data, maxdist2, rnd, comm, grid_out); break;
    case 3: e = launch_iter<false, true, true>(ctx, v, model, data, maxdist2, rnd, comm, grid_out); break;
    case 6: e = launch_iter<true, true, false>(ctx, v, model, data, maxdist2, rnd, comm, grid_out); break;
    case 7: e = launch_iter<true, true, true>(ctx, v, model, data, maxdist2, rnd, comm, grid_out); break;
    default: return -1;
  }
  return e == cudaSuccess ? 0 : -2;
}

int max_iter_grid(const b200icp_ctx* ctx) { return ctx->sm_count * kMaxBlocksPerSm; }

// ---- two-kernel iteration (stream_kernels.cuh): streaming pass over every point, then the queued searches + solve
template <bool PLANE, bool EXACT>
cudaError_t launch_split(b200icp_ctx* ctx, int variant, const b200icp_scan* model, const b200icp_scan* data,
                         double maxdist2, const CommDev& comm, cudaEvent_t mid) {
  auto kstream = icp_stream_kernel<PLANE>;
  auto ksearch = icp_search_kernel<PLANE, EXACT>;
  const size_t smem_a = sizeof(StreamSmem), smem_b = sizeof(SearchKSmem);
  if (ctx->search_blocks_per_sm[variant] == 0) {
    cudaError_t e = cudaFuncSetAttribute(kstream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ksearch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    if (e != cudaSuccess) return e;
    int b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, ksearch, kBlock, smem_b);
    if (e != cudaSuccess) return e;
    ctx->search_blocks_per_sm[variant] = std::max(b, 1);
  }
  const uint32_t nd = (uint32_t)data->n;
  const uint32_t ntiles = (nd + kBlock - 1) / kBlock;
  uint32_t nseg = std::min<uint32_t>(ntiles, std::min<uint32_t>((uint32_t)ctx->sm_count * 2u, (uint32_t)kMaxSegments));
  const uint32_t tiles_per_seg = (ntiles + nseg - 1) / nseg;
  nseg = (ntiles + tiles_per_seg - 1) / tiles_per_seg;
  const int grid_b = (int)std::min<uint32_t>(
      (ntiles + 0u), (uint32_t)(ctx->sm_count * std::min(ctx->search_blocks_per_sm[variant], kMaxBlocksPerSm - 2)));
  // both kernels are launched with programmatic stream serialization (PDL): launch latency and prologue of each
  // overlap the tail of its predecessor; the kernels order themselves with griddepcontrol.wait
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kBlock);
  cfg.stream = ctx->stream;
  static const bool no_pdl = getenv("B200ICP_NO_PDL") != nullptr;   // A/B switch
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  cfg.gridDim = dim3(nseg);
  cfg.dynamicSmemBytes = smem_a;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kstream, model->g, (const double4*)data->g.p64, (const double4*)data->g.nrm, nd,
                                     (const IterState*)ctx->d_state.p, maxdist2, (const double4*)ctx->nn_pm.p,
                                     ctx->nn_budget.p, ctx->queue.p, ctx->seg_count.p, tiles_per_seg, ctx->partials.p);
  if (e != cudaSuccess) return e;
  if (mid) cudaEventRecord(mid, ctx->stream);
  cfg.gridDim = dim3(grid_b);
  cfg.dynamicSmemBytes = smem_b;
  return cudaLaunchKernelEx(&cfg, ksearch, model->g, (const double4*)data->g.p64, (const double4*)data->g.nrm, nd,
                            ctx->d_state.p, maxdist2, ctx->nn_cache.p, ctx->nn_budget.p, ctx->nn_pm.p,
                            (const int*)ctx->queue.p, (const unsigned*)ctx->seg_count.p, (int)nseg, tiles_per_seg,
                            ctx->partials.p, ctx->stage2_counter.p, ctx->rms_log.p, ctx->npairs_log.p,
                            ctx->stage2_log.p, comm);
}

int launch_split_dispatch(b200icp_ctx* ctx, bool plane, bool exact, const b200icp_scan* model,
                          const b200icp_scan* data, double maxdist2, const CommDev& comm, cudaEvent_t mid) {
  const int v = (plane ? 2 : 0) | (exact ? 1 : 0);
  cudaError_t e;
  switch (v) {
    case 0: e = launch_split<false, false>(ctx, v, model, data, maxdist2, comm, mid); break;
    case 1: e = launch_split<false, true>(ctx, v, model, data, maxdist2, comm, mid); break;
    case 2: e = launch_split<true, false>(ctx, v, model, data, maxdist2, comm, mid); break;
    default: e = launch_split<true, true>(ctx, v, model, data, maxdist2, comm, mid); break;
  }
  return e == cudaSuccess ? 0 : -2;
}

}  // namespace

extern "C" {

const char* b200icp_last_error(void) { return g_last_error.c_str(); }
// internal: lets the library's other translation units (lum_graph.cpp) report through the same channel
int b200icp_set_error_(int code, const char* msg) { return fail(code, msg ? msg : ""); }

const char* b200icp_version(void) {
  return "b200icp 0.1 (sm_100a; exact fp64-verified grid NN; fused match; no CPU fallback)";
}

int b200icp_create(int device, b200icp_ctx** out) {
  if (!out) return fail(B200ICP_EINVAL, "b200icp_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(B200ICP_ENODEV, std::string("b200icp_create: no CUDA device (") +
                                    cudaGetErrorString(e) + "); this library has no CPU fallback");
  if (device < 0 || device >= count) return fail(B200ICP_EINVAL, "b200icp_create: bad device index");
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(B200ICP_ENODEV, std::string("b200icp_create: device '") + prop.name +
                                    "' is not sm_100; kernels are built for sm_100a only");
  CU_TRY(cudaSetDevice(device));
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;   // keep freed blocks cached in the pool
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  b200icp_ctx* ctx = new b200icp_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete ctx; return fail(B200ICP_ECUDA, cudaGetErrorString(e)); }
  ctx->stream = ctx->own;
  ctx->own_stream = true;
  bool ok = ctx->partials.alloc((size_t)max_iter_grid(ctx) * NS_MAX) == cudaSuccess &&
            ctx->d_state.alloc(1) == cudaSuccess && ctx->stage2_counter.alloc(4) == cudaSuccess &&
            ctx->d_small.alloc(64) == cudaSuccess &&
            cudaMallocHost((void**)&ctx->h_state, 2 * sizeof(IterState)) == cudaSuccess &&
            cudaMallocHost((void**)&ctx->h_small, 64 * sizeof(double)) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->poll_ev[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->poll_ev[1], cudaEventDisableTiming) == cudaSuccess &&
            cudaMemsetAsync(ctx->stage2_counter.p, 0, 4 * sizeof(unsigned), ctx->stream) == cudaSuccess;
  if (!ok) {
    std::string msg = cudaGetErrorString(cudaGetLastError());
    b200icp_destroy(ctx);
    return fail(B200ICP_ENOMEM, "b200icp_create: workspace allocation failed: " + msg);
  }
  *out = ctx;
  return B200ICP_OK;
}

void b200icp_destroy(b200icp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (cudaEvent_t ev : ctx->events) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i)
    if (ctx->poll_ev[i]) cudaEventDestroy(ctx->poll_ev[i]);
  if (ctx->h_state) cudaFreeHost(ctx->h_state);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  ctx->partials.release();
  ctx->d_state.release();
  ctx->rms_log.release();
  ctx->pose_log.release();
  ctx->npairs_log.release();
  ctx->stage2_counter.release();
  ctx->stage2_log.release();
  ctx->nn_cache.release();
  ctx->nn_budget.release();
  ctx->nn_pm.release();
  ctx->queue.release();
  ctx->seg_count.release();
  ctx->d_small.release();
  b200icp_comm_destroy(ctx);
  if (ctx->own) cudaStreamDestroy(ctx->own);
  delete ctx;
}

int b200icp_set_stream(b200icp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return fail(B200ICP_EINVAL, "ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
  else { ctx->stream = ctx->own; ctx->own_stream = true; }
  return B200ICP_OK;
}

int b200icp_synchronize(b200icp_ctx* ctx) {
  if (!ctx) return fail(B200ICP_EINVAL, "ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B200ICP_OK;
}

// --------------------------------------------------------------------------------------- scans
int b200icp_scan_create_device(b200icp_ctx* ctx, const double* d_xyz, const double* d_normals,
                               size_t n, double cell_edge, double max_dist_hint,
                               b200icp_scan** out) {
  if (!ctx || !out) return fail(B200ICP_EINVAL, "scan_create: NULL argument");
  *out = nullptr;
  if (n == 0) return fail(B200ICP_EEMPTY, "scan_create: cannot build a search grid over zero points");
  if (!d_xyz) return fail(B200ICP_EINVAL, "scan_create: xyz is NULL");
  if (n >= (1ull << 31)) return fail(B200ICP_EINVAL, "scan_create: more than 2^31-1 points");
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;

  // ---- bbox
  const int bb_blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 4);
  DevBuf<double> bb_part, bb_out;
  CU_TRY(bb_part.alloc_async((size_t)bb_blocks * 6, st));
  CU_TRY(bb_out.alloc_async(6, st));
  bbox_partial_kernel<<<bb_blocks, 256, 0, st>>>(d_xyz, n, bb_part.p);
  bbox_final_kernel<<<1, 32, 0, st>>>(bb_part.p, bb_blocks, bb_out.p);
  double bb[6];
  CU_TRY(cudaMemcpyAsync(bb, bb_out.p, sizeof bb, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  for (int k = 0; k < 6; ++k)
    if (!std::isfinite(bb[k])) return fail(B200ICP_EINVAL, "scan_create: non-finite coordinate in xyz");
  double ext[3] = {bb[3] - bb[0], bb[4] - bb[1], bb[5] - bb[2]};
  const double max_ext = std::max(ext[0], std::max(ext[1], ext[2]));

  // ---- cell edge
  int dims[3];
  double h_floor = std::max(max_ext * 1e-6, 1e-9);  // smallest edge whose dense table fits the cap
  while (cells_for(ext, h_floor, dims) > (double)kCellCap) h_floor *= 1.08;
  const char* env_ppc = getenv("B200ICP_TARGET_PPC");
  const double target_ppc = env_ppc ? std::max(atof(env_ppc), 0.25) : 4.5;   // (sweep on the final kernels, 1M/1M: 3.5 / 4 / 4.5 / 5 / 6 -> 7.09 / 7.13 / 7.02 / 7.07 / 7.12 ms)
  DevBuf<uint32_t> counts;
  DevBuf<unsigned long long> occ_dev;
  CU_TRY(occ_dev.alloc_async(1, st));
  GridShape gs;
  unsigned long long occ = 0;
  auto bin_points = [&](double h, uint32_t* keys, uint32_t* vals) -> cudaError_t {
    const double nc = cells_for(ext, h, dims);
    gs.h = h; gs.inv_h = 1.0 / h;
    gs.nx = dims[0]; gs.ny = dims[1]; gs.nz = dims[2];
    for (int k = 0; k < 3; ++k) gs.g0[k] = bb[k];
    cudaError_t e = counts.ensure_async((size_t)nc + 1, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(counts.p, 0, ((size_t)nc + 1) * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    cell_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyz, n, gs, keys, vals, counts.p);
    e = cudaMemsetAsync(occ_dev.p, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    count_occupied_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(counts.p, (size_t)nc, occ_dev.p);
    e = cudaMemcpyAsync(&occ, occ_dev.p, sizeof occ, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
  };
  double h;
  if (cell_edge > 0.0) {
    h = std::max(cell_edge, h_floor);
  } else {
    // density-driven choice: aim at `target_ppc` points per OCCUPIED cell (surface-like data: ppc ~ h^2)
    double vol = 1.0;
    for (int k = 0; k < 3; ++k) vol *= std::max(ext[k], max_ext * 1e-3 + 1e-9);
    h = std::max(h_floor, std::cbrt(vol / (double)n));
    const double h_hint_floor = max_dist_hint > 0.0 ? max_dist_hint / 64.0 : 0.0;  // bounds ring count
    h = std::max(h, h_hint_floor);
    for (int trial = 0; trial < 5; ++trial) {
      CU_TRY(bin_points(h, nullptr, nullptr));
      const double ppc = (double)n / (double)std::max<unsigned long long>(occ, 1);
      if (ppc > 0.75 * target_ppc && ppc < 1.35 * target_ppc) break;
      double f = std::sqrt(target_ppc / ppc);
      f = std::min(std::max(f, 0.25), 4.0);
      const double hn = std::max(std::max(h * f, h_floor), h_hint_floor);
      if (std::fabs(hn - h) < 1e-3 * h) break;
      h = hn;
    }
  }
  if (!(h > 0.0) || !std::isfinite(h)) return fail(B200ICP_EINVAL, "scan_create: bad cell edge");

  // ---- final binning + sort
  DevBuf<uint32_t> keys_in, keys_out, vals_in;
  b200icp_scan* sc = new b200icp_scan();
  auto bail = [&](int code, const std::string& msg) { delete sc; return fail(code, msg); };
  if (keys_in.alloc_async(n, st) != cudaSuccess || keys_out.alloc_async(n, st) != cudaSuccess ||
      vals_in.alloc_async(n, st) != cudaSuccess || sc->perm.alloc_async(n, st) != cudaSuccess)
    return bail(B200ICP_ENOMEM, "scan_create: device allocation failed (sort buffers)");
  cudaError_t e = bin_points(h, keys_in.p, vals_in.p);
  if (e != cudaSuccess) return bail(B200ICP_ECUDA, std::string("scan_create: binning: ") + cudaGetErrorString(e));
  const size_t ncells = (size_t)dims[0] * dims[1] * dims[2];
  int key_bits = 1;
  while ((1ull << key_bits) < ncells && key_bits < 32) ++key_bits;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.p, keys_out.p, vals_in.p, sc->perm.p,
                                  (int)n, 0, key_bits, st);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts.p, counts.p, (int)(ncells + 1), st);
  DevBuf<unsigned char> tmp;
  if (tmp.alloc_async(std::max(tmp_bytes, scan_bytes) + 256, st) != cudaSuccess ||
      sc->cell_start.alloc_async(ncells + 1, st) != cudaSuccess || sc->p32.alloc_async(n, st) != cudaSuccess ||
      sc->p64.alloc_async(n, st) != cudaSuccess || (d_normals && sc->nrm.alloc_async(n, st) != cudaSuccess))
    return bail(B200ICP_ENOMEM, "scan_create: device allocation failed (grid buffers)");
  cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.p, keys_out.p, vals_in.p, sc->perm.p, (int)n,
                                  0, key_bits, st);
  cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, counts.p, sc->cell_start.p, (int)(ncells + 1), st);
  const unsigned gblocks = (unsigned)((n + 255) / 256);
  DevBuf<float> bmax_part;
  if (bmax_part.alloc_async(gblocks, st) != cudaSuccess) return bail(B200ICP_ENOMEM, "scan_create: alloc");
  const double c[3] = {0.5 * (bb[0] + bb[3]), 0.5 * (bb[1] + bb[4]), 0.5 * (bb[2] + bb[5])};
  gather_kernel<<<gblocks, 256, 0, st>>>(d_xyz, d_normals, n, sc->perm.p, c[0], c[1], c[2], sc->p32.p,
                                         sc->p64.p, sc->nrm.p, bmax_part.p);
  std::vector<float> bm(gblocks);
  e = cudaMemcpyAsync(bm.data(), bmax_part.p, gblocks * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return bail(B200ICP_ECUDA, std::string("scan_create: ") + cudaGetErrorString(e));
  float bmax = 0.f;
  for (float v : bm) bmax = std::max(bmax, v);

  sc->n = n;
  sc->has_normals = d_normals != nullptr;
  sc->n_cells = ncells;
  sc->n_occupied = occ;
  GridDev& g = sc->g;
  for (int k = 0; k < 3; ++k) { g.g0[k] = bb[k]; g.c[k] = c[k]; g.bbox_lo[k] = bb[k]; g.bbox_hi[k] = bb[3 + k]; }
  g.h = h; g.inv_h = 1.0 / h;
  g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
  g.n = (uint32_t)n;
  g.bmax = bmax * 1.0000002f + 1e-30f;
  g.cell_start = sc->cell_start.p;
  g.p32 = sc->p32.p;
  g.p64 = sc->p64.p;
  g.nrm = sc->nrm.p;
  m4_identity(sc->transMat);
  m4_identity(sc->dalignxf);
  for (int i = 0; i < 9; ++i) sc->nmat[i] = (i % 4 == 0) ? 1.0 : 0.0;
  *out = sc;
  return B200ICP_OK;
}

int b200icp_scan_create(b200icp_ctx* ctx, const double* xyz, const double* normals, size_t n,
                        double cell_edge, double max_dist_hint, b200icp_scan** out) {
  if (!ctx || !out) return fail(B200ICP_EINVAL, "scan_create: NULL argument");
  *out = nullptr;
  if (n == 0) return fail(B200ICP_EEMPTY, "scan_create: cannot build a search grid over zero points");
  if (!xyz) return fail(B200ICP_EINVAL, "scan_create: xyz is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  DevBuf<double> d_xyz, d_nrm;
  CU_TRY(d_xyz.alloc_async(3 * n, ctx->stream));
  CU_TRY(cudaMemcpyAsync(d_xyz.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (normals) {
    CU_TRY(d_nrm.alloc_async(3 * n, ctx->stream));
    CU_TRY(cudaMemcpyAsync(d_nrm.p, normals, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  int rc = b200icp_scan_create_device(ctx, d_xyz.p, normals ? d_nrm.p : nullptr, n, cell_edge,
                                      max_dist_hint, out);
  cudaStreamSynchronize(ctx->stream);
  return rc;
}

void b200icp_scan_destroy(b200icp_ctx* ctx, b200icp_scan* scan) {
  if (!scan) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto it = ctx->lum_seeds.begin(); it != ctx->lum_seeds.end();) {   // seeds of links this scan was part of
      if (it->first.first == scan->id || it->first.second == scan->id) {
        ctx->lum_seed_bytes -= it->second.count * sizeof(int);
        it = ctx->lum_seeds.erase(it);
      } else {
        ++it;
      }
    }
    scan->retarget(ctx->stream);   // buffers return to the stream-ordered pool through a stream that is alive
  } else {
    cudaDeviceSynchronize();       // no context left (e.g. interpreter teardown): the legacy stream is always valid
    scan->retarget(nullptr);
  }
  delete scan;
}

size_t b200icp_scan_size(const b200icp_scan* scan) { return scan ? scan->n : 0; }

int b200icp_scan_grid_info(const b200icp_scan* scan, int dims[3], double* cell_edge, uint64_t* n_cells,
                           uint64_t* n_occupied) {
  if (!scan) return fail(B200ICP_EINVAL, "scan is NULL");
  if (dims) { dims[0] = scan->g.nx; dims[1] = scan->g.ny; dims[2] = scan->g.nz; }
  if (cell_edge) *cell_edge = scan->g.h;
  if (n_cells) *n_cells = scan->n_cells;
  if (n_occupied) *n_occupied = scan->n_occupied;
  return B200ICP_OK;
}

int b200icp_scan_get_pose(const b200icp_scan* scan, double transMat[16], double dalignxf[16]) {
  if (!scan) return fail(B200ICP_EINVAL, "scan is NULL");
  if (transMat) memcpy(transMat, scan->transMat, sizeof scan->transMat);
  if (dalignxf) memcpy(dalignxf, scan->dalignxf, sizeof scan->dalignxf);
  return B200ICP_OK;
}

int b200icp_scan_set_pose(b200icp_scan* scan, const double transMat[16], const double dalignxf[16]) {
  if (!scan) return fail(B200ICP_EINVAL, "scan is NULL");
  if (transMat) memcpy(scan->transMat, transMat, sizeof scan->transMat);
  if (dalignxf) {
    memcpy(scan->dalignxf, dalignxf, sizeof scan->dalignxf);
    // normals follow the rotation history (transform3normal); a freshly set pose restarts it with
    // the transposed rotation block of dalignxf, which is what a single Scan::transform would leave.
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) scan->nmat[3 * r + c] = dalignxf[4 * r + c];
  }
  return B200ICP_OK;
}

int b200icp_scan_transform(b200icp_scan* scan, const double alignxf[16]) {
  if (!scan || !alignxf) return fail(B200ICP_EINVAL, "scan_transform: NULL argument");
  double tmp[16];
  m4_mul(alignxf, scan->transMat, tmp);
  memcpy(scan->transMat, tmp, sizeof tmp);
  m4_mul(alignxf, scan->dalignxf, tmp);
  memcpy(scan->dalignxf, tmp, sizeof tmp);
  // transform3normal (globals.icc:1465-1475) multiplies by the transposed rotation block, as in solve_step_serial
  double nn[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      nn[3 * r + c] = alignxf[4 * r + 0] * scan->nmat[c] + alignxf[4 * r + 1] * scan->nmat[3 + c] +
                      alignxf[4 * r + 2] * scan->nmat[6 + c];
  memcpy(scan->nmat, nn, sizeof nn);
  return B200ICP_OK;
}

int b200icp_metascan_create(b200icp_ctx* ctx, const b200icp_scan* const* scans, int n_scans, double cell_edge,
                            double max_dist_hint, b200icp_scan** out) {
  if (!ctx || !scans || !out) return fail(B200ICP_EINVAL, "metascan_create: NULL argument");
  *out = nullptr;
  if (n_scans <= 0) return fail(B200ICP_EEMPTY, "metascan_create: no member scans");
  size_t total = 0;
  for (int i = 0; i < n_scans; ++i) {
    if (!scans[i]) return fail(B200ICP_EINVAL, "metascan_create: NULL member scan");
    total += scans[i]->n;
  }
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevBuf<double> d_xyz, d_xf;
  CU_TRY(d_xyz.alloc_async(3 * total, st));
  CU_TRY(d_xf.alloc_async((size_t)25 * n_scans, st));
  std::vector<double> hx((size_t)25 * n_scans);
  for (int i = 0; i < n_scans; ++i) {
    memcpy(&hx[(size_t)25 * i], scans[i]->dalignxf, 16 * sizeof(double));
    memcpy(&hx[(size_t)25 * i + 16], scans[i]->nmat, 9 * sizeof(double));
  }
  CU_TRY(cudaMemcpyAsync(d_xf.p, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  size_t off = 0;
  for (int i = 0; i < n_scans; ++i) {   // member i's current "xyz reduced", in its original row order
    const b200icp_scan* s = scans[i];
    scan_export_kernel<<<(unsigned)((s->n + 255) / 256), 256, 0, st>>>(s->g.p64, nullptr, s->perm.p, (uint32_t)s->n,
                                                                       d_xf.p + (size_t)25 * i, d_xyz.p + 3 * off,
                                                                       nullptr);
    off += s->n;
  }
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(st));   // hx must outlive the copy
  return b200icp_scan_create_device(ctx, d_xyz.p, nullptr, total, cell_edge, max_dist_hint, out);
}

int b200icp_scan_download(b200icp_ctx* ctx, const b200icp_scan* scan, double* xyz_out, double* nrm_out) {
  if (!ctx || !scan || !xyz_out) return fail(B200ICP_EINVAL, "scan_download: NULL argument");
  if (nrm_out && !scan->has_normals) return fail(B200ICP_ESTATE, "scan_download: scan has no normals");
  CU_TRY(cudaSetDevice(ctx->device));
  DevBuf<double> d_xyz, d_nrm, d_xf;
  CU_TRY(d_xyz.alloc_async(3 * scan->n, ctx->stream));
  if (nrm_out) CU_TRY(d_nrm.alloc_async(3 * scan->n, ctx->stream));
  CU_TRY(d_xf.alloc_async(25, ctx->stream));
  double hx[25];
  memcpy(hx, scan->dalignxf, 16 * sizeof(double));
  memcpy(hx + 16, scan->nmat, 9 * sizeof(double));
  CU_TRY(cudaMemcpyAsync(d_xf.p, hx, sizeof hx, cudaMemcpyHostToDevice, ctx->stream));
  scan_export_kernel<<<(unsigned)((scan->n + 255) / 256), 256, 0, ctx->stream>>>(
      scan->g.p64, scan->g.nrm, scan->perm.p, (uint32_t)scan->n, d_xf.p, d_xyz.p, nrm_out ? d_nrm.p : nullptr);
  CU_TRY(cudaMemcpyAsync(xyz_out, d_xyz.p, 3 * scan->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (nrm_out)
    CU_TRY(cudaMemcpyAsync(nrm_out, d_nrm.p, 3 * scan->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  CU_TRY(cudaGetLastError());
  return B200ICP_OK;
}

// ----------------------------------------------------------------------------- API search path
int b200icp_nn_batch_device(b200icp_ctx* ctx, const b200icp_scan* model, const double* d_q_xyz,
                            const double* d_q_nrm, size_t n, const double source_alignxf[16],
                            double maxdist2, int pairing_mode, int32_t* d_idx_out, double* d_d2_out,
                            double sums_out[8]) {
  if (!ctx || !model) return fail(B200ICP_EINVAL, "nn_batch: NULL argument");
  if (pairing_mode != B200ICP_CLOSEST_POINT && pairing_mode != B200ICP_CLOSEST_PLANE_SIMPLE)
    return fail(B200ICP_EINVAL, "nn_batch: pairing mode not on the accelerated path");
  if (pairing_mode == B200ICP_CLOSEST_PLANE_SIMPLE && !d_q_nrm)
    return fail(B200ICP_EINVAL, "nn_batch: CLOSEST_PLANE_SIMPLE needs query normals");
  if (!(maxdist2 >= 0.0)) return fail(B200ICP_EINVAL, "nn_batch: maxdist2 must be >= 0");
  if (sums_out) for (int k = 0; k < 8; ++k) sums_out[k] = 0.0;
  if (n == 0) return B200ICP_OK;
  if (!d_q_xyz) return fail(B200ICP_EINVAL, "nn_batch: q_xyz is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  double id[16];
  m4_identity(id);
  const double* S = source_alignxf ? source_alignxf : id;
  BatchXf xfs;   // by value: no staging buffer shared with later calls (the device-pointer form returns unsynchronised)
  memcpy(xfs.S, S, sizeof xfs.S);
  m4_inverse(S, xfs.Sinv);
  if (ctx->blocks_per_sm_batch == 0) {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, nn_batch_kernel<false>, kBlock, 0);
    ctx->blocks_per_sm_batch = std::max(b, 1);
  }
  const size_t ntiles = (n + kBlock - 1) / kBlock;
  const int grid = (int)std::min<size_t>(ntiles, (size_t)std::min(ctx->sm_count * ctx->blocks_per_sm_batch,
                                                                  max_iter_grid(ctx)));
  if (pairing_mode == B200ICP_CLOSEST_PLANE_SIMPLE)
    nn_batch_kernel<true><<<grid, kBlock, 0, ctx->stream>>>(model->g, d_q_xyz, d_q_nrm, n, xfs,
                                                            maxdist2, d_idx_out, d_d2_out, ctx->partials.p);
  else
    nn_batch_kernel<false><<<grid, kBlock, 0, ctx->stream>>>(model->g, d_q_xyz, d_q_nrm, n, xfs,
                                                             maxdist2, d_idx_out, d_d2_out, ctx->partials.p);
  CU_TRY(cudaGetLastError());
  if (sums_out) {
    std::vector<double> hp((size_t)grid * 8);
    CU_TRY(cudaMemcpyAsync(hp.data(), ctx->partials.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost,
                           ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < 8; ++k) sums_out[k] += hp[(size_t)b * 8 + k];
  }
  return B200ICP_OK;
}

int b200icp_nn_batch(b200icp_ctx* ctx, const b200icp_scan* model, const double* q_xyz,
                     const double* q_nrm, size_t n, const double source_alignxf[16], double maxdist2,
                     int pairing_mode, int32_t* idx_out, double* d2_out, double sums_out[8]) {
  if (!ctx || !model) return fail(B200ICP_EINVAL, "nn_batch: NULL argument");
  if (n == 0) {
    if (sums_out) for (int k = 0; k < 8; ++k) sums_out[k] = 0.0;
    return B200ICP_OK;
  }
  if (!q_xyz) return fail(B200ICP_EINVAL, "nn_batch: q_xyz is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  DevBuf<double> dq, dn, dd2;
  DevBuf<int32_t> didx;
  // stream-ordered pool (release threshold raised in b200icp_create): no cudaMalloc / cudaFree per call
  CU_TRY(dq.alloc_async(3 * n, ctx->stream));
  CU_TRY(didx.alloc_async(n, ctx->stream));
  if (d2_out) CU_TRY(dd2.alloc_async(n, ctx->stream));
  CU_TRY(cudaMemcpyAsync(dq.p, q_xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (q_nrm) {
    CU_TRY(dn.alloc_async(3 * n, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dn.p, q_nrm, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  int rc = b200icp_nn_batch_device(ctx, model, dq.p, q_nrm ? dn.p : nullptr, n, source_alignxf, maxdist2,
                                   pairing_mode, didx.p, d2_out ? dd2.p : nullptr, sums_out);
  if (rc != B200ICP_OK) { cudaStreamSynchronize(ctx->stream); return rc; }
  if (idx_out)
    CU_TRY(cudaMemcpyAsync(idx_out, didx.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (d2_out)
    CU_TRY(cudaMemcpyAsync(d2_out, dd2.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  CU_TRY(cudaGetLastError());
  return B200ICP_OK;
}

int b200icp_find_closest(b200icp_ctx* ctx, const b200icp_scan* model, const double p[3], double maxdist2,
                         int64_t* idx_out) {
  if (!p || !idx_out) return fail(B200ICP_EINVAL, "find_closest: NULL argument");
  int32_t idx = -1;
  int rc = b200icp_nn_batch(ctx, model, p, nullptr, 1, nullptr, maxdist2, B200ICP_CLOSEST_POINT, &idx,
                            nullptr, nullptr);
  if (rc != B200ICP_OK) return rc;
  *idx_out = idx;
  return B200ICP_OK;
}

// -------------------------------------------------------------------------------- fused match
int b200icp_match(b200icp_ctx* ctx, const b200icp_scan* model, b200icp_scan* data,
                  const b200icp_match_params* prm, double* rms_per_iter, uint64_t* npairs_per_iter,
                  b200icp_match_result* result) {
  if (!ctx || !model || !data || !prm) return fail(B200ICP_EINVAL, "match: NULL argument");
  const int algo = prm->algo;
  if (algo != 1 && algo != 2 && algo != 3 && algo != 4 && algo != 5 && algo != 6 && algo != 10)
    return fail(B200ICP_EINVAL, "match: algo must be 1 (QUAT), 2 (SVD), 3 (ORTHO), 4 (DUAL), 5 (HELIX), 6 (APX) or 10 (NAPX)");
  if (prm->pairing_mode != B200ICP_CLOSEST_POINT && prm->pairing_mode != B200ICP_CLOSEST_PLANE_SIMPLE)
    return fail(B200ICP_EINVAL, "match: pairing mode not on the accelerated path");
  // reference ctor checks, icp6D.cc:67-78 (there: exit(1))
  if (prm->max_dist_match < 0.0) return fail(B200ICP_EINVAL, "match: max_dist_match has to be >= 0");
  if (prm->max_num_iterations < 0) return fail(B200ICP_EINVAL, "match: max_num_iterations has to be >= 0");
  const bool plane = prm->pairing_mode == B200ICP_CLOSEST_PLANE_SIMPLE;
  const bool napx = algo == 10;
  if (plane && !data->has_normals) return fail(B200ICP_ESTATE, "match: CLOSEST_PLANE_SIMPLE needs data normals");
  if (napx && !plane) return fail(B200ICP_EINVAL, "match: NAPX consumes pair normals; use CLOSEST_PLANE_SIMPLE");
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int max_iter = prm->max_num_iterations;
  b200icp_match_result res;
  memset(&res, 0, sizeof res);
  res.queries = data->n;
  if (max_iter == 0) {  // icp6D.cc:112-114
    if (result) *result = res;
    return B200ICP_OK;
  }
  CU_TRY(ctx->rms_log.ensure((size_t)max_iter));
  CU_TRY(ctx->pose_log.ensure(16 * (size_t)max_iter));
  ctx->pose_log_count = 0;
  CU_TRY(ctx->npairs_log.ensure((size_t)max_iter));
  CU_TRY(ctx->stage2_log.ensure(2 * (size_t)max_iter));
  CU_TRY(ctx->nn_cache.ensure(data->n));
  CU_TRY(ctx->nn_budget.ensure(data->n));
  // Two-kernel iteration (stream_kernels.cuh: TMA-streamed pass over every point + queued searches), available for
  // the point-to-point moment set when every point is visited every iteration.  Opt-in (B200ICP_SPLIT=1): measured
  // on the 1M/1M bench pair it is 2.6 % slower per match than the single fused kernel (8.46 vs 8.24 ms; faster
  // only once < 20 % of the points still search, profiles/r01_split_vs_fused.md), so the fused kernel stays the
  // default.  Read per call so that tests can exercise both paths in one process.
  const char* env_split = getenv("B200ICP_SPLIT");
  const bool split = !napx && prm->rnd <= 1 && env_split && env_split[0] == '1';
  if (split) {
    CU_TRY(ctx->nn_pm.ensure(data->n));
    CU_TRY(ctx->queue.ensure(((data->n + kBlock - 1) / kBlock) * kBlock));
    CU_TRY(ctx->seg_count.ensure(kMaxSegments));
  }

  IterState* hs = ctx->h_state;
  memset(hs, 0, sizeof(IterState));
  memcpy(hs->X, data->dalignxf, sizeof hs->X);
  memcpy(hs->Xprev, data->dalignxf, sizeof hs->Xprev);
  memcpy(hs->T, data->transMat, sizeof hs->T);
  memcpy(hs->S, model->dalignxf, sizeof hs->S);
  m4_inverse(hs->S, hs->Sinv);
  memcpy(hs->Nm, data->nmat, sizeof hs->Nm);
  xf_point(hs->S, model->g.c, hs->o);
  m4_identity(hs->alignxf);
  hs->eps = prm->epsilon_icp;
  hs->algo = algo;
  hs->napx_weighted = prm->napx_weighted;
  hs->max_iter = max_iter;
  hs->pose_log = ctx->pose_log.p;
  // fixed-point scales of the point-to-point moment sums (FixAcc, icp_kernels.cuh).  Every addend is bounded through
  // |p - o| <= R for both points of a pair (o = the model's bbox centre; a data point pairs only within maxdist of
  // a model point) and |p1 - p2| < maxdist; R, D carry a factor 2 of head room.
  {
    double hd = 0.0;
    for (int k = 0; k < 3; ++k) {
      const double e = 0.5 * (model->g.bbox_hi[k] - model->g.bbox_lo[k]);
      hd += e * e;
    }
    const double md = prm->max_dist_match;
    const double R = 2.0 * (std::sqrt(hd) + 2.0 * md) + 1.0, D = 2.0 * md + 1e-3;
    double bound[NS_P2P];
    for (int k = 0; k < NS_P2P; ++k) bound[k] = R * R;
    bound[MP_N] = 2.0;
    bound[MP_D2] = D * D;
    for (int k = 0; k < 3; ++k) bound[MP_M + k] = bound[MP_D + k] = R;
    for (int k = 0; k < NS_P2P; ++k) {
      int ex = 0, ex1 = 0;
      std::frexp(bound[k] * (double)std::max<size_t>(data->n, 1), &ex);   // nd * bound < 2^ex
      std::frexp(bound[k], &ex1);                                          //      bound < 2^ex1
      const int sh = std::min(60 - ex, 50 - ex1);    // totals below 2^60, every addend below 2^50 (FixRef)
      hs->fix_s[k] = std::ldexp(1.0, sh);
      hs->fix_inv[k] = std::ldexp(1.0, -sh);
    }
    hs->fixed_point = 1;
  }
  CU_TRY(cudaMemcpyAsync(ctx->d_state.p, hs, sizeof(IterState), cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemsetAsync(ctx->partials.p, 0, NS_MAX * sizeof(double), st));   // row of integer totals (fixed-point kernels)
  CU_TRY(cudaMemsetAsync(ctx->nn_cache.p, 0xFF, data->n * sizeof(int), st));  // -1: no cached neighbour
  CU_TRY(cudaMemsetAsync(ctx->nn_budget.p, 0, data->n * sizeof(float), st));

  const double maxdist2 = prm->max_dist_match * prm->max_dist_match;
  const bool exact = prm->exact != 0;
  CommDev comm = {0, 1, 0, {nullptr}};
  if (prm->sharded) {
    if (ctx->comm.world < 2) return fail(B200ICP_ESTATE, "match: params.sharded set but no communicator is connected");
    if (prm->rnd > 1) return fail(B200ICP_EINVAL, "match: sharded match does not support rnd > 1");
    comm = ctx->comm;
    comm.seq_base = (++ctx->comm_matches) << 24;   // every rank calls match the same number of times
  }
  const bool profile = prm->profile != 0;
  if (profile) {
    const size_t need = std::min<size_t>((size_t)max_iter * 3 + 3, kMaxProfileEvents);
    while (ctx->events.size() < need) {
      cudaEvent_t ev;
      CU_TRY(cudaEventCreate(&ev));
      ctx->events.push_back(ev);
    }
  }
  const int chunk = 4;
  int launched = 0, grid = 0;
  uint32_t launches = 0;
  int slot = 0;
  bool done = false;
  IterState* h_slots = ctx->h_state;  // slot 0 doubles as the upload buffer; copy is ordered after it
  int pending = -1;
  size_t ev_i = 0;
  while (!done && launched < max_iter) {
    const int todo = std::min(chunk, max_iter - launched);
    for (int k = 0; k < todo; ++k) {
      const bool rec = profile && ev_i + 3 <= ctx->events.size();
      if (rec) CU_TRY(cudaEventRecord(ctx->events[ev_i], st));
      if (split) {
        // events: [0] start, [1] between the stream and the search kernel, [2] end
        if (launch_split_dispatch(ctx, plane, exact, model, data, maxdist2, comm, rec ? ctx->events[ev_i + 1] : nullptr) != 0)
          return fail(B200ICP_ECUDA, std::string("match: launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        if (rec) { CU_TRY(cudaEventRecord(ctx->events[ev_i + 2], st)); ev_i += 3; }
        launches += 2;
        continue;
      }
      if (launch_iter_dispatch(ctx, napx, plane, exact, model, data, maxdist2, prm->rnd, comm, &grid) != 0)
        return fail(B200ICP_EINVAL, "match: unsupported kernel variant");
      if (rec) { CU_TRY(cudaEventRecord(ctx->events[ev_i + 1], st)); CU_TRY(cudaEventRecord(ctx->events[ev_i + 2], st)); ev_i += 3; }
      launches += 1;
    }
    launched += todo;
    CU_TRY(cudaGetLastError());
    // poll the previous chunk's state while this chunk runs (speculative launch hides the round trip)
    if (pending >= 0) {
      CU_TRY(cudaEventSynchronize(ctx->poll_ev[pending]));
      if (h_slots[pending].done) done = true;
    }
    CU_TRY(cudaMemcpyAsync(&h_slots[slot], ctx->d_state.p, sizeof(IterState), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(ctx->poll_ev[slot], st));
    pending = slot;
    slot ^= 1;
  }
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaGetLastError());
  const IterState& fin = h_slots[pending];
  if (fin.done && fin.ret_iter <= -1000)
    return fail(B200ICP_ECUDA, "match: sharded match timed out waiting for a peer's moments");
  res.iterations = fin.done ? fin.ret_iter : max_iter;
  res.iterations_run = fin.iters_run;
  ctx->pose_log_count = fin.iters_run;
  res.kernel_launches = launches;
  res.stage2_queries_last = fin.stage2_last;
  if (fin.iters_run > 0) {
    std::vector<double> rms(fin.iters_run);
    std::vector<unsigned long long> np(fin.iters_run);
    CU_TRY(cudaMemcpy(rms.data(), ctx->rms_log.p, rms.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(np.data(), ctx->npairs_log.p, np.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    res.rms_last = rms.back();
    res.npairs_last = np.back();
    for (int i = 0; i < fin.iters_run; ++i) {
      if (rms_per_iter) rms_per_iter[i] = rms[i];
      if (npairs_per_iter) npairs_per_iter[i] = np[i];
    }
  }
  ctx->prof_nn_ms.clear(); ctx->prof_solve_ms.clear(); ctx->prof_stage2.clear();
  if (fin.iters_run > 0) {
    ctx->prof_stage2.resize(2 * (size_t)fin.iters_run);
    CU_TRY(cudaMemcpy(ctx->prof_stage2.data(), ctx->stage2_log.p, 2 * (size_t)fin.iters_run * sizeof(unsigned),
                      cudaMemcpyDeviceToHost));
  }
  if (profile && ev_i >= 3) {
    double nn_ms = 0, sv_ms = 0;
    int cnt = 0;
    const int lim = std::min<int>(fin.iters_run, (int)(ev_i / 3));
    for (int i = 0; i < lim; ++i) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, ctx->events[3 * i], ctx->events[3 * i + 1]);
      cudaEventElapsedTime(&b, ctx->events[3 * i + 1], ctx->events[3 * i + 2]);
      // fused kernel: a = the iteration kernel, b = nothing (the solve runs in its last block).
      // two-kernel iteration: a = streaming kernel, b = search kernel (+ solve) -> report the search kernel as
      // "nn" and the streaming kernel in the second slot
      if (split) std::swap(a, b);
      ctx->prof_nn_ms.push_back(a); ctx->prof_solve_ms.push_back(b);
      nn_ms += a; sv_ms += b; ++cnt;
    }
    if (cnt) { res.nn_kernel_ms = nn_ms / cnt; res.solve_kernel_ms = sv_ms / cnt; }
  }
  memcpy(data->dalignxf, fin.X, sizeof fin.X);
  memcpy(data->transMat, fin.T, sizeof fin.T);
  memcpy(data->nmat, fin.Nm, sizeof fin.Nm);
  if (result) *result = res;
  return B200ICP_OK;
}

#if defined(B200_TILE_STATS) || defined(B200_COUNT_UNSETTLED)
// debug builds only: counters of the cooperative tile search (or, with B200_COUNT_UNSETTLED, of the fp32-ambiguous searches) (batches tiled, staged points, batches refused); reset on read
int b200icp_debug_tile_stats(unsigned long long* out8) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out8, g_tile_stats, sizeof z) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(g_tile_stats, z, sizeof z) == cudaSuccess ? 0 : -1;
}
#endif

#ifdef B200_TIMING
// debug builds only: per-block (start, tile loop done, partials stored, SM id) of the last fused iteration launch
int b200icp_debug_blocks(unsigned long long* out, int nblocks) {
  return cudaMemcpyFromSymbol(out, g_blk, (size_t)nblocks * 4 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -1;
}
// debug builds only: per-warp (walk start, walk end, leftover start, leftover end, search batches, -) of the last launch
int b200icp_debug_warps(unsigned long long* out, int nblocks) {
  return cudaMemcpyFromSymbol(out, g_wrp, (size_t)nblocks * 48 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -1;
}
// debug builds only: the in-kernel timeline of the last iteration (globaltimer ns), see tl_mark
int b200icp_debug_timing(unsigned long long* out32) {
  return cudaMemcpyFromSymbol(out32, g_tl, 32 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -1;
}
#endif

int b200icp_last_poses(b200icp_ctx* ctx, int cap, double* transmats) {
  if (!ctx) return fail(B200ICP_EINVAL, "ctx is NULL");
  const int n = ctx->pose_log_count;
  const int m = std::min(n, cap);
  if (m > 0 && transmats) {
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaMemcpy(transmats, ctx->pose_log.p, 16 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return n;
}

int b200icp_last_profile(b200icp_ctx* ctx, int cap, double* nn_ms, double* solve_ms, uint32_t* stage2,
                         uint32_t* searches) {
  if (!ctx) return fail(B200ICP_EINVAL, "ctx is NULL");
  const int n = (int)(ctx->prof_stage2.size() / 2);
  for (int i = 0; i < n && i < cap; ++i) {
    if (nn_ms) nn_ms[i] = i < (int)ctx->prof_nn_ms.size() ? ctx->prof_nn_ms[i] : 0.0;
    if (solve_ms) solve_ms[i] = i < (int)ctx->prof_solve_ms.size() ? ctx->prof_solve_ms[i] : 0.0;
    if (stage2) stage2[i] = ctx->prof_stage2[2 * i];
    if (searches) searches[i] = ctx->prof_stage2[2 * i + 1];
  }
  return n;
}

// ------------------------------------------------------------------------------------ normals
int b200icp_normals_knn(b200icp_ctx* ctx, const double* xyz, size_t n, int k, const double rPos[3],
                        double* normals_out) {
  if (!ctx || !xyz || !rPos || !normals_out) return fail(B200ICP_EINVAL, "normals_knn: NULL argument");
  if (k < 1 || k > 32) return fail(B200ICP_EINVAL, "normals_knn: k must be in [1, 32]");
  if (n == 0) return fail(B200ICP_EEMPTY, "normals_knn: zero points");
  b200icp_scan* sc = nullptr;
  int rc = b200icp_scan_create(ctx, xyz, nullptr, n, 0.0, 0.0, &sc);
  if (rc != B200ICP_OK) return rc;
  DevBuf<double> d_out;
  cudaError_t e = d_out.alloc(3 * n);
  if (e == cudaSuccess) {
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (k <= 16)
      normals_knn_kernel<16><<<blocks, 128, 0, ctx->stream>>>(sc->g, sc->perm.p, k, rPos[0], rPos[1], rPos[2], d_out.p);
    else
      normals_knn_kernel<32><<<blocks, 128, 0, ctx->stream>>>(sc->g, sc->perm.p, k, rPos[0], rPos[1], rPos[2], d_out.p);
    e = cudaMemcpyAsync(normals_out, d_out.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  b200icp_scan_destroy(ctx, sc);
  if (e != cudaSuccess) return fail(B200ICP_ECUDA, std::string("normals_knn: ") + cudaGetErrorString(e));
  return B200ICP_OK;
}

int b200icp_scan_calc_normals(b200icp_ctx* ctx, b200icp_scan* scan, int k, const double rPos[3]) {
  if (!ctx || !scan || !rPos) return fail(B200ICP_EINVAL, "scan_calc_normals: NULL argument");
  if (k < 1 || k > 32) return fail(B200ICP_EINVAL, "scan_calc_normals: k must be in [1, 32]");
  CU_TRY(cudaSetDevice(ctx->device));
  // normals live in the frame the grid was built in ("normal reduced" of the un-moved scan), like uploaded ones;
  // the cumulative normal map (nmat) keeps applying on load
  if (!scan->nrm.p || scan->nrm.count < scan->n) {
    scan->nrm.stream = ctx->stream;
    CU_TRY(scan->nrm.alloc(scan->n));
    scan->g.nrm = scan->nrm.p;
  }
  const unsigned blocks = (unsigned)((scan->n + 127) / 128);
  if (k <= 16)
    normals_knn_kernel<16><<<blocks, 128, 0, ctx->stream>>>(scan->g, nullptr, k, rPos[0], rPos[1], rPos[2],
                                                            reinterpret_cast<double*>(scan->nrm.p));
  else
    normals_knn_kernel<32><<<blocks, 128, 0, ctx->stream>>>(scan->g, nullptr, k, rPos[0], rPos[1], rPos[2],
                                                            reinterpret_cast<double*>(scan->nrm.p));
  CU_TRY(cudaGetLastError());
  scan->has_normals = true;
  return B200ICP_OK;
}

// ------------------------------------------------------------------------------------ LUM link
extern "C++" {
namespace {
// Gaussian elimination with partial pivoting, N <= 7 (the reference inverts MM with newmat's .i())
template <int N>
bool gauss_solve(double (&A)[N][N], double* b) {
  for (int c = 0; c < N; ++c) {
    int piv = c;
    for (int r = c + 1; r < N; ++r) if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
    if (A[piv][c] == 0.0) return false;
    if (piv != c) { for (int k = 0; k < N; ++k) std::swap(A[c][k], A[piv][k]); std::swap(b[c], b[piv]); }
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int k = c; k < N; ++k) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  for (int r = N - 1; r >= 0; --r) {
    double t = b[r];
    for (int k = r + 1; k < N; ++k) t -= A[r][k] * b[k];
    b[r] = t / A[r][r];
  }
  return true;
}

// shared body of b200icp_lum_link (Euler, 6 parameters) and b200icp_lum_link_quat (7 parameters)
template <bool QUAT>
int lum_link_impl(b200icp_ctx* ctx, const b200icp_scan* first, const b200icp_scan* second, double max_dist_match2,
                  double* C, double* CD, uint64_t* npairs) {
  constexpr int N = QUAT ? 7 : 6;
  if (!ctx || !first || !second || !C || !CD) return fail(B200ICP_EINVAL, "lum_link: NULL argument");
  if (!(max_dist_match2 >= 0.0)) return fail(B200ICP_EINVAL, "lum_link: max_dist_match2 must be >= 0");
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (int i = 0; i < N * N; ++i) C[i] = 0.0;
  for (int i = 0; i < N; ++i) CD[i] = 0.0;
  if (npairs) *npairs = 0;
  // per-link neighbour cache: the pairs of this link's previous evaluation seed the searches of this one (the graph
  // relaxation evaluates every link once per iteration and the poses move little in between)
  int* cache = nullptr;
  int seeded = 0;
  {
    const LinkKey key{first->id, second->id};
    auto it = ctx->lum_seeds.find(key);
    if (it != ctx->lum_seeds.end() && it->second.count == second->n) {
      cache = it->second.p;
      seeded = 1;
    } else if (ctx->lum_seed_limit > 0) {
      const size_t need = second->n * sizeof(int);
      if (ctx->lum_seed_bytes + need > ctx->lum_seed_limit) { ctx->lum_seeds.clear(); ctx->lum_seed_bytes = 0; }
      if (need <= ctx->lum_seed_limit) {
        DevBuf<int>& b = ctx->lum_seeds[key];
        if (b.alloc(second->n) == cudaSuccess) {
          cache = b.p;
          ctx->lum_seed_bytes += need;
        } else {
          cudaGetLastError();
          ctx->lum_seeds.erase(key);
        }
      }
    }
  }
  if (!cache) {
    CU_TRY(ctx->nn_cache.ensure(second->n));
    cache = ctx->nn_cache.p;
  }
  double* hx = ctx->h_small;   // pinned, 64 doubles
  memcpy(hx, second->dalignxf, 16 * sizeof(double));
  memcpy(hx + 16, first->dalignxf, 16 * sizeof(double));
  m4_inverse(first->dalignxf, hx + 32);
  for (int i = 0; i < 7; ++i) hx[48 + i] = 0.0;
  CU_TRY(cudaMemcpyAsync(ctx->d_small.p, hx, 55 * sizeof(double), cudaMemcpyHostToDevice, st));
  const uint32_t nd = (uint32_t)second->n;
  const uint32_t ntiles = (nd + kBlock - 1) / kBlock;
  const int grid = (int)std::min<uint32_t>(ntiles, (uint32_t)(ctx->sm_count * 2));
  std::vector<double> hp((size_t)grid * kLumSums);
  lum_link_kernel<1, QUAT><<<grid, kBlock, 0, st>>>(first->g, second->g.p64, nd, ctx->d_small.p, max_dist_match2,
                                                    cache, ctx->partials.p, seeded);
  CU_TRY(cudaMemcpyAsync(hp.data(), ctx->partials.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaGetLastError());
  double s[kLumSums] = {0};
  for (int b = 0; b < grid; ++b)
    for (int k = 0; k < kLumSums; ++k) s[k] += hp[(size_t)b * kLumSums + k];
  const double m = s[0];
  if (npairs) *npairs = (uint64_t)(m + 0.5);
  if (!(m > 2.0)) return B200ICP_OK;   // "This case should not occur": C = CD = 0 (lum6Deuler.cc:243-259)
  const double sx = s[1], sy = s[2], sz = s[3], xpy = s[4], xpz = s[5], ypz = s[6], xy = s[7], xz = s[8], yz = s[9];
  double MM[N][N] = {{0}}, MZ[N];
  MM[0][0] = MM[1][1] = MM[2][2] = m;
  if (!QUAT) {                         // lum6Deuler.cc:177-191
    for (int i = 0; i < 6; ++i) MZ[i] = s[10 + i];
    MM[3][3] = ypz; MM[4][4] = xpy; MM[5][5] = xpz;
    MM[0][4] = MM[4][0] = -sy; MM[0][5] = MM[5][0] = sz;
    MM[1][3] = MM[3][1] = -sz; MM[1][4] = MM[4][1] = sx;
    MM[2][3] = MM[3][2] = sy;  MM[2][5] = MM[5][2] = -sx;
    MM[3][4] = MM[4][3] = -xz; MM[3][5] = MM[5][3] = -xy; MM[4][5] = MM[5][4] = -yz;
  } else {                             // lum6Dquat.cc:166-188 (1-based there)
    MZ[0] = s[10]; MZ[1] = s[11]; MZ[2] = s[12]; MZ[3] = s[16]; MZ[4] = s[13]; MZ[5] = s[14]; MZ[6] = s[15];
    MM[3][3] = s[17]; MM[4][4] = ypz; MM[5][5] = xpz; MM[6][6] = xpy;
    MM[0][3] = MM[3][0] = sx; MM[0][5] = MM[5][0] = -sz; MM[0][6] = MM[6][0] = sy;
    MM[1][3] = MM[3][1] = sy; MM[1][4] = MM[4][1] = sz;  MM[1][6] = MM[6][1] = -sx;
    MM[2][3] = MM[3][2] = sz; MM[2][4] = MM[4][2] = -sy; MM[2][5] = MM[5][2] = sx;
    MM[4][5] = MM[5][4] = -xy; MM[4][6] = MM[6][4] = -xz; MM[5][6] = MM[6][5] = -yz;
  }
  double A[N][N], D[N];
  memcpy(A, MM, sizeof A);
  memcpy(D, MZ, sizeof D);
  if (!gauss_solve<N>(A, D)) return B200ICP_OK;
  for (int i = 0; i < 7; ++i) hx[48 + i] = i < N ? D[i] : 0.0;
  CU_TRY(cudaMemcpyAsync(ctx->d_small.p + 48, hx + 48, 7 * sizeof(double), cudaMemcpyHostToDevice, st));
  lum_link_kernel<2, QUAT><<<grid, kBlock, 0, st>>>(first->g, second->g.p64, nd, ctx->d_small.p, max_dist_match2,
                                                    cache, ctx->partials.p, 0);
  CU_TRY(cudaMemcpyAsync(hp.data(), ctx->partials.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaGetLastError());
  double ss = 0.0;
  for (int b = 0; b < grid; ++b) ss += hp[(size_t)b * kLumSums];
  ss = ss / (2.0 * m - 3.0);
  if (!QUAT && ss < 0.0000000000001) return B200ICP_OK;   // identical clouds (lum6Deuler.cc:219-231; the quaternion
  ss = 1.0 / ss;                                           // form has no such guard, lum6Dquat.cc:208-209)
  for (int i = 0; i < N; ++i) {
    CD[i] = MZ[i] * ss;
    for (int k = 0; k < N; ++k) C[N * i + k] = MM[i][k] * ss;
  }
  return B200ICP_OK;
}
}  // namespace
}  // extern "C++"

int b200icp_lum_seed_cache(b200icp_ctx* ctx, size_t limit_bytes) {
  if (!ctx) return fail(B200ICP_EINVAL, "ctx is NULL");
  CU_TRY(cudaSetDevice(ctx->device));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->lum_seeds.clear();
  ctx->lum_seed_bytes = 0;
  ctx->lum_seed_limit = limit_bytes;
  return B200ICP_OK;
}

int b200icp_lum_link(b200icp_ctx* ctx, const b200icp_scan* first, const b200icp_scan* second,
                     double max_dist_match2, double C[36], double CD[6], uint64_t* npairs) {
  return lum_link_impl<false>(ctx, first, second, max_dist_match2, C, CD, npairs);
}

int b200icp_lum_link_quat(b200icp_ctx* ctx, const b200icp_scan* first, const b200icp_scan* second,
                          double max_dist_match2, double C[49], double CD[7], uint64_t* npairs) {
  return lum_link_impl<true>(ctx, first, second, max_dist_match2, C, CD, npairs);
}

// ------------------------------------------------------------------------------------ octree reduction
int b200icp_reduce_octree_center(b200icp_ctx* ctx, const double* xyz, size_t n, double voxel_size,
                                 double* xyz_out, size_t* n_out) {
  if (!ctx || !xyz || !xyz_out || !n_out) return fail(B200ICP_EINVAL, "reduce_octree_center: NULL argument");
  if (!(voxel_size > 0.0)) return fail(B200ICP_EINVAL, "reduce_octree_center: voxel_size must be > 0");
  *n_out = 0;
  if (n == 0) return B200ICP_OK;
  if (n >= (1ull << 31)) return fail(B200ICP_EINVAL, "reduce_octree_center: more than 2^31-1 points");
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevBuf<double> d_xyz, bb_part, bb_out, d_out;
  CU_TRY(d_xyz.alloc_async(3 * n, st));
  CU_TRY(cudaMemcpyAsync(d_xyz.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  const int bb_blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 4);
  CU_TRY(bb_part.alloc_async((size_t)bb_blocks * 6, st));
  CU_TRY(bb_out.alloc_async(6, st));
  bbox_partial_kernel<<<bb_blocks, 256, 0, st>>>(d_xyz.p, n, bb_part.p);
  bbox_final_kernel<<<1, 32, 0, st>>>(bb_part.p, bb_blocks, bb_out.p);
  double bb[6];
  CU_TRY(cudaMemcpyAsync(bb, bb_out.p, sizeof bb, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  for (int k = 0; k < 6; ++k)
    if (!std::isfinite(bb[k])) return fail(B200ICP_EINVAL, "reduce_octree_center: non-finite coordinate");
  OctRoot root;
  for (int k = 0; k < 3; ++k) root.c[k] = 0.5 * (bb[k] + bb[3 + k]);             // Boctree.h:249-251
  root.size = std::max(std::max(0.5 * (bb[3] - bb[0]), 0.5 * (bb[4] - bb[1])), 0.5 * (bb[5] - bb[2]));
  root.size += 1.0;                                                              // Boctree.h:255
  root.levels = 1;
  for (double hs = root.size / 2.0; hs > voxel_size; hs /= 2.0) ++root.levels;   // leaf: child half-size <= voxel
  if (root.levels > 21) return fail(B200ICP_EINVAL, "reduce_octree_center: voxel too small for a 63-bit octree key");
  DevBuf<unsigned long long> k_in, k_sorted, k_uniq, d_count;
  DevBuf<unsigned char> tmp;
  CU_TRY(k_in.alloc_async(n, st));
  CU_TRY(k_sorted.alloc_async(n, st));
  CU_TRY(k_uniq.alloc_async(n, st));
  CU_TRY(d_count.alloc_async(1, st));
  oct_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyz.p, n, root, k_in.p);
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, b1, k_in.p, k_sorted.p, (int)n, 0, 3 * root.levels, st);
  cub::DeviceSelect::Unique(nullptr, b2, k_sorted.p, k_uniq.p, d_count.p, (int)n, st);
  CU_TRY(tmp.alloc_async(std::max(b1, b2) + 256, st));
  cub::DeviceRadixSort::SortKeys(tmp.p, b1, k_in.p, k_sorted.p, (int)n, 0, 3 * root.levels, st);
  cub::DeviceSelect::Unique(tmp.p, b2, k_sorted.p, k_uniq.p, d_count.p, (int)n, st);
  unsigned long long m = 0;
  CU_TRY(cudaMemcpyAsync(&m, d_count.p, sizeof m, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  // DeviceSelect::Unique writes an int-sized count for int num_items
  m &= 0xFFFFFFFFull;
  if (m == 0 || m > n) return fail(B200ICP_ECUDA, "reduce_octree_center: bad unique count");
  CU_TRY(d_out.alloc_async(3 * (size_t)m, st));
  oct_centre_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(k_uniq.p, (size_t)m, root, d_out.p);
  CU_TRY(cudaMemcpyAsync(xyz_out, d_out.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaGetLastError());
  *n_out = (size_t)m;
  return B200ICP_OK;
}

// glibc's rand() stream (random_r.c TYPE_3: r[i] = r[i-3] + r[i-31] over 32-bit words, seeded by the Lehmer
// generator 16807 x mod 2^31-1, first 310 outputs discarded, result >> 1) -- what std::rand() returns on the Linux
// builds of the reference; `skip` values are dropped first (calls the process made before the reduction).
extern "C" int b200icp_glibc_rand(unsigned seed, size_t skip, size_t count, int* out) {
  if (!out && count) return fail(B200ICP_EINVAL, "glibc_rand: out is NULL");
  if (seed == 0) seed = 1;
  std::vector<uint32_t> r(34 + 310 + skip + count);
  r[0] = seed;
  for (int i = 1; i < 31; ++i) {
    long long w = (16807LL * (long long)(int32_t)r[i - 1]) % 2147483647LL;
    if (w < 0) w += 2147483647LL;
    r[i] = (uint32_t)w;
  }
  for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
  for (size_t i = 34; i < r.size(); ++i) r[i] = r[i - 31] + r[i - 3];
  for (size_t k = 0; k < count; ++k) out[k] = (int)(r[344 + skip + k] >> 1);
  return B200ICP_OK;
}

int b200icp_reduce_octree(b200icp_ctx* ctx, const double* xyz, const double* normals, size_t n, double voxel_size,
                          int nrpts, unsigned rand_seed, size_t rand_skip, double* xyz_out, double* nrm_out,
                          size_t* n_out) {
  if (!ctx || !xyz || !xyz_out || !n_out) return fail(B200ICP_EINVAL, "reduce_octree: NULL argument");
  if (nrpts == 0) {
    if (normals || nrm_out)   // GetOctTreeCenter copies POINTDIM values out of a 3-value centre (Boctree.h:938-941)
      return fail(B200ICP_EINVAL, "reduce_octree: centre extraction does not carry normals (the reference reads past its centre array there)");
    return b200icp_reduce_octree_center(ctx, xyz, n, voxel_size, xyz_out, n_out);
  }
  if (nrpts != -1 && nrpts != 1)
    return fail(B200ICP_EINVAL, "reduce_octree: nrpts must be 0 (centre), -1 (average) or 1 (one random point per voxel)");
  if (!(voxel_size > 0.0)) return fail(B200ICP_EINVAL, "reduce_octree: voxel_size must be > 0");
  if ((normals == nullptr) != (nrm_out == nullptr)) return fail(B200ICP_EINVAL, "reduce_octree: normals and nrm_out go together");
  *n_out = 0;
  if (n == 0) return B200ICP_OK;
  if (n >= (1ull << 31)) return fail(B200ICP_EINVAL, "reduce_octree: more than 2^31-1 points");
  CU_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevBuf<double> d_xyz, d_nrm, bb_part, bb_out, d_out, d_out_n;
  CU_TRY(d_xyz.alloc_async(3 * n, st));
  CU_TRY(cudaMemcpyAsync(d_xyz.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (normals) {
    CU_TRY(d_nrm.alloc_async(3 * n, st));
    CU_TRY(cudaMemcpyAsync(d_nrm.p, normals, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  const int bb_blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 4);
  CU_TRY(bb_part.alloc_async((size_t)bb_blocks * 6, st));
  CU_TRY(bb_out.alloc_async(6, st));
  bbox_partial_kernel<<<bb_blocks, 256, 0, st>>>(d_xyz.p, n, bb_part.p);
  bbox_final_kernel<<<1, 32, 0, st>>>(bb_part.p, bb_blocks, bb_out.p);
  double bb[6];
  CU_TRY(cudaMemcpyAsync(bb, bb_out.p, sizeof bb, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  for (int k = 0; k < 6; ++k)
    if (!std::isfinite(bb[k])) return fail(B200ICP_EINVAL, "reduce_octree: non-finite coordinate");
  OctRoot root;
  for (int k = 0; k < 3; ++k) root.c[k] = 0.5 * (bb[k] + bb[3 + k]);             // Boctree.h:249-251
  root.size = std::max(std::max(0.5 * (bb[3] - bb[0]), 0.5 * (bb[4] - bb[1])), 0.5 * (bb[5] - bb[2]));
  root.size += 1.0;                                                              // Boctree.h:255
  root.levels = 1;
  for (double hs = root.size / 2.0; hs > voxel_size; hs /= 2.0) ++root.levels;
  if (root.levels > 21) return fail(B200ICP_EINVAL, "reduce_octree: voxel too small for a 63-bit octree key");
  DevBuf<unsigned long long> k_in, k_sorted;
  DevBuf<uint32_t> r_in, r_sorted, heads;
  DevBuf<unsigned char> flags, tmp;
  DevBuf<int> d_count, d_rnd;
  CU_TRY(k_in.alloc_async(n, st)); CU_TRY(k_sorted.alloc_async(n, st));
  CU_TRY(r_in.alloc_async(n, st)); CU_TRY(r_sorted.alloc_async(n, st));
  CU_TRY(heads.alloc_async(n, st)); CU_TRY(flags.alloc_async(n, st)); CU_TRY(d_count.alloc_async(1, st));
  oct_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyz.p, n, root, k_in.p);
  iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(r_in.p, n);
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, k_in.p, k_sorted.p, r_in.p, r_sorted.p, (int)n, 0, 3 * root.levels, st);
  cub::CountingInputIterator<uint32_t> counting(0);
  cub::DeviceSelect::Flagged(nullptr, b2, counting, flags.p, heads.p, d_count.p, (int)n, st);
  CU_TRY(tmp.alloc_async(std::max(b1, b2) + 256, st));
  // (radix sort is stable: the rows of one voxel stay in input order, the order the octree keeps them in its leaf)
  cub::DeviceRadixSort::SortPairs(tmp.p, b1, k_in.p, k_sorted.p, r_in.p, r_sorted.p, (int)n, 0, 3 * root.levels, st);
  oct_heads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(k_sorted.p, n, flags.p);
  cub::DeviceSelect::Flagged(tmp.p, b2, counting, flags.p, heads.p, d_count.p, (int)n, st);
  int m = 0;
  CU_TRY(cudaMemcpyAsync(&m, d_count.p, sizeof m, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (m <= 0 || (size_t)m > n) return fail(B200ICP_ECUDA, "reduce_octree: bad voxel count");
  CU_TRY(d_out.alloc_async(3 * (size_t)m, st));
  if (normals) CU_TRY(d_out_n.alloc_async(3 * (size_t)m, st));
  std::vector<int> h_rnd;
  if (nrpts == 1) {
    h_rnd.resize((size_t)m);
    int rc = b200icp_glibc_rand(rand_seed, rand_skip, (size_t)m, h_rnd.data());
    if (rc != B200ICP_OK) return rc;
    CU_TRY(d_rnd.alloc_async((size_t)m, st));
    CU_TRY(cudaMemcpyAsync(d_rnd.p, h_rnd.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, st));
    oct_extract_kernel<1><<<(unsigned)((m + 127) / 128), 128, 0, st>>>(d_xyz.p, d_nrm.p, r_sorted.p, heads.p, (size_t)m, n,
                                                                      d_rnd.p, d_out.p, d_out_n.p);
  } else {
    oct_extract_kernel<-1><<<(unsigned)((m + 127) / 128), 128, 0, st>>>(d_xyz.p, d_nrm.p, r_sorted.p, heads.p, (size_t)m, n,
                                                                       nullptr, d_out.p, d_out_n.p);
  }
  CU_TRY(cudaMemcpyAsync(xyz_out, d_out.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (normals) CU_TRY(cudaMemcpyAsync(nrm_out, d_out_n.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaGetLastError());
  *n_out = (size_t)m;
  return B200ICP_OK;
}

// ------------------------------------------------------------------------------------ communicator
int b200icp_comm_create(b200icp_ctx* ctx, int rank, int world, void* ipc_handle_out) {
  if (!ctx) return fail(B200ICP_EINVAL, "ctx is NULL");
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
    return fail(B200ICP_EINVAL, "comm_create: need 0 <= rank < world <= 8");
  CU_TRY(cudaSetDevice(ctx->device));
  b200icp_comm_destroy(ctx);
  CU_TRY(cudaMalloc((void**)&ctx->mailbox, sizeof(Mailbox)));   // cudaMalloc (not the pool): IPC-exportable
  CU_TRY(cudaMemset(ctx->mailbox, 0, sizeof(Mailbox)));
  CU_TRY(cudaDeviceSynchronize());
  ctx->comm.rank = rank;
  ctx->comm.world = 1;            // becomes `world` once the peers are connected
  ctx->comm.peer[rank] = ctx->mailbox;
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == B200ICP_COMM_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, ctx->mailbox));
    memcpy(ipc_handle_out, &h, sizeof h);
  }
  (void)world;
  return B200ICP_OK;
}

void* b200icp_comm_mailbox(b200icp_ctx* ctx) { return ctx ? (void*)ctx->mailbox : nullptr; }

int b200icp_comm_connect_ipc(b200icp_ctx* ctx, int world, const void* all_handles) {
  if (!ctx || !ctx->mailbox || !all_handles) return fail(B200ICP_EINVAL, "comm_connect_ipc: create the communicator first");
  if (world < 1 || world > kMaxRanks || ctx->comm.rank >= world) return fail(B200ICP_EINVAL, "comm_connect_ipc: bad world");
  CU_TRY(cudaSetDevice(ctx->device));
  for (int r = 0; r < world; ++r) {
    if (r == ctx->comm.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)all_handles + (size_t)r * sizeof h, sizeof h);
    void* p = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->comm.peer[r] = (Mailbox*)p;
  }
  ctx->comm_ipc = true;
  ctx->comm.world = world;
  return B200ICP_OK;
}

int b200icp_comm_connect_local(b200icp_ctx* ctx, int world, b200icp_ctx* const* all_ctx) {
  if (!ctx || !ctx->mailbox || !all_ctx) return fail(B200ICP_EINVAL, "comm_connect_local: create the communicator first");
  if (world < 1 || world > kMaxRanks || ctx->comm.rank >= world) return fail(B200ICP_EINVAL, "comm_connect_local: bad world");
  CU_TRY(cudaSetDevice(ctx->device));
  for (int r = 0; r < world; ++r) {
    if (r == ctx->comm.rank) continue;
    if (!all_ctx[r] || !all_ctx[r]->mailbox) return fail(B200ICP_EINVAL, "comm_connect_local: peer has no mailbox");
    if (all_ctx[r]->device != ctx->device) {
      int can = 0;
      CU_TRY(cudaDeviceCanAccessPeer(&can, ctx->device, all_ctx[r]->device));
      if (!can) return fail(B200ICP_ENODEV, "comm_connect_local: no peer access between the devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(all_ctx[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU_TRY(e);
      cudaGetLastError();
    }
    ctx->comm.peer[r] = all_ctx[r]->mailbox;
  }
  ctx->comm_ipc = false;
  ctx->comm.world = world;
  return B200ICP_OK;
}

int b200icp_comm_destroy(b200icp_ctx* ctx) {
  if (!ctx) return B200ICP_OK;
  cudaSetDevice(ctx->device);
  if (ctx->comm_ipc)
    for (int r = 0; r < ctx->comm.world; ++r)
      if (r != ctx->comm.rank && ctx->comm.peer[r]) cudaIpcCloseMemHandle(ctx->comm.peer[r]);
  if (ctx->mailbox) cudaFree(ctx->mailbox);
  ctx->mailbox = nullptr;
  ctx->comm = {0, 1, 0, {nullptr}};
  ctx->comm_ipc = false;
  return B200ICP_OK;
}

}  // extern "C"
