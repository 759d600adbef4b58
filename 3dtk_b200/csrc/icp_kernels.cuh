// icp_kernels.cuh -- the per-iteration kernels of the fused match and the API-compatible batch search.
//
//   icp_iter_kernel   one launch per ICP iteration: for every data point  t = X d0,  s = Sinv t,
//                     exact NN of s in the model grid, rejection d^2 < maxdist2, pair formation
//                     (incl. CLOSEST_PLANE_SIMPLE projection) and accumulation of the pair moments.
//                     Replaces Scan::getPtPairs + SearchTree::getPtPairs + the pair walk of Align
//                     (reference src/slam6d/scan.cc:1220-1260, searchTree.cc:92-188,
//                     icp6Dquat.cc:57-71) and makes Scan::transformReduced (scan.cc:851-875)
//                     disappear: the cumulative transform is applied on load.
//                     The scan is handed out in 32-point groups, statically or -- while nearly all points
//                     search -- dynamically from one counter; the point-to-point variants sum their
//                     moments as 64-bit integers (order-independent: reruns stay bit-equal under any
//                     hand-out, and the block sums are added into one row of totals with atomics).
//   solve_step        run by the LAST block of that launch (no second launch): totals of the moments, the
//                     6-DoF solve (solve.h; for QUAT spread over the 32 lanes of a warp,
//                     solve_quat_warp), pose composition (scan.cc:878-898) and the convergence test of
//                     icp6D::match (icp6D.cc:266-279), all in fp64 on device.  The next launch is
//                     released (programmatic dependent launch) when this tail starts.
//   nn_batch_kernel   SearchTree::getPtPairs for caller-supplied queries (API path).
//   lum_link_kernel   lum6DEuler::covarianceEuler / lum6DQuat::covarianceQuat sums of one link.
#pragma once
#include <type_traits>
#include "nn_search.cuh"

namespace b200 {

// ---- optional in-kernel timeline (build with -DB200_TIMING; read with b200icp_debug_timing) ----------------
#ifdef B200_TIMING
__device__ unsigned long long g_tl[32];
__device__ __forceinline__ void tl_mark(int k) {
  if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_tl[k] = t; }
}
// per-block record of the fused iteration kernel: [block][0..3] = start, tile loop done, partials stored, SM id
__device__ unsigned long long g_blk[2048][4];
__device__ __forceinline__ void blk_mark(int k) {
  if (threadIdx.x == 0 && blockIdx.x < 2048) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_blk[blockIdx.x][k] = t;
    if (k == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); g_blk[blockIdx.x][3] = s; }
  }
}
// per-warp record of the fused iteration kernel: [block][warp][0..5] = walk start, walk end (before the block
// barrier), leftover batches start, leftover batches end, search batches run by the warp, (unused)
__device__ unsigned long long g_wrp[2048][8][6];
__device__ __forceinline__ void wrp_mark(int k) {
  if ((threadIdx.x & 31) == 0 && blockIdx.x < 2048) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_wrp[blockIdx.x][threadIdx.x >> 5][k] = t;
  }
}
__device__ __forceinline__ void wrp_count(int k, bool reset) {
  if ((threadIdx.x & 31) == 0 && blockIdx.x < 2048) {
    if (reset) g_wrp[blockIdx.x][threadIdx.x >> 5][k] = 0; else g_wrp[blockIdx.x][threadIdx.x >> 5][k] += 1;
  }
}
#else
__device__ __forceinline__ void tl_mark(int) {}
__device__ __forceinline__ void blk_mark(int) {}
__device__ __forceinline__ void wrp_mark(int) {}
__device__ __forceinline__ void wrp_count(int, bool) {}
#endif

struct XfSmem {
  double X[16], Sinv[16], S[16], Nm[9], o[3];
  float dX[12];     // X - Xprev (rotation block + translation), fp32: per-iteration motion of a data point
  float dXabs;      // max |entry| of the rotation block of dX (rounding bound)
  int need_dd;      // the running minimizer reads the MP_DD moments (HELIX, APX); the others skip their 6 sums
  double fs[NS_P2P]; // fixed-point scales (accumulate_p2p_fix)
};

__device__ __forceinline__ void xf_apply(const double* M, double x, double y, double z, double& ox,
                                         double& oy, double& oz) {
  ox = x * M[0] + y * M[4] + z * M[8] + M[12];
  oy = x * M[1] + y * M[5] + z * M[9] + M[13];
  oz = x * M[2] + y * M[6] + z * M[10] + M[14];
}

// transform3 with the reference's rounding (left-to-right, no contraction), for the API path where
// the query array arrives bit-identical to the reference's.
__device__ __forceinline__ void xf_apply_strict(const double* M, double x, double y, double z,
                                                double& ox, double& oy, double& oz) {
  ox = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, M[0]), __dmul_rn(y, M[4])), __dmul_rn(z, M[8])), M[12]);
  oy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, M[1]), __dmul_rn(y, M[5])), __dmul_rn(z, M[9])), M[13]);
  oz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, M[2]), __dmul_rn(y, M[6])), __dmul_rn(z, M[10])), M[14]);
}

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// thread-private accumulator column in shared memory: element k of this thread lives at base[k*kBlock]
struct SmemAcc {
  double* base;
  __device__ __forceinline__ double& operator[](int k) const { return base[k * kBlock]; }
};

// Order-independent sums: the same thread-private columns, holding 64-bit integers.  A pair adds round(v * s[k]) to
// column k; integer addition is associative, so the totals do not depend on which thread summed which pair, in
// which order -- that is what lets the kernel hand the scan out dynamically and still return bit-identical results
// from run to run.  s[k] is a power of two with nd * max|v| < 2^60 (b200icp_match); rounding < 0.5 / s[k] per pair,
// unbiased: at 1M pairs the totals keep 13-14 digits, like the fp64 sums they replace.
// The conversion is one fused multiply-add: v * s + 1.5 * 2^52 is rounded (to nearest) to an integer that sits in the
// low mantissa bits, so the integer is the bit pattern minus the constant's -- no F2I, which costs the walk 35 % when
// done 17 times per pair.  Needs |v * s| < 2^51 per addend; b200icp_match chooses s accordingly.
constexpr double kFixMagic = 6755399441055744.0;                // 1.5 * 2^52
constexpr long long kFixMagicBits = 0x4338000000000000LL;       // its bit pattern
__device__ __forceinline__ void fix_add(long long* col, int k, double v, double s) {
  col[k * kBlock] += __double_as_longlong(__fma_rn(v, s, kFixMagic)) - kFixMagicBits;
}
// accumulate_p2p (solve.h) on integer columns.  fs = the scales: one for the count, one for sum |p1 - p2|^2, one for
// the first moments, one for the second moments (fs[MP_N], fs[MP_D2], fs[MP_M], fs[MP_DM]; MP_D = MP_M, MP_DD = MP_DM).
// The data point is scaled once, so a second moment is one fma like in the fp64 version.
__device__ __forceinline__ void accumulate_p2p_fix(long long* col, const double* __restrict__ fs, const double* p1,
                                                   const double* p2, const double* o, bool dd) {
  const double a[3] = {p1[0] - o[0], p1[1] - o[1], p1[2] - o[2]};
  const double b[3] = {p2[0] - o[0], p2[1] - o[1], p2[2] - o[2]};
  const double e0 = p1[0] - p2[0], e1 = p1[1] - p2[1], e2 = p1[2] - p2[2];
  const double s1 = fs[MP_M], s2 = fs[MP_DM];
  fix_add(col, MP_N, 1.0, fs[MP_N]);
  fix_add(col, MP_D2, e0 * e0 + e1 * e1 + e2 * e2, fs[MP_D2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) { fix_add(col, MP_M + i, a[i], s1); fix_add(col, MP_D + i, b[i], s1); }
  const double bs[3] = {b[0] * s2, b[1] * s2, b[2] * s2};   // exact: s2 is a power of two
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) fix_add(col, MP_DM + 3 * i + j, bs[i], a[j]);
  if (dd) {
    fix_add(col, MP_DD + 0, bs[0], b[0]); fix_add(col, MP_DD + 1, bs[0], b[1]); fix_add(col, MP_DD + 2, bs[0], b[2]);
    fix_add(col, MP_DD + 3, bs[1], b[1]); fix_add(col, MP_DD + 4, bs[1], b[2]); fix_add(col, MP_DD + 5, bs[2], b[2]);
  }
}

template <int NS, class Acc>
__device__ __forceinline__ void block_reduce_store(Acc acc, double* __restrict__ out) {
  __shared__ double red[kWarps][NS_MAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 4
  for (int k = 0; k < NS; ++k) {
    double v = acc[k];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    double v = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) v += red[w][threadIdx.x];
    out[threadIdx.x] = v;
  }
}

template <int NS>
__device__ __forceinline__ void block_reduce_store_regs(double (&acc)[NS], double* __restrict__ out) {
  __shared__ double red[kWarps][NS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    double v = acc[k];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    double v = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) v += red[w][threadIdx.x];
    out[threadIdx.x] = v;
  }
}

// Runs in the LAST block of icp_iter_kernel to finish (all kBlock threads): fixed-order reduction of the
// per-block moments, the 6-DoF solve and the loop bookkeeping of icp6D::match.  Deterministic: block b's
// partial always enters the sum at the same place.
// The serial part of one ICP iteration (one thread; `st` lives in shared memory).
__device__ __forceinline__ void solve_step_serial(IterState* st, const double* mom, double* __restrict__ rms_log,
                                                  unsigned long long* __restrict__ npairs_log,
                                                  unsigned* __restrict__ stage2_log,
                                                  unsigned* __restrict__ stage2_counter) {
  // ---- icp6D::match loop body after getPtPairs (icp6D.cc:124-125, :229-279)
  tl_mark(20);
  const int iter = st->iter;
  st->prev_prev_ret = st->prev_ret;
  st->prev_ret = st->ret;
  st->stage2_last = atomicExch(stage2_counter, 0u);
  const unsigned searches_now = atomicExch(stage2_counter + 1, 0u);
  const double np = mom[0];
  if (!(np > 3.0)) {  // "do we have enough point pairs?" -> break before any transform
    st->done = 1;
    st->ret_iter = iter;
    return;
  }
  tl_mark(21);
  double alignxf[16];
  for (int i = 0; i < 16; ++i) alignxf[i] = st->alignxf[i];  // kept when the Cholesky path bails out
  const double ret = solve_any(st->algo, mom, st->o, st->napx_weighted, alignxf);
  tl_mark(22);
  st->ret = ret;
  for (int i = 0; i < 16; ++i) st->alignxf[i] = alignxf[i];
  rms_log[st->iters_run] = ret;
  npairs_log[st->iters_run] = (unsigned long long)(np + 0.5);
  stage2_log[2 * st->iters_run] = st->stage2_last;
  stage2_log[2 * st->iters_run + 1] = searches_now;
  st->searches_last = searches_now;
  st->iters_run += 1;
  // Scan::transformMatrix (scan.cc:878-898): transMat <- alignxf*transMat, dalignxf <- alignxf*dalignxf
  tl_mark(23);
  double tmp[16];
  m4_mul(alignxf, st->X, tmp);
  for (int i = 0; i < 16; ++i) { st->Xprev[i] = st->X[i]; st->X[i] = tmp[i]; }
  m4_mul(alignxf, st->T, tmp);
  for (int i = 0; i < 16; ++i) st->T[i] = tmp[i];
  if (st->pose_log)
    for (int i = 0; i < 16; ++i) st->pose_log[16 * (st->iters_run - 1) + i] = tmp[i];
  // transform3normal (globals.icc:1465-1475) multiplies by the transposed rotation block:
  // Nm <- R^T Nm, R(r,c) = alignxf[4c+r]
  double nn[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      nn[3 * r + c] = alignxf[4 * r + 0] * st->Nm[c] + alignxf[4 * r + 1] * st->Nm[3 + c] +
                      alignxf[4 * r + 2] * st->Nm[6 + c];
  for (int i = 0; i < 9; ++i) st->Nm[i] = nn[i];
  tl_mark(24);
  if ((fabs(ret - st->prev_ret) < st->eps && fabs(ret - st->prev_prev_ret) < st->eps) ||
      iter == st->max_iter - 1) {
    st->done = 1;
    st->ret_iter = iter;
  } else {
    st->iter = iter + 1;
  }
}

// ---- the same serial part on the 32 lanes of one warp, for icp6D_QUAT (the default minimizer) -------------------
// One thread running solve_quat inside this kernel inherits its 80-register budget: the 4x4 work arrays go to local
// memory and the solve takes 4.6 us against 1.8 us in a kernel of its own (tools/bench_src/solve_bench.cu), on the
// serial tail of EVERY iteration.  Here the small dense pieces are spread over lanes instead: the 3x3 S, Horn's 4x4 Q,
// the characteristic quartic from Q's principal minors (x^4 - e1 x^3 + e2 x^2 - e3 x + e4), the adjugate of
// Q - lambda I (16 cofactors, one per lane), R, t and the two pose compositions; only Newton's iteration for lambda_max
// stays serial (every lane runs it redundantly).  Same formulas as solve_quat / sym4_max_eigvec up to the order of a
// few additions (1e-16 relative); w: >= 80 doubles of shared memory.
__device__ __forceinline__ double det3_of(const double* A, int r0, int r1, int r2, int c0, int c1, int c2) {
  return det3(A[4 * r0 + c0], A[4 * r0 + c1], A[4 * r0 + c2], A[4 * r1 + c0], A[4 * r1 + c1], A[4 * r1 + c2],
              A[4 * r2 + c0], A[4 * r2 + c1], A[4 * r2 + c2]);
}
// cofactor (r, c) of the 4x4 row-major A
__device__ __forceinline__ double cof4(const double* A, int r, int c) {
  const int r0 = r == 0 ? 1 : 0, r1 = r <= 1 ? 2 : 1, r2 = r <= 2 ? 3 : 2;
  const int c0 = c == 0 ? 1 : 0, c1 = c <= 1 ? 2 : 1, c2 = c <= 2 ? 3 : 2;
  const double d = det3_of(A, r0, r1, r2, c0, c1, c2);
  return ((r + c) & 1) ? -d : d;
}

__device__ __forceinline__ double solve_quat_warp(const double* __restrict__ mom, const double* __restrict__ o,
                                                  double* __restrict__ alignxf, double* __restrict__ w) {
  const int lane = threadIdx.x & 31;
  double* S = w;          // [9]  centred cross-covariance / n
  double* Q = w + 16;     // [16] Horn's N, row-major
  double* A = w + 32;     // [16] Q - lambda I
  double* C = w + 48;     // [16] cofactors of A
  double* term = w + 64;  // [16] principal minors of Q
  const double n = mom[MP_N];
  const double inv = 1.0 / n;
  const double cm[3] = {mom[MP_M] * inv, mom[MP_M + 1] * inv, mom[MP_M + 2] * inv};
  const double cd[3] = {mom[MP_D] * inv, mom[MP_D + 1] * inv, mom[MP_D + 2] * inv};
  if (lane < 9) S[lane] = mom[MP_DM + lane] * inv - cd[lane / 3] * cm[lane % 3];
  __syncwarp();
  if (lane < 16) {
    const int i = lane >> 2, j = lane & 3;
    const double trace = S[0] + S[4] + S[8];
    double q;
    if (i == 0 && j == 0) q = trace;
    else if (i == 0 || j == 0) {
      const int k = i + j;   // 1, 2, 3 -> S12 - S21, S20 - S02, S01 - S10
      q = k == 1 ? S[5] - S[7] : (k == 2 ? S[6] - S[2] : S[1] - S[3]);
    } else {
      q = S[3 * (i - 1) + (j - 1)] + S[3 * (j - 1) + (i - 1)] - (i == j ? trace : 0.0);
    }
    Q[lane] = q;
  }
  __syncwarp();
  // coefficients: e2 = sum of the 6 principal 2x2 minors, e3 = sum of the 4 principal 3x3 minors, e4 = det Q
  {
    double t = 0.0;
    if (lane < 6) {
      const int i = lane < 3 ? 0 : (lane < 5 ? 1 : 2), j = lane < 3 ? lane + 1 : (lane < 5 ? lane - 1 : 3);
      t = Q[5 * i] * Q[5 * j] - Q[4 * i + j] * Q[4 * j + i];
    } else if (lane < 10) {
      t = cof4(Q, lane - 6, lane - 6);
    } else if (lane < 14) {
      t = Q[lane - 10] * cof4(Q, 0, lane - 10);
    }
    if (lane < 14) term[lane] = t;
  }
  __syncwarp();
  double fro = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) fro += Q[k] * Q[k];
  const double e1 = (Q[0] + Q[5]) + (Q[10] + Q[15]);
  const double e2 = ((term[0] + term[1]) + (term[2] + term[3])) + (term[4] + term[5]);
  const double e3 = (term[6] + term[7]) + (term[8] + term[9]);
  const double e4 = (term[10] + term[11]) + (term[12] + term[13]);
  const double c1 = -e1, c2 = e2, c3 = -e3, c4 = e4;
  double x = sqrt(fro) * (1.0 + 1e-12) + 1e-300;   // >= spectral radius >= lambda_max (as sym4_max_eigvec)
  for (int it = 0; it < 100; ++it) {
    const double p = (((x + c1) * x + c2) * x + c3) * x + c4;
    const double dp = ((4.0 * x + 3.0 * c1) * x + 2.0 * c2) * x + c3;
    if (!(dp > 0.0)) break;
    const double xn = x - p / dp;
    if (!(xn < x)) break;
    const bool tiny = (x - xn) <= 4e-16 * fabs(x);
    x = xn;
    if (tiny) break;
  }
  if (lane < 16) A[lane] = Q[lane] - ((lane >> 2) == (lane & 3) ? x : 0.0);
  __syncwarp();
  if (lane < 16) C[lane] = cof4(A, lane >> 2, lane & 3);
  __syncwarp();
  int best = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (fabs(C[5 * i]) > fabs(C[5 * best])) best = i;
  double q0 = C[best], q1 = C[4 + best], q2 = C[8 + best], q3 = C[12 + best];   // column `best` of the adjugate
  const double n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
  if (!(n2 > 0.0)) { q0 = 1.0; q1 = q2 = q3 = 0.0; }
  const double ql = 1.0 / sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  q0 *= ql; q1 *= ql; q2 *= ql; q3 *= ql;
  const double q00 = q0 * q0, q11 = q1 * q1, q22 = q2 * q2, q33 = q3 * q3;
  const double q03 = q0 * q3, q13 = q1 * q3, q23 = q2 * q3, q02 = q0 * q2, q12 = q1 * q2, q01 = q0 * q1;
  double R[3][3];
  R[0][0] = q00 + q11 - q22 - q33;
  R[1][1] = q00 - q11 + q22 - q33;
  R[2][2] = q00 - q11 - q22 + q33;
  R[0][1] = 2.0 * (q12 - q03);
  R[1][0] = 2.0 * (q12 + q03);
  R[0][2] = 2.0 * (q13 + q02);
  R[2][0] = 2.0 * (q13 - q02);
  R[1][2] = 2.0 * (q23 - q01);
  R[2][1] = 2.0 * (q23 + q01);
  if (lane == 0) {
    m4_identity(alignxf);
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) alignxf[4 * c + r] = R[r][c];
    const double cmo[3] = {cm[0] + o[0], cm[1] + o[1], cm[2] + o[2]};
    const double cdo[3] = {cd[0] + o[0], cd[1] + o[1], cd[2] + o[2]};
    set_translation_from_centroids(alignxf, cmo, cdo);
  }
  __syncwarp();
  return sqrt(mom[MP_D2] * inv);
}

// solve_step_serial for icp6D_QUAT, run by ALL lanes of warp 0
__device__ __forceinline__ void solve_step_quat_warp(IterState* st, const double* mom, double* __restrict__ rms_log,
                                                     unsigned long long* __restrict__ npairs_log,
                                                     unsigned* __restrict__ stage2_log,
                                                     unsigned* __restrict__ stage2_counter, double* __restrict__ w) {
  const int lane = threadIdx.x & 31;
  tl_mark(20);
  const int iter = st->iter;
  const double prev_ret = st->ret, prev_prev_ret = st->prev_ret;
  unsigned searches_now = 0;
  __syncwarp();
  if (lane == 0) {
    st->prev_prev_ret = prev_prev_ret;
    st->prev_ret = prev_ret;
    st->stage2_last = atomicExch(stage2_counter, 0u);
    searches_now = atomicExch(stage2_counter + 1, 0u);
  }
  const double np = mom[0];
  if (!(np > 3.0)) {  // "do we have enough point pairs?" -> break before any transform
    if (lane == 0) { st->done = 1; st->ret_iter = iter; }
    return;
  }
  tl_mark(21);
  const double ret = solve_quat_warp(mom, st->o, st->alignxf, w);
  tl_mark(22);
  const int run = st->iters_run;
  __syncwarp();
  if (lane == 0) {
    st->ret = ret;
    rms_log[run] = ret;
    npairs_log[run] = (unsigned long long)(np + 0.5);
    stage2_log[2 * run] = st->stage2_last;
    stage2_log[2 * run + 1] = searches_now;
    st->searches_last = searches_now;
    st->iters_run = run + 1;
  }
  tl_mark(23);
  // Scan::transformMatrix (scan.cc:878-898): lanes 0-15 X <- alignxf*X, lanes 16-31 T <- alignxf*T (one entry each,
  // MMult's operand order)
  {
    const double* A = st->alignxf;
    const double* B = lane < 16 ? st->X : st->T;
    const int e = lane & 15, r = e & 3, c = e >> 2;
    const double v = A[r] * B[4 * c] + A[r + 4] * B[4 * c + 1] + A[r + 8] * B[4 * c + 2] + A[r + 12] * B[4 * c + 3];
    // transform3normal (globals.icc:1465-1475): Nm <- R^T Nm
    double nv = 0.0;
    if (lane < 9) {
      const int nr = lane / 3, nc = lane % 3;
      nv = A[4 * nr + 0] * st->Nm[nc] + A[4 * nr + 1] * st->Nm[3 + nc] + A[4 * nr + 2] * st->Nm[6 + nc];
    }
    __syncwarp();
    if (lane < 16) { st->Xprev[e] = st->X[e]; st->X[e] = v; }
    else {
      st->T[e] = v;
      if (st->pose_log) st->pose_log[16 * run + e] = v;
    }
    if (lane < 9) st->Nm[lane] = nv;
  }
  tl_mark(24);
  if (lane == 0) {
    if ((fabs(ret - prev_ret) < st->eps && fabs(ret - prev_prev_ret) < st->eps) || iter == st->max_iter - 1) {
      st->done = 1;
      st->ret_iter = iter;
    } else {
      st->iter = iter + 1;
    }
  }
  __syncwarp();
}

// fixed: the partial rows hold 64-bit integers (accumulate_p2p_fix): summed exactly, then scaled back by fix_inv
__device__ __noinline__ void solve_step(IterState* __restrict__ gst, const double* __restrict__ partials,
                                        int nblocks, bool fixed, double* __restrict__ rms_log,
                                        unsigned long long* __restrict__ npairs_log,
                                        unsigned* __restrict__ stage2_log,
                                        unsigned* __restrict__ stage2_counter, double* scratch,
                                        const CommDev& comm) {
  // scratch: >= (kWarps+1)*NS_MAX + sizeof(IterState)/8 doubles of shared memory (the accumulator columns,
  // free by now) -- no static shared memory of its own, the kernel sits at the 3-blocks/SM limit
  static_assert(sizeof(IterState) % 8 == 0, "IterState is copied as 8-byte words");
  constexpr int kWords = (int)(sizeof(IterState) / 8);
  double (*wpart)[NS_MAX] = reinterpret_cast<double (*)[NS_MAX]>(scratch);
  double* mom = scratch + kWarps * NS_MAX;
  double* wq = mom + NS_MAX;          // 80 doubles of work space for the warp-parallel QUAT solve
  double* st_raw = wq + 80;
  IterState* st = reinterpret_cast<IterState*>(st_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // stage the loop state in shared memory: one coalesced read now, one coalesced write at the end, instead
  // of a single thread chasing ~100 dependent global accesses
  tl_mark(8);
  for (int i = tid; i < kWords; i += kBlock)
    st_raw[i] = __ldcg(reinterpret_cast<const double*>(gst) + i);
  const int NS = moment_count(gst->algo);
  __syncthreads();   // scratch aliases the accumulators other warps may still be reading
  tl_mark(9);
  // reduction: lane = moment, warp w takes blocks w, w+8, ... -> every load is one coalesced 184/352-byte row
  // of a block's partials; fixed shape, so the sums do not depend on timing
  // (groups of 16 predicated loads, all in flight together, no remainder loop: a rolled remainder exposes one L2
  //  round trip per block on the serial tail of the iteration -- 5.4 -> 3.5 us measured)
  constexpr int kRedGroup = 16;
  const int trips = (nblocks - warp + kWarps - 1) / kWarps;   // blocks this warp sums: warp, warp + 8, ...
  if (!fixed)
  for (int k0 = 0; k0 < NS; k0 += 32) {
    const int k = k0 + lane;
    double v = 0.0;
    if (k < NS) {
      for (int j0 = 0; j0 < trips; j0 += kRedGroup) {
        double x[kRedGroup];
#pragma unroll
        for (int u = 0; u < kRedGroup; ++u) {
          const int b = warp + (j0 + u) * kWarps;
          x[u] = b < nblocks ? __ldcg(partials + (size_t)b * NS_MAX + k) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kRedGroup; ++u) v += x[u];
      }
      wpart[warp][k] = v;
    }
  }
  __syncthreads();
  if (tid < NS) {
    if (fixed) {
      // the blocks have added their integer sums into row 0 (icp_iter_kernel); read, scale back, clear for the next launch
      unsigned long long* tot = reinterpret_cast<unsigned long long*>(const_cast<double*>(partials)) + tid;
      const long long iv = (long long)__ldcg(tot);
      *tot = 0ull;
      mom[tid] = (double)iv * st->fix_inv[tid];
    } else {
      mom[tid] = ((wpart[0][tid] + wpart[1][tid]) + (wpart[2][tid] + wpart[3][tid])) +
                 ((wpart[4][tid] + wpart[5][tid]) + (wpart[6][tid] + wpart[7][tid]));
    }
  }
  __syncthreads();
  tl_mark(10);
  if (comm.world > 1) {
    // ---- fused all-reduce over NVLink peer memory (query-sharded match): this rank's moments are stored
    // straight into every rank's mailbox, a flag with the iteration's sequence number follows, and every
    // rank sums the world's rows in rank order -- bit-identical totals everywhere, so the replicated solve
    // below keeps all ranks' loop state identical without a broadcast.
    const unsigned long long seq = comm.seq_base + (unsigned long long)st->iter;
    const int slot = (int)(seq & 1ull);
    for (int p = warp; p < comm.world; p += kWarps)
      for (int k = lane; k < NS; k += 32) comm.peer[p]->box[slot][comm.rank][k] = mom[k];
    __threadfence_system();
    __syncthreads();
    if (tid < comm.world) {
      volatile unsigned long long* f = &comm.peer[tid]->flag[slot][comm.rank];
      *f = seq;
      __threadfence_system();
    }
    __shared__ int comm_ok;
    if (tid == 0) comm_ok = 1;
    __syncthreads();
    if (tid < comm.world) {
      volatile unsigned long long* f = &comm.peer[comm.rank]->flag[slot][tid];
      const long long t0 = clock64();
      while (*f != seq)
        if (clock64() - t0 > 4000000000LL) { comm_ok = 0; break; }   // ~2 s: a peer died; do not hang the GPU
      __threadfence_system();
    }
    __syncthreads();
    if (!comm_ok) {
      if (tid == 0) { st->done = 1; st->ret_iter = -1000 - st->iter; }
    } else if (tid < NS) {
      const volatile double* row = &comm.peer[comm.rank]->box[slot][0][tid];
      double t = 0.0;
      for (int p = 0; p < comm.world; ++p) t += row[(size_t)p * NS_MAX];
      mom[tid] = t;
    }
    __syncthreads();
    if (comm_ok) {
      if (st->algo == 1) { if (warp == 0) solve_step_quat_warp(st, mom, rms_log, npairs_log, stage2_log, stage2_counter, wq); }
      else if (tid == 0) solve_step_serial(st, mom, rms_log, npairs_log, stage2_log, stage2_counter);
    }
  } else if (st->algo == 1) {
    if (warp == 0) solve_step_quat_warp(st, mom, rms_log, npairs_log, stage2_log, stage2_counter, wq);
  } else if (tid == 0) {
    solve_step_serial(st, mom, rms_log, npairs_log, stage2_log, stage2_counter);
  }
  __syncthreads();
  tl_mark(11);
  for (int i = tid; i < kWords; i += kBlock)
    reinterpret_cast<double*>(gst)[i] = st_raw[i];
  tl_mark(12);
}


constexpr int kQueueCap = 64;  // per-warp queue of queries waiting for a full search
#ifndef B200_TILE_SPAN
#define B200_TILE_SPAN 96
#endif
constexpr int kTileSpan = B200_TILE_SPAN;   // a batch whose queries span at most this many rows of the data scan is "dense"

// dynamic shared memory of icp_iter_kernel
template <int NS>
struct IterSmem {
  SearchSmem search;
  int queue[kWarps][kQueueCap];
  int leftover[kWarps * 32];      // end-of-kernel merge of the warps' partial queues
  int left_count[kWarps];
  double acc[NS][kBlock];
};

struct PairCtx {   // everything needed to turn (query, neighbour) into moments
  const GridDev* model;
  const double4* dn;
  const XfSmem* xf;
};

// forms the pair of data point i (current position t, neighbour at sorted position bj) and accumulates it
// (mx,my,mz) = the neighbour's coordinates in the model grid's frame (p64 of the model scan)
template <bool NAPX, bool PLANE, bool FIXP = false, class Acc>
__device__ __forceinline__ void accumulate_pair_pt(const PairCtx& pc, Acc&& acc, uint32_t i, double mx, double my,
                                                   double mz, double tx, double ty, double tz) {
  const XfSmem& xf = *pc.xf;
  double p1[3], p2[3] = {tx, ty, tz};
  xf_apply(xf.S, mx, my, mz, p1[0], p1[1], p1[2]);
  double nv[3] = {0, 0, 0};
  if (PLANE) {
    const double2 na = __ldg(reinterpret_cast<const double2*>(pc.dn + i));
    const double n2 = __ldg(reinterpret_cast<const double*>(pc.dn + i) + 2);
    nv[0] = xf.Nm[0] * na.x + xf.Nm[1] * na.y + xf.Nm[2] * n2;
    nv[1] = xf.Nm[3] * na.x + xf.Nm[4] * na.y + xf.Nm[5] * n2;
    nv[2] = xf.Nm[6] * na.x + xf.Nm[7] * na.y + xf.Nm[8] * n2;
    const double nl = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);  // Normalize3
    nv[0] /= nl; nv[1] /= nl; nv[2] /= nl;
    // s <- (n.(s-t)) n + t   (searchTree.cc:149-162)
    const double dot = nv[0] * (p1[0] - p2[0]) + nv[1] * (p1[1] - p2[1]) + nv[2] * (p1[2] - p2[2]);
    p1[0] = nv[0] * dot + p2[0];
    p1[1] = nv[1] * dot + p2[1];
    p1[2] = nv[2] * dot + p2[2];
  }
  if constexpr (NAPX) accumulate_napx(acc, p1, p2, nv, xf.o);
  else if constexpr (FIXP) accumulate_p2p_fix(reinterpret_cast<long long*>(acc.base), xf.fs, p1, p2, xf.o, xf.need_dd != 0);
  else accumulate_p2p(acc, p1, p2, xf.o, xf.need_dd != 0);
}

template <bool NAPX, bool PLANE, bool FIXP = false, class Acc>
__device__ __forceinline__ void accumulate_pair(const PairCtx& pc, Acc&& acc, uint32_t i, int bj, double tx,
                                                double ty, double tz) {
  const double2 pa = __ldg(reinterpret_cast<const double2*>(pc.model->p64 + bj));
  const double pz = __ldg(reinterpret_cast<const double*>(pc.model->p64 + bj) + 2);
  accumulate_pair_pt<NAPX, PLANE, FIXP>(pc, acc, i, pa.x, pa.y, pz, tx, ty, tz);
}

// current position of data point i, its position one iteration ago and the margin worth asking for
struct Pt3 { double x, y; };
__device__ __forceinline__ void query_state_pt(const XfSmem& xf, const Pt3 a, const double z0,
                                               bool can_skip, float dmin, float dmax, double& tx, double& ty,
                                               double& tz, double& sx, double& sy, double& sz, float& step,
                                               float& delta) {
  xf_apply(xf.X, a.x, a.y, z0, tx, ty, tz);
  xf_apply(xf.Sinv, tx, ty, tz, sx, sy, sz);
  step = 0.f;
  delta = 0.f;
  if (can_skip) {
    // |s - s_prev| = |t - t_prev| (S is rigid) = |(X - Xprev) d0|, evaluated in fp32 from the fp64 difference
    // matrix and inflated by its rounding bound: only an UPPER bound of the motion is needed
    const float ax = (float)a.x, ay = (float)a.y, az = (float)z0;
    const float vx = fmaf(xf.dX[0], ax, fmaf(xf.dX[3], ay, fmaf(xf.dX[6], az, xf.dX[9])));
    const float vy = fmaf(xf.dX[1], ax, fmaf(xf.dX[4], ay, fmaf(xf.dX[7], az, xf.dX[10])));
    const float vz = fmaf(xf.dX[2], ax, fmaf(xf.dX[5], ay, fmaf(xf.dX[8], az, xf.dX[11])));
    const float bound = 4e-7f * (xf.dXabs * (fabsf(ax) + fabsf(ay) + fabsf(az)) + fabsf(xf.dX[9]) +
                                 fabsf(xf.dX[10]) + fabsf(xf.dX[11]));
    step = sqrtf(fmaf(vx, vx, fmaf(vy, vy, vz * vz))) * 1.00001f + 1.8f * bound + 1e-12f;
    // the scanned cells certify a margin for free (nn_search.cuh); an extra one is worth scanning for only
    // when it is small next to the typical runner-up gap and covers the motion still to come
    delta = 8.0f * step <= dmax ? fmaxf(8.0f * step, dmin) : 0.f;
  }
}

__device__ __forceinline__ void query_state(const XfSmem& xf, const double4* __restrict__ dq, uint32_t i,
                                            bool can_skip, float dmin, float dmax, double& tx, double& ty,
                                            double& tz, double& sx, double& sy, double& sz, float& step,
                                            float& delta) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(dq + i));
  const double z0 = __ldg(reinterpret_cast<const double*>(dq + i) + 2);
  query_state_pt(xf, Pt3{a.x, a.y}, z0, can_skip, dmin, dmax, tx, ty, tz, sx, sy, sz, step, delta);
}

// stages the transforms of the running iteration in shared memory (all threads call; caller syncs)
__device__ __forceinline__ void load_xf(XfSmem& xf, const IterState* __restrict__ st) {
  const int tid = threadIdx.x;
  if (tid < 16) { xf.X[tid] = st->X[tid]; xf.Sinv[tid] = st->Sinv[tid]; xf.S[tid] = st->S[tid]; }
  if (tid < 9) xf.Nm[tid] = st->Nm[tid];
  if (tid < 3) xf.o[tid] = st->o[tid];
  if (tid >= 32 && tid < 44) {   // dX[3c+r] = (X - Xprev)(r, c): columns 0..2 rotation block, column 3 translation
    const int k = tid - 32, c = k / 3, r = k % 3;
    xf.dX[k] = (float)(st->X[4 * c + r] - st->Xprev[4 * c + r]);
  }
  if (tid == 65) xf.need_dd = algo_needs_dd(st->algo) ? 1 : 0;
  if (tid >= 128 && tid < 128 + NS_P2P) xf.fs[tid - 128] = st->fix_s[tid - 128];
  if (tid == 64) {
    double m = 0.0;
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) m = fmax(m, fabs(st->X[4 * c + r] - st->Xprev[4 * c + r]));
    xf.dXabs = (float)m * 1.000001f + 1e-30f;
  }
}

#ifndef B200_LEFT_SCAN_BATCH
#define B200_LEFT_SCAN_BATCH 4   // candidate loads in flight per lane in the leftover batches (2: 7.60, 4: 7.56 ms)
#endif
#ifndef B200_PDL
#define B200_PDL 1
#endif
#ifndef B200_FIXPOINT
#define B200_FIXPOINT 1   // 0: fp64 sums, static hand-out only (A/B builds)
#endif
#ifndef B200_DUAL_TILE
#define B200_DUAL_TILE 0
#endif
#ifndef B200_TRIP_CHUNK
#define B200_TRIP_CHUNK 1   // > 1: see the tile loop of icp_iter_kernel (experiment, default off)
#endif
#ifndef B200_ITER_MINBLOCKS
#define B200_ITER_MINBLOCKS 3
#endif
template <bool NAPX, bool PLANE, bool EXACT>
__global__ void __launch_bounds__(kBlock, B200_ITER_MINBLOCKS)
icp_iter_kernel(GridDev model, const double4* __restrict__ dq, const double4* __restrict__ dn,
                uint32_t nd, IterState* __restrict__ st, double maxdist2, int rnd,
                int* __restrict__ nn_cache, float* __restrict__ nn_budget, double* __restrict__ partials,
                unsigned* __restrict__ stage2_counter, double* __restrict__ rms_log,
                unsigned long long* __restrict__ npairs_log, unsigned* __restrict__ stage2_log,
                const __grid_constant__ CommDev comm) {
#if B200_PDL
  // programmatic dependent launch: this launch may have been started while its predecessor's last block was still
  // reducing and solving (see the trigger below); nothing of the predecessor may be read before this returns
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  if (st->done) return;
  constexpr int NS = NAPX ? (int)NS_NAPX : (int)NS_P2P;
  constexpr bool FIXP = B200_FIXPOINT && !NAPX && !PLANE;   // order-independent sums + dynamic hand-out
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  IterSmem<NS>& sm = *reinterpret_cast<IterSmem<NS>*>(dyn_smem);
  __shared__ XfSmem xf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  blk_mark(0);
  load_xf(xf, st);
  const unsigned iter_salt = (unsigned)st->iter * 0x9E3779B9u;
  // a query may skip its search while it has moved less than its certified budget (nn_search.cuh);
  // needs the previous pose (iter > 0) and every point visited every iteration (no subsampling)
  const bool can_skip = st->iters_run > 0 && rnd <= 1;
  const float dmax = 0.1f * (float)model.h, dmin = 1e-3f * (float)model.h;
  // Hand-out mode of this launch (every block takes the same decision from state the previous launch left).  While
  // (nearly) all points search, a warp's share of the scan is ~9 batches of 32 searches, each as long as its slowest
  // lane (10-60 us): with a static split the slowest warp takes 1.6x the median (tools/warp_times.py).  Then the
  // warps draw their 32-point groups from one counter instead; the order-independent sums keep the result
  // reproducible.  (Runs of 2 / 4 groups per draw: 8.2 / 9.8 ms per match against 7.4 -- one group it is.)
  const bool dyn = FIXP && rnd <= 1 && (st->iters_run == 0 || st->searches_last > nd - (nd >> 4));   // > 15/16
  SmemAcc acc{&sm.acc[0][tid]};
#pragma unroll 4
  for (int k = 0; k < NS; ++k) acc[k] = 0.0;
  __syncthreads();

  PairCtx pc{&model, dn, &xf};
  unsigned stage2 = 0, searches = 0;
  int qcount = 0;            // warp-uniform number of queued queries
  int* queue = sm.queue[warp];

  // full search of up to 32 queued queries (list[0..nb)), one per lane
  // tag: std::integral_constant<int, 0> = batch inside the walk, <int, 1> = leftover batch after it (four candidate
  // loads in flight per lane instead of two: latency matters there, not registers or issue slots)
  auto run_batch = [&](const int* list, int nb, auto tag) {
    constexpr bool kLeft = decltype(tag)::value != 0;
    wrp_count(4, false);
    const bool on = lane < nb;
    uint32_t i = 0;
    double tx = 0, ty = 0, tz = 0, sx = 0, sy = 0, sz = 0;
    float step = 0.f, delta = 0.f;
    int seed = -1;
    if (on) {
      i = (uint32_t)list[lane];
      seed = nn_cache[i];
      query_state(xf, dq, i, can_skip, dmin, dmax, tx, ty, tz, sx, sy, sz, step, delta);
    }
    int bj;
    double bd2;
    unsigned boidx;
    float newbud;
    // neighbours in the data scan's cell order -> overlapping stencils -> cooperative tile search
    const bool dense = nb == 32 && (unsigned)(list[31] - list[0]) <= (unsigned)kTileSpan;
    nn_warp_search<EXACT, kLeft ? B200_LEFT_SCAN_BATCH : kScanBatch>(
        model, sm.search, on, sx, sy, sz, maxdist2, seed, delta, bj, bd2, boidx, newbud, stage2, dense);
    if (on) {
      nn_cache[i] = bj;
      nn_budget[i] = newbud;
      if (bj >= 0) accumulate_pair<NAPX, PLANE, FIXP>(pc, acc, i, bj, tx, ty, tz);
    }
    __syncwarp();
  };

  const uint32_t ntiles = (nd + kBlock - 1) / kBlock;
  wrp_count(4, true);
  wrp_mark(0);
  // streaming part of one query: skip path inline, otherwise report "needs a search"
  auto stream_one = [&](uint32_t i, bool active) -> bool {
    if (rnd > 1 && active) active = (hash32(i ^ iter_salt) % (unsigned)rnd) == 0u;
    bool search = false;
    if (active) {
      double tx, ty, tz, sx, sy, sz;
      float step, delta;
      // (the cached neighbour's index is fetched together with the point and its budget, not after the budget test:
      //  one dependent round trip less per trip of the walk)
      const int seed = can_skip ? nn_cache[i] : -1;
      const float bud0 = can_skip ? nn_budget[i] : 0.f;
      query_state(xf, dq, i, can_skip, dmin, dmax, tx, ty, tz, sx, sy, sz, step, delta);
      const float bud = bud0 - step;
      search = !(bud > 0.f);
      if (!search) {
        // the cached neighbour is still THE nearest neighbour; only its distance is re-evaluated
        nn_budget[i] = bud;
        if (seed >= 0) {
          const double d2 = EXACT ? exact_d2(model, seed, sx, sy, sz)
                                  : (double)dist32(__ldg(model.p32 + seed), (float)(sx - model.c[0]),
                                                   (float)(sy - model.c[1]), (float)(sz - model.c[2]));
          if (d2 < maxdist2) accumulate_pair<NAPX, PLANE, FIXP>(pc, acc, i, seed, tx, ty, tz);
        }
      }
    }
    return search;
  };
  // queue the queries that need a full search; run them 32 at a time so every lane works
  auto enqueue = [&](uint32_t i, bool search) {
    const unsigned smask = __ballot_sync(0xffffffffu, search);
    if (smask) {
      if (search) queue[qcount + __popc(smask & ((1u << lane) - 1u))] = (int)i;
      qcount += __popc(smask);
      if (lane == 0) searches += __popc(smask);
      __syncwarp();
      if (qcount >= 32) { qcount -= 32; run_batch(queue + qcount, 32, std::integral_constant<int, 0>{}); }
    }
  };
#if B200_DUAL_TILE
  // two independent queries per thread and loop trip: twice the loads in flight in the latency-bound skip path
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += 2 * gridDim.x) {
    const uint32_t i0 = tile * kBlock + tid;
    const uint32_t tile1 = tile + gridDim.x;
    const uint32_t i1 = tile1 * kBlock + tid;
    const bool s0 = stream_one(i0, i0 < nd);
    const bool s1 = stream_one(i1, tile1 < ntiles && i1 < nd);
    enqueue(i0, s0);
    enqueue(i1, s1);
  }
#elif B200_TRIP_CHUNK > 1
  // Untested experiment for the next round (profiles/r01_sched_cert_experiment.md, section 2): a warp's consecutive
  // trips are kTripChunk NEIGHBOURING groups of 32 points instead of groups gridDim * 256 points apart, so a batch
  // that the queue assembles from two trips still covers one region of the model.
  {
    constexpr uint32_t kTripChunk = B200_TRIP_CHUNK;
    const uint32_t span = kBlock * kTripChunk;                     // points per block and stride
    const uint32_t nspans = (nd + span - 1) / span;
    for (uint32_t sp = blockIdx.x; sp < nspans; sp += gridDim.x)
      for (uint32_t c = 0; c < kTripChunk; ++c) {
        const uint32_t i = sp * span + (warp * kTripChunk + c) * 32u + lane;
        enqueue(i, stream_one(i, i < nd));
      }
  }
#else
  {
    // one loop for both modes: static = tile b, b + grid, ... of the block (warp w takes the tile's group w);
    // dynamic = the next group of the scan, whoever asks (the draw for the group after this one is issued before the
    // work and read after it, so its latency is hidden)
    const uint32_t ngroups = (nd + 31u) / 32u;
    uint32_t g = blockIdx.x * kWarps + warp;
    unsigned gnext = 0;
    if (dyn) {
      if (lane == 0) gnext = atomicAdd(stage2_counter + 3, 1u);
      g = __shfl_sync(0xffffffffu, gnext, 0);
    }
    while (g < ngroups) {
      if (dyn && lane == 0) gnext = atomicAdd(stage2_counter + 3, 1u);
      const uint32_t i = g * 32u + lane;
      enqueue(i, stream_one(i, i < nd));
      g = dyn ? __shfl_sync(0xffffffffu, gnext, 0) : g + gridDim.x * kWarps;
    }
  }
#endif
  wrp_mark(1);
  // leftovers: merge the warps' partial queues so the remaining searches run in full batches
  if (lane == 0) sm.left_count[warp] = qcount;
  __syncthreads();
  blk_mark(1);
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
    wrp_mark(2);
    // total < kWarps * 32 (every warp kept fewer than 32).  The remaining searches are spread over ALL warps: a
    // batch lasts as long as its slowest lane, and the block waits for its slowest batch, so eight batches of 12
    // finish earlier than three of 32 while five warps idle (late iterations: ~90 searches per block).
    {
      const int per = (total + kWarps - 1) / kWarps;
      const int base = warp * per;
      if (base < total) run_batch(sm.leftover + base, min(per, total - base), std::integral_constant<int, 1>{});
    }
    wrp_mark(3);
  }

  // block sums of the NS accumulator columns: warp w takes moments w, w + 8, ...; a lane adds 8 of the column's 256
  // thread entries, five shuffle steps finish it -- 3 x (8 loads + 10 shuffles) per warp instead of NS x 10 shuffles,
  // fixed order
  __syncthreads();
  for (int k = warp; k < NS; k += kWarps) {
    const double* col = &sm.acc[k][0];
    double v = 0.0;
    if (FIXP) {
      long long iv = 0;
#pragma unroll
      for (int u = 0; u < kBlock / 32; ++u) iv += __double_as_longlong(col[u * 32 + lane]);
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) iv += __shfl_xor_sync(0xffffffffu, iv, m);
      v = __longlong_as_double(iv);
    } else {
#pragma unroll
      for (int u = 0; u < kBlock / 32; ++u) v += col[u * 32 + lane];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    }
    // integer sums need no row per block: they are added straight into ONE row of totals (integer atomics are exact
    // and order-independent), which the last block only has to read -- the 444-row reduction leaves the serial tail
    if (FIXP) { if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(partials) + k, (unsigned long long)__double_as_longlong(v)); }
    else if (lane == 0) partials[(size_t)blockIdx.x * NS_MAX + k] = v;
  }
  blk_mark(2);
  if (lane == 0 && stage2) atomicAdd(stage2_counter, stage2);
  if (lane == 0 && searches) atomicAdd(stage2_counter + 1, searches);
  // ---- the last block to get here reduces all partials and runs the solve (no second launch)
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(stage2_counter + 2, 1u);
    is_last = ticket == gridDim.x - 1;
    if (is_last) { stage2_counter[2] = 0; stage2_counter[3] = 0; }   // ready for the next launch
  }
  __syncthreads();
  if (is_last) {
#if B200_PDL
    // every other block has exited: with this trigger the next launch's blocks are scheduled while the tail runs
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    __threadfence();
    solve_step(st, partials, (int)gridDim.x, FIXP, rms_log, npairs_log, stage2_log, stage2_counter, &sm.acc[0][0], comm);
  }
}

// ---- API path: SearchTree::getPtPairs over caller-supplied queries ---------------------------------
// xfs: source_alignxf and its inverse, passed BY VALUE (kernel parameter space): a call never depends on a staging
// buffer that a later call could overwrite before this one has run.  partials: [grid][8] = n, sum, cm[3], cd[3].
struct BatchXf { double S[16], Sinv[16]; };
template <bool PLANE>
__global__ void __launch_bounds__(kBlock)
nn_batch_kernel(GridDev model, const double* __restrict__ q_xyz, const double* __restrict__ q_nrm,
                size_t n, const __grid_constant__ BatchXf xfs, double maxdist2, int32_t* __restrict__ idx_out,
                double* __restrict__ d2_out, double* __restrict__ partials) {
  __shared__ SearchSmemSmall sm;
  __shared__ double S[16], Sinv[16];
  const int tid = threadIdx.x;
  if (tid < 16) { S[tid] = xfs.S[tid]; Sinv[tid] = xfs.Sinv[tid]; }
  __syncthreads();
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0;
  unsigned stage2 = 0;
  const size_t ntiles = (n + kBlock - 1) / kBlock;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t i = tile * kBlock + tid;
    const bool active = i < n;
    double tx = 0, ty = 0, tz = 0, sx = 0, sy = 0, sz = 0;
    if (active) {
      tx = q_xyz[3 * i]; ty = q_xyz[3 * i + 1]; tz = q_xyz[3 * i + 2];
      xf_apply_strict(Sinv, tx, ty, tz, sx, sy, sz);
    }
    int bj;
    double bd2;
    unsigned boidx;
    float unused_budget;
    nn_warp_search<true>(model, sm, active, sx, sy, sz, maxdist2, -1, 0.f, bj, bd2, boidx, unused_budget, stage2);
    if (active) {
      if (idx_out) idx_out[i] = bj >= 0 ? (int32_t)boidx : -1;
      if (d2_out) d2_out[i] = bj >= 0 ? bd2 : -1.0;
      if (bj >= 0) {
        const double2 pa = __ldg(reinterpret_cast<const double2*>(model.p64 + bj));
        const double pz = __ldg(reinterpret_cast<const double*>(model.p64 + bj) + 2);
        double p1x, p1y, p1z;
        xf_apply_strict(S, pa.x, pa.y, pz, p1x, p1y, p1z);
        if (PLANE) {
          double nx = q_nrm[3 * i], ny = q_nrm[3 * i + 1], nz = q_nrm[3 * i + 2];
          const double nl = sqrt(nx * nx + ny * ny + nz * nz);
          nx /= nl; ny /= nl; nz /= nl;
          const double dot = nx * (p1x - tx) + ny * (p1y - ty) + nz * (p1z - tz);
          p1x = nx * dot + tx; p1y = ny * dot + ty; p1z = nz * dot + tz;
        }
        const double ex = p1x - tx, ey = p1y - ty, ez = p1z - tz;
        acc[0] += 1.0;
        acc[1] += ex * ex + ey * ey + ez * ez;
        acc[2] += p1x; acc[3] += p1y; acc[4] += p1z;
        acc[5] += tx; acc[6] += ty; acc[7] += tz;
      }
    }
  }
  block_reduce_store_regs<8>(acc, partials + (size_t)blockIdx.x * 8);
}

// ---- LUM link sums: lum6DEuler::covarianceEuler (reference src/slam6d/lum6Deuler.cc:94-260) and, with QUAT,
// lum6DQuat::covarianceQuat (src/slam6d/lum6Dquat.cc:83-240, the 7-parameter form ELCH-SLERP uses) ---------------
// PASS 1: pairs of (model grid, data scan) -> neighbour cache + the sums that make MM and MZ (16 Euler / 18 quat).
// PASS 2: the pairs are re-formed from the cache (nothing moved) and the residual sum ss is taken with the
//         pose-difference estimate D the host solved from pass 1 -- the reference's two walks over `uk`.
// xfs: [0..15] data dalignxf, [16..31] model dalignxf, [32..47] its inverse, [48..54] D.
// partials: [grid][kLumSums].  Sums are taken in absolute coordinates like the reference (the linearisation of
// the LUM error is about the origin, so these moments are not shift-invariant).
constexpr int kLumSums = 18;
template <int PASS, bool QUAT>
__global__ void __launch_bounds__(kBlock, 2)
lum_link_kernel(GridDev model, const double4* __restrict__ dq, uint32_t nd, const double* __restrict__ xfs,
                double maxdist2, int* __restrict__ nn_cache, double* __restrict__ partials, int seeded) {
  __shared__ SearchSmemSmall sm;
  __shared__ double X[16], S[16], Sinv[16], D[7];
  const int tid = threadIdx.x;
  if (tid < 16) { X[tid] = xfs[tid]; S[tid] = xfs[16 + tid]; Sinv[tid] = xfs[32 + tid]; }
  if (tid < 7) D[tid] = xfs[48 + tid];
  __syncthreads();
  double acc[kLumSums];
#pragma unroll
  for (int k = 0; k < kLumSums; ++k) acc[k] = 0.0;
  unsigned stage2 = 0;
  const uint32_t ntiles = (nd + kBlock - 1) / kBlock;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint32_t i = tile * kBlock + tid;
    const bool active = i < nd;
    double tx = 0, ty = 0, tz = 0, sx = 0, sy = 0, sz = 0;
    if (active) {
      const double2 a = __ldg(reinterpret_cast<const double2*>(dq + i));
      const double z0 = __ldg(reinterpret_cast<const double*>(dq + i) + 2);
      xf_apply(X, a.x, a.y, z0, tx, ty, tz);
      xf_apply(Sinv, tx, ty, tz, sx, sy, sz);
    }
    int bj = -1;
    if (PASS == 1) {
      double bd2;
      unsigned boidx;
      float unused_budget;
      // seeded: nn_cache holds this link's neighbours from the previous LUM iteration (an exact upper bound on the
      // search radius; the result does not depend on it)
      const int seed = (seeded && active) ? nn_cache[i] : -1;
      nn_warp_search<true>(model, sm, active, sx, sy, sz, maxdist2, seed, 0.f, bj, bd2, boidx, unused_budget, stage2);
      if (active) nn_cache[i] = bj;
    } else if (active) {
      bj = nn_cache[i];
    }
    if (active && bj >= 0) {
      const double2 pa = __ldg(reinterpret_cast<const double2*>(model.p64 + bj));
      const double pz = __ldg(reinterpret_cast<const double*>(model.p64 + bj) + 2);
      double ax, ay, az;
      xf_apply(S, pa.x, pa.y, pz, ax, ay, az);           // ak = p1 (model side), bk = p2 = t (data side)
      const double x = (ax + tx) / 2.0, y = (ay + ty) / 2.0, z = (az + tz) / 2.0;
      const double dx = ax - tx, dy = ay - ty, dz = az - tz;
      if (PASS == 1) {
        acc[0] += 1.0;
        acc[1] += x; acc[2] += y; acc[3] += z;
        acc[4] += x * x + y * y; acc[5] += x * x + z * z; acc[6] += y * y + z * z;
        acc[7] += x * y; acc[8] += x * z; acc[9] += y * z;
        acc[10] += dx; acc[11] += dy; acc[12] += dz;
        if (!QUAT) {
          acc[13] += -z * dy + y * dz; acc[14] += -y * dx + x * dy; acc[15] += z * dx - x * dz;
        } else {   // MZ(5..7), MZ(4) and the trace sum xpypz of lum6Dquat.cc:146-164
          acc[13] += z * dy - y * dz; acc[14] += x * dz - z * dx; acc[15] += y * dx - x * dy;
          acc[16] += x * dx + y * dy + z * dz;
          acc[17] += x * x + y * y + z * z;
        }
      } else if (!QUAT) {
        const double rx = dx - (D[0] - y * D[4] + z * D[5]);
        const double ry = dy - (D[1] - z * D[3] + x * D[4]);
        const double rz = dz - (D[2] + y * D[3] - x * D[5]);
        acc[0] += rx * rx + ry * ry + rz * rz;
      } else {     // lum6Dquat.cc:203-205
        const double rx = dx - (D[0] + x * D[3] - z * D[5] + y * D[6]);
        const double ry = dy - (D[1] + y * D[3] + z * D[4] - x * D[6]);
        const double rz = dz - (D[2] + z * D[3] - y * D[4] + x * D[5]);
        acc[0] += rx * rx + ry * ry + rz * rz;
      }
    }
  }
  block_reduce_store_regs<kLumSums>(acc, partials + (size_t)blockIdx.x * kLumSums);
}

// current "xyz reduced" = X * original (Scan::transformReduced), written back in original row order
__global__ void scan_export_kernel(const double4* __restrict__ p64, const double4* __restrict__ nrm,
                                   const uint32_t* __restrict__ perm, uint32_t n,
                                   const double* __restrict__ xf /* X[16], Nm[9] */,
                                   double* __restrict__ xyz_out, double* __restrict__ nrm_out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t dst = perm[j];
  const double4 p = p64[j];
  double x, y, z;
  xf_apply(xf, p.x, p.y, p.z, x, y, z);
  xyz_out[3 * dst] = x; xyz_out[3 * dst + 1] = y; xyz_out[3 * dst + 2] = z;
  if (nrm_out && nrm) {
    const double4 q = nrm[j];
    const double* Nm = xf + 16;
    nrm_out[3 * dst] = Nm[0] * q.x + Nm[1] * q.y + Nm[2] * q.z;
    nrm_out[3 * dst + 1] = Nm[3] * q.x + Nm[4] * q.y + Nm[5] * q.z;
    nrm_out[3 * dst + 2] = Nm[6] * q.x + Nm[7] * q.y + Nm[8] * q.z;
  }
}

}  // namespace b200
