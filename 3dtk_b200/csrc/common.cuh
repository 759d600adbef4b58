// common.cuh -- device-side data layout of a scan and the per-match iteration state.
//
// HBM layout of one scan (all arrays sorted by grid cell, x fastest, then original row):
//   p32  float4[n]   xyz relative to the fp32 origin `c`, w = original row (uint bits)   16 B/pt
//   p64  double4[n]  xyz in the frame the grid was built in ("xyz reduced original"), w=0 32 B/pt
//   nrm  double4[n]  optional "normal reduced"                                            32 B/pt
//   cell_start uint32[ncells+1]  exclusive prefix of per-cell counts (dense table)         4 B/cell
//   perm uint32[n]   sorted position -> original row
// The fp32 copy is what the correspondence kernel streams; the fp64 copy is only gathered for the few
// candidates that survive the fp32 filter (exact verification) and for accepted pairs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "solve.h"

namespace b200 {

struct GridDev {
  double g0[3];   // min corner of the grid
  double c[3];    // origin of the fp32 relative coordinates (bbox centre)
  double h;       // cell edge
  double inv_h;
  double bbox_lo[3], bbox_hi[3];
  int nx, ny, nz;
  uint32_t n;
  float bmax;     // max |relative coordinate| over the scan's points (fp32 error bound input)
  const uint32_t* __restrict__ cell_start;
  const float4* __restrict__ p32;
  const double4* __restrict__ p64;
  const double4* __restrict__ nrm;
};

// State of one icp6D::match loop, resident on the device between the correspondence kernel and the
// solve kernel (reference: locals of icp6D::match, src/slam6d/icp6D.cc:117-123, plus Scan::transMat /
// Scan::dalignxf, src/slam6d/scan.cc:878-898).
struct IterState {
  double X[16];       // data scan dalignxf   (current = X * original)
  double Xprev[16];   // X before the last alignxf (motion of each query since the previous iteration)
  double T[16];       // data scan transMat
  double S[16];       // model scan dalignxf  (Source->dalignxf, scan.cc:1240)
  double Sinv[16];    // M4inv(S)             (searchTree.cc:109-110)
  double Nm[9];       // cumulative normal map, row-major: n_cur = Nm * n_original (scan.cc:864-869)
  double o[3];        // shift origin of the moment sums
  double alignxf[16]; // last alignxf
  double ret, prev_ret, prev_prev_ret;
  double eps;
  int iter;           // loop index of the iteration about to run
  int done;           // loop has exited
  int ret_iter;       // value icp6D::match returns
  int iters_run;      // iterations that produced a transform
  int algo;
  int napx_weighted;
  int max_iter;
  unsigned int stage2_last;
  unsigned int searches_last;  // full searches of the previous iteration (picks the hand-out mode of the next launch)
  int fixed_point;             // the launch's partial sums are fixed-point integers (point-to-point moment set)
  // fixed-point scales of the point-to-point moment sums (powers of two, chosen per match from the scene's extent so
  // that nd * max|addend| < 2^60) and their inverses, see FixAcc in icp_kernels.cuh
  double fix_s[NS_P2P], fix_inv[NS_P2P];
  double* pose_log;   // [max_iter][16] transMat after every iteration that produced a transform (frames, see
                      // b200icp_last_poses); may be NULL
};

// Peer mailboxes of a query-sharded match.  Every rank owns one mailbox in its HBM, mapped into
// all peers (cudaIpc across processes, peer access inside one process):
//   box [2 slots][kMaxRanks][NS_MAX] doubles   moments of rank r for the iteration using that slot
//   flag[2 slots][kMaxRanks]          u64      sequence number of the iteration whose moments are complete
constexpr int kMaxRanks = 8;
struct Mailbox {
  double box[2][kMaxRanks][NS_MAX];
  unsigned long long flag[2][kMaxRanks];
};
struct CommDev {
  int rank, world;                 // world <= 1: single-GPU match, nothing is exchanged
  unsigned long long seq_base;     // + loop index = sequence number of an iteration
  Mailbox* peer[kMaxRanks];        // peer[r] = rank r's mailbox as mapped in this process (peer[rank] = own)
};

}  // namespace b200
