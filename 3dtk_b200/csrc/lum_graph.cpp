// lum_graph.cpp -- the global relaxation back-end around the LUM link kernel.
//
// Replaces, on the host side of the C ABI:
//   Graph::Graph(int nodes, double cldist2, int loopsize)   reference src/slam6d/graph.cc:108-127
//   lum6DEuler::FillGB3D                                    reference src/slam6d/lum6Deuler.cc:265-304
//   lum6DEuler::doGraphSlam6D                               reference src/slam6d/lum6Deuler.cc:314-479
//   Scan::transformToEuler (matrix bookkeeping only)        reference src/slam6d/scan.cc:1061-1083
//   Matrix4ToEuler                                          reference include/slam6d/globals.icc:540-578
// The per-link work (pairs, sums, C, CD) is b200icp_lum_link (lum_link_kernel); everything here is O(scans):
// the assembly of G and B, one dense Cholesky solve of the (6(n-1))^2 system and the pose updates.  The
// reference solves the same SPD system with CXSparse's cs_cholsol (graphSlam6D.cc:305-345; SuiteSparse is not
// vendored, version unpinned); its in-tree dense alternative graphSlam6D::solveCholesky (:245-293, the
// Numerical-Recipes choldc/cholsl pair) is what is restated here -- same solution up to rounding.
// Points are never moved: a scan's pose lives in transMat / dalignxf and the kernels apply dalignxf on load.
// This file only uses the public C ABI (no CUDA), so sharded callers can split the loop:
//   fill_gb(own links) -> all-reduce of [G|B] -> solve_update on every rank   (3dtk_b200/parallel.py).
#include "../../include/b200icp.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

extern "C" int b200icp_set_error_(int code, const char* msg);   // b200icp.cu: records the thread's last error

namespace {

// in-place inverse of a 6x6 by Gauss-Jordan with partial pivoting (newmat's Ha.i() in the reference)
bool invert6(double A[6][6]) {
  double W[6][12];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) { W[r][c] = A[r][c]; W[r][6 + c] = r == c ? 1.0 : 0.0; }
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(W[r][c]) > fabs(W[piv][c])) piv = r;
    if (W[piv][c] == 0.0) return false;
    if (piv != c) for (int k = 0; k < 12; ++k) { const double t = W[c][k]; W[c][k] = W[piv][k]; W[piv][k] = t; }
    const double d = 1.0 / W[c][c];
    for (int k = 0; k < 12; ++k) W[c][k] *= d;
    for (int r = 0; r < 6; ++r) {
      if (r == c) continue;
      const double f = W[r][c];
      if (f == 0.0) continue;
      for (int k = 0; k < 12; ++k) W[r][k] -= f * W[c][k];
    }
  }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) A[r][c] = W[r][6 + c];
  return true;
}

// choldc + cholsl (graphSlam6D.cc:245-293 -> Numerical Recipes): A = L L^T on the lower triangle, then two
// triangular solves.  A is row-major n x n and is overwritten.
bool cholesky_solve(int n, std::vector<double>& A, const double* b, double* x) {
  std::vector<double> diag(n);
  for (int i = 0; i < n; ++i) {
    for (int j = i; j < n; ++j) {
      double sum = A[(size_t)i * n + j];
      for (int k = i - 1; k >= 0; --k) sum -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      if (i == j) {
        // not positive definite; the reference's default solver (cs_cholsol, graphSlam6D.cc:330) fails on the same
        // condition -- no pivot floor, so weakly constrained but valid systems still solve
        if (!(sum > 0.0)) return false;
        diag[i] = sqrt(sum);
      } else {
        A[(size_t)j * n + i] = sum / diag[i];
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double sum = b[i];
    for (int k = i - 1; k >= 0; --k) sum -= A[(size_t)i * n + k] * x[k];
    x[i] = sum / diag[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double sum = x[i];
    for (int k = i + 1; k < n; ++k) sum -= A[(size_t)k * n + i] * x[k];
    x[i] = sum / diag[i];
  }
  return true;
}

}  // namespace

extern "C" {

void b200icp_matrix4_to_euler(const double m[16], double rPosTheta[3], double rPos[3]) {
  if (m[0] > 0.0) rPosTheta[1] = asin(m[8]);
  else rPosTheta[1] = M_PI - asin(m[8]);
  const double C = cos(rPosTheta[1]);
  if (fabs(C) > 0.005) {                       // no gimbal lock
    rPosTheta[0] = atan2(-m[9] / C, m[10] / C);
    rPosTheta[2] = atan2(-m[4] / C, m[0] / C);
  } else {
    rPosTheta[0] = 0.0;
    rPosTheta[2] = atan2(m[1], m[5]);
  }
  if (rPos) { rPos[0] = m[12]; rPos[1] = m[13]; rPos[2] = m[14]; }
}

int b200icp_graph_from_poses(const double* rpos, int n_scans, double cldist2, int loopsize, int* links,
                             int cap, int* n_links) {
  if (!rpos || !n_links || n_scans < 1) return b200icp_set_error_(B200ICP_EINVAL, "graph_from_poses: bad argument");
  int n = 0;
  auto push = [&](int a, int b) {
    if (links && n < cap) { links[2 * n] = a; links[2 * n + 1] = b; }
    ++n;
  };
  for (int i = 0; i < n_scans - 1; ++i) push(i, i + 1);
  for (int j = 0; j < n_scans; ++j)
    for (int k = j + 1; k < n_scans; ++k) {
      const double dx = rpos[3 * j] - rpos[3 * k], dy = rpos[3 * j + 1] - rpos[3 * k + 1],
                   dz = rpos[3 * j + 2] - rpos[3 * k + 2];
      // Dist2, globals.icc:237-245: (x2-x1)^2 summed left to right
      if (k - j > loopsize && dx * dx + dy * dy + dz * dz < cldist2) push(j, k);
    }
  *n_links = n;
  if (links && n > cap) return b200icp_set_error_(B200ICP_EINVAL, "graph_from_poses: links array too small");
  return B200ICP_OK;
}

int b200icp_graph_chain(int n_scans, int loop, int* links, int cap, int* n_links) {
  // Graph::Graph(int nScans, bool loop), graph.cc:76-105: (i, i+1) for every scan, the last one back to 0 when `loop`
  if (!n_links || n_scans < 0) return b200icp_set_error_(B200ICP_EINVAL, "graph_chain: bad argument");
  const int nl = loop ? n_scans : (n_scans > 0 ? n_scans - 1 : 0);
  *n_links = nl;
  if (!links) return B200ICP_OK;
  if (nl > cap) return b200icp_set_error_(B200ICP_EINVAL, "graph_chain: links array too small");
  for (int i = 0; i < nl; ++i) {
    links[2 * i] = i;
    links[2 * i + 1] = loop ? (i != nl - 1 ? i + 1 : 0) : i + 1;
  }
  return B200ICP_OK;
}

int b200icp_lum_fill_gb(b200icp_ctx* ctx, b200icp_scan* const* scans, int n_scans, const int* links,
                        int n_links, double max_dist_match2, double* G, double* B, uint64_t* npairs_out) {
  if (!ctx || !scans || !links || !G || !B || n_scans < 2 || n_links < 0)
    return b200icp_set_error_(B200ICP_EINVAL, "lum_fill_gb: bad argument");
  const int dim = 6 * (n_scans - 1);
  for (int l = 0; l < n_links; ++l) {
    const int first = links[2 * l], second = links[2 * l + 1];
    if (first < 0 || second < 0 || first >= n_scans || second >= n_scans || first == second)
      return b200icp_set_error_(B200ICP_EINVAL, "lum_fill_gb: link references a scan outside the graph");
    double C[36], CD[6];
    uint64_t np = 0;
    const int rc = b200icp_lum_link(ctx, scans[first], scans[second], max_dist_match2, C, CD, &np);
    if (rc != B200ICP_OK) return rc;
    if (npairs_out) npairs_out[l] = np;
    const int a = first - 1, b = second - 1;        // scan 0 is fixed (lum6Deuler.cc:273-274)
    if (a >= 0) {
      for (int r = 0; r < 6; ++r) {
        B[6 * a + r] += CD[r];
        for (int c = 0; c < 6; ++c) G[(size_t)(6 * a + r) * dim + 6 * a + c] += C[6 * r + c];
      }
    }
    if (b >= 0) {
      for (int r = 0; r < 6; ++r) {
        B[6 * b + r] -= CD[r];
        for (int c = 0; c < 6; ++c) G[(size_t)(6 * b + r) * dim + 6 * b + c] += C[6 * r + c];
      }
    }
    if (a >= 0 && b >= 0) {
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          G[(size_t)(6 * a + r) * dim + 6 * b + c] -= C[6 * r + c];
          G[(size_t)(6 * b + r) * dim + 6 * a + c] -= C[6 * r + c];
        }
    }
  }
  return B200ICP_OK;
}

int b200icp_lum_solve_update(b200icp_scan* const* scans, int n_scans, const double* G, const double* B,
                             double* sum_position_diff, b200icp_frames* frames) {
  if (!scans || !G || !B || n_scans < 2) return b200icp_set_error_(B200ICP_EINVAL, "lum_solve_update: bad argument");
  const int dim = 6 * (n_scans - 1);
  std::vector<double> A(G, G + (size_t)dim * dim), X(dim);
  if (!cholesky_solve(dim, A, B, X.data()))
    return b200icp_set_error_(B200ICP_ESTATE, "lum_solve_update: G is not positive definite (graph not connected to scan 0, or links without pairs)");
  double sum = 0.0;
  for (int i = 1; i < n_scans; ++i) {
    double T[16], dal[16], rPos[3], rTh[3];
    b200icp_scan_get_pose(scans[i], T, dal);
    b200icp_matrix4_to_euler(T, rTh, rPos);
    const double xa = rPos[0], ya = rPos[1], za = rPos[2];
    const double ctx_ = cos(rTh[0]), stx = sin(rTh[0]), cty = cos(rTh[1]), sty = sin(rTh[1]);
    double Ha[6][6] = {{0}};
    for (int k = 0; k < 6; ++k) Ha[k][k] = 1.0;
    Ha[0][4] = -za * ctx_ + ya * stx;
    Ha[0][5] = ya * cty * ctx_ + za * stx * cty;
    Ha[1][3] = za;
    Ha[1][4] = -xa * stx;
    Ha[1][5] = -xa * ctx_ * cty + za * sty;
    Ha[2][3] = -ya;
    Ha[2][4] = xa * ctx_;
    Ha[2][5] = -xa * cty * stx - ya * sty;
    Ha[3][5] = sty;
    Ha[4][4] = stx;
    Ha[4][5] = ctx_ * cty;
    Ha[5][4] = ctx_;
    Ha[5][5] = -stx * cty;
    if (!invert6(Ha)) return b200icp_set_error_(B200ICP_ESTATE, "lum_solve_update: singular pose Jacobian Ha");
    double result[6];
    for (int r = 0; r < 6; ++r) {
      double t = 0.0;
      for (int c = 0; c < 6; ++c) t += Ha[r][c] * X[6 * (i - 1) + c];
      result[r] = t;
    }
    double nPos[3], nTh[3];
    for (int k = 0; k < 3; ++k) { nPos[k] = rPos[k] - result[k]; nTh[k] = rTh[k] - result[k + 3]; }
    // Scan::transformToEuler: transform(M4inv(transMat)) then transform(EulerToMatrix4(new pose))
    // (two Scan::transform calls, so that the scan's cumulative normal map follows as transform3normal would move the
    //  normals -- a later CLOSEST_PLANE_SIMPLE / NAPX match or a download sees what the reference's scan holds)
    double tinv[16], alignxf[16];
    if (!b200icp_m4inv(T, tinv)) return b200icp_set_error_(B200ICP_ESTATE, "lum_solve_update: singular transMat");
    b200icp_euler_to_matrix4(nPos, nTh, alignxf);
    b200icp_scan_transform(scans[i], tinv);
    b200icp_scan_transform(scans[i], alignxf);
    if (frames) {   // transformToEuler(.., Scan::LUM, i != last ? 1 : 2), lum6Deuler.cc:447-451
      std::vector<double> all((size_t)16 * n_scans);
      for (int k = 0; k < n_scans; ++k) b200icp_scan_get_pose(scans[k], &all[(size_t)16 * k], nullptr);
      b200icp_frames_transform(frames, i, all.data(), B200ICP_FRAME_LUM, i != n_scans - 1 ? 1 : 2);
    }
    sum += sqrt(result[0] * result[0] + result[1] * result[1] + result[2] * result[2]);
  }
  if (sum_position_diff) *sum_position_diff = sum;
  return B200ICP_OK;
}

int b200icp_lum_graph_slam(b200icp_ctx* ctx, b200icp_scan* const* scans, int n_scans, const int* links,
                           int n_links, double max_dist_match2, int nr_it, double epsilon_lum,
                           double* ret_out, int* iterations_out, b200icp_frames* frames) {
  if (!ctx || !scans || !links) return b200icp_set_error_(B200ICP_EINVAL, "lum_graph_slam: NULL argument");
  if (n_scans <= 0) return b200icp_set_error_(B200ICP_EINVAL, "Zero scans in graph");   // lum6Deuler.cc:316-318
  double ret = DBL_MAX;
  int it = 0;
  if (n_scans >= 2) {
    const int dim = 6 * (n_scans - 1);
    std::vector<double> G((size_t)dim * dim), B(dim);
    for (; it < nr_it && ret > epsilon_lum; ++it) {
      std::fill(G.begin(), G.end(), 0.0);
      std::fill(B.begin(), B.end(), 0.0);
      int rc = b200icp_lum_fill_gb(ctx, scans, n_scans, links, n_links, max_dist_match2, G.data(), B.data(), nullptr);
      if (rc != B200ICP_OK) return rc;
      double sum = 0.0;
      rc = b200icp_lum_solve_update(scans, n_scans, G.data(), B.data(), &sum, frames);
      if (rc != B200ICP_OK) return rc;
      ret = sum / (double)n_scans;
    }
  }
  if (ret_out) *ret_out = ret;
  if (iterations_out) *iterations_out = it;
  return B200ICP_OK;
}

int b200icp_lum_graph_slam_sharded(b200icp_ctx* ctx, b200icp_scan* const* scans, int n_scans, const int* links,
                                   int n_links, double max_dist_match2, int nr_it, double epsilon_lum, int rank,
                                   int world, b200icp_allreduce_fn allreduce, void* user, double* ret_out,
                                   int* iterations_out, b200icp_frames* frames) {
  if (!ctx || !scans || !links) return b200icp_set_error_(B200ICP_EINVAL, "lum_graph_slam_sharded: NULL argument");
  if (n_scans <= 0) return b200icp_set_error_(B200ICP_EINVAL, "Zero scans in graph");   // lum6Deuler.cc:316-318
  if (world < 1 || rank < 0 || rank >= world)
    return b200icp_set_error_(B200ICP_EINVAL, "lum_graph_slam_sharded: need 0 <= rank < world");
  if (world > 1 && !allreduce)
    return b200icp_set_error_(B200ICP_EINVAL, "lum_graph_slam_sharded: world > 1 needs an all-reduce callback");
  // this rank's links: round-robin (consecutive links cost about the same; the reference hands them out with
  // `schedule(dynamic)`, lum6Deuler.cc:271)
  std::vector<int> mine;
  for (int l = rank; l < n_links; l += world) { mine.push_back(links[2 * l]); mine.push_back(links[2 * l + 1]); }
  double ret = DBL_MAX;
  int it = 0;
  if (n_scans >= 2) {
    const int dim = 6 * (n_scans - 1);
    std::vector<double> GB((size_t)dim * dim + dim);          // [G | B] packed: ONE all-reduce per LUM iteration
    for (; it < nr_it && ret > epsilon_lum; ++it) {
      std::fill(GB.begin(), GB.end(), 0.0);
      int rc = b200icp_lum_fill_gb(ctx, scans, n_scans, mine.data(), (int)(mine.size() / 2), max_dist_match2,
                                   GB.data(), GB.data() + (size_t)dim * dim, nullptr);
      if (rc != B200ICP_OK) return rc;
      if (world > 1 && allreduce(GB.data(), GB.size(), user) != 0)
        return b200icp_set_error_(B200ICP_ESTATE, "lum_graph_slam_sharded: the all-reduce callback failed");
      // every rank now holds the same system and runs the same O(scans) solve + pose update on its replica of the
      // scans: the replicas stay bit-identical without a broadcast
      double sum = 0.0;
      rc = b200icp_lum_solve_update(scans, n_scans, GB.data(), GB.data() + (size_t)dim * dim, &sum, frames);
      if (rc != B200ICP_OK) return rc;
      ret = sum / (double)n_scans;
    }
  }
  if (ret_out) *ret_out = ret;
  if (iterations_out) *iterations_out = it;
  return B200ICP_OK;
}

}  // extern "C"
