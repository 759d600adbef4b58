// gpu_search_tree.cc -- see gpu_search_tree.h.  Errors follow the reference's style for search-tree
// construction (std::runtime_error from the constructor, kdTreeImpl.h:86-88; basicScan.cc:723-726).
#include "gpu_search_tree.h"

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "slam6d/globals.icc"

namespace {
[[noreturn]] void raise(const char* what) {
  throw std::runtime_error(std::string("GpuSearchTree: ") + what + ": " + b200icp_last_error());
}
}  // namespace

GpuSearchTree::GpuSearchTree(double** pts, int n, double max_dist_hint, double cell_edge, int device)
    : pts_(pts), n_(n), device_(device), scan_(nullptr) {
  if (n <= 0) throw std::runtime_error("cannot create kdtree with zero points");  // same text as the k-d tree
  b200icp_ctx* c = context(0);
  std::vector<double> flat(3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    flat[3 * (size_t)i] = pts[i][0];
    flat[3 * (size_t)i + 1] = pts[i][1];
    flat[3 * (size_t)i + 2] = pts[i][2];
  }
  if (b200icp_scan_create(c, flat.data(), nullptr, (size_t)n, cell_edge, max_dist_hint, &scan_) != B200ICP_OK)
    raise("scan_create");
}

GpuSearchTree::~GpuSearchTree() {
  if (scan_ && !ctx_.empty() && ctx_[0]) b200icp_scan_destroy(ctx_[0], scan_);
  for (b200icp_ctx* c : ctx_)
    if (c) b200icp_destroy(c);
}

b200icp_ctx* GpuSearchTree::context(int thread_num) const {
  std::lock_guard<std::mutex> lock(mu_);
  if (thread_num < 0) thread_num = 0;
  if ((size_t)thread_num >= ctx_.size()) ctx_.resize(thread_num + 1, nullptr);
  if (!ctx_[thread_num] && b200icp_create(device_, &ctx_[thread_num]) != B200ICP_OK) raise("create");
  return ctx_[thread_num];
}

double* GpuSearchTree::FindClosest(double* _p, double maxdist2, int threadNum) const {
  int64_t idx = -1;
  if (b200icp_find_closest(context(threadNum), scan_, _p, maxdist2, &idx) != B200ICP_OK) raise("find_closest");
  return idx < 0 ? 0 : pts_[idx];
}

void GpuSearchTree::pairs_from_batch(std::vector<PtPair>* pairs, double* source_alignxf, const double* q_xyz,
                                     const double* q_nrm, size_t n, int thread_num, double max_dist_match2,
                                     double& sum, double* centroid_m, double* centroid_d,
                                     PairingMode pairing_mode) {
  if (pairing_mode == CLOSEST_POINT_ALONG_NORMAL_SIMPLE)
    throw std::runtime_error("Method FindClosestAlongDir is not implemented");  // searchTree.cc:23-29
  std::vector<int32_t> idx(n);
  const int mode = pairing_mode == CLOSEST_PLANE_SIMPLE ? B200ICP_CLOSEST_PLANE_SIMPLE : B200ICP_CLOSEST_POINT;
  if (b200icp_nn_batch(context(thread_num), scan_, q_xyz, mode ? q_nrm : nullptr, n, source_alignxf,
                       max_dist_match2, mode, idx.data(), nullptr, nullptr) != B200ICP_OK)
    raise("nn_batch");
  // rebuild the pairs exactly as SearchTree::getPtPairs does (searchTree.cc:120-181)
  double t[3], s[3], normal[3] = {0, 0, 0};
  pairs->reserve(pairs->size() + n);
  for (size_t i = 0; i < n; ++i) {
    if (idx[i] < 0) continue;
    t[0] = q_xyz[3 * i]; t[1] = q_xyz[3 * i + 1]; t[2] = q_xyz[3 * i + 2];
    if (pairing_mode != CLOSEST_POINT) {
      normal[0] = q_nrm[3 * i]; normal[1] = q_nrm[3 * i + 1]; normal[2] = q_nrm[3 * i + 2];
      Normalize3(normal);
    }
    transform3(source_alignxf, pts_[idx[i]], s);
    if (pairing_mode == CLOSEST_PLANE_SIMPLE) {
      double tmp[3], s_[3];
      sub3(s, t, tmp);
      double dot = Dot(normal, tmp);
      scal_mul3(normal, dot, tmp);
      add3(tmp, t, s_);
      s[0] = s_[0]; s[1] = s_[1]; s[2] = s_[2];
    }
    centroid_m[0] += s[0]; centroid_m[1] += s[1]; centroid_m[2] += s[2];
    centroid_d[0] += t[0]; centroid_d[1] += t[1]; centroid_d[2] += t[2];
    PtPair myPair(s, t, normal);
    double p12[3] = {myPair.p1.x - myPair.p2.x, myPair.p1.y - myPair.p2.y, myPair.p1.z - myPair.p2.z};
    sum += Len2(p12);
    pairs->push_back(myPair);
  }
}

void GpuSearchTree::getPtPairs(std::vector<PtPair>* pairs, double* source_alignxf, const DataXYZ& xyz_r,
                               const DataNormal& normal_r, unsigned int startindex, unsigned int endindex,
                               int thread_num, int rnd, double max_dist_match2, double& sum,
                               double* centroid_m, double* centroid_d, PairingMode pairing_mode) {
  if (endindex <= startindex) return;
  lock();
  const bool use_n = pairing_mode != CLOSEST_POINT;
  if (rnd > 1) {
    // keep the reference's consumption of the global rand() stream (searchTree.cc:118): pick first, batch after
    std::vector<double> q, qn;
    for (unsigned int i = startindex; i < endindex; i++) {
      if (rand(rnd) != 0) continue;
      q.insert(q.end(), {xyz_r[i][0], xyz_r[i][1], xyz_r[i][2]});
      if (use_n) qn.insert(qn.end(), {normal_r[i][0], normal_r[i][1], normal_r[i][2]});
    }
    if (!q.empty())
      pairs_from_batch(pairs, source_alignxf, q.data(), use_n ? qn.data() : nullptr, q.size() / 3, thread_num,
                       max_dist_match2, sum, centroid_m, centroid_d, pairing_mode);
  } else {
    // DataXYZ / DataNormal are contiguous double[3] rows (include/slam6d/data_types.h)
    pairs_from_batch(pairs, source_alignxf, &xyz_r[startindex][0], use_n ? &normal_r[startindex][0] : nullptr,
                     (size_t)(endindex - startindex), thread_num, max_dist_match2, sum, centroid_m, centroid_d,
                     pairing_mode);
  }
  unlock();
}

void GpuSearchTree::getPtPairs(std::vector<PtPair>* pairs, double* source_alignxf, double* const* q_points,
                               unsigned int startindex, unsigned int endindex, int thread_num, int rnd,
                               double max_dist_match2, double& sum, double* centroid_m, double* centroid_d) {
  if (endindex <= startindex) return;
  lock();
  std::vector<double> q;
  q.reserve(3 * (size_t)(endindex - startindex));
  for (unsigned int i = startindex; i < endindex; i++) {
    if (rnd > 1 && rand(rnd) != 0) continue;
    q.insert(q.end(), {q_points[i][0], q_points[i][1], q_points[i][2]});
  }
  // this overload builds PtPair(s, t) without a normal (searchTree.cc:73); CLOSEST_POINT pairs carry zeros too
  if (!q.empty())
    pairs_from_batch(pairs, source_alignxf, q.data(), nullptr, q.size() / 3, thread_num, max_dist_match2, sum,
                     centroid_m, centroid_d, CLOSEST_POINT);
  unlock();
}
