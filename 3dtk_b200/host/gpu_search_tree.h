// gpu_search_tree.h -- reference-side adapter: a 3DTK SearchTree backed by the B200 engine.
//
// Drop-in for KDtree (include/slam6d/kd.h) wherever the reference instantiates a search tree:
//   BasicScan::createSearchTreePrivate (src/slam6d/basicScan.cc:702-728) gets one more `case`
//   (see INTEGRATION.md).  Derives from the reference's own SearchTree (include/slam6d/searchTree.h:38),
//   so Scan::getPtPairs / getPtPairsParallel, lum6DEuler::covarianceEuler, ELCH ... use it unchanged.
//
//   FindClosest   one exact query through b200icp_find_closest (API compatibility; one launch per call)
//   getPtPairs    both virtual overloads are overridden: the whole [start,end) range goes to the device in
//                 ONE b200icp_nn_batch call, PtPairs are rebuilt on the host with the reference's own
//                 arithmetic (transform3, plane projection, centroid / sum accumulation, push order).
// Thread contract of the reference (icp6D.cc:159-166: same tree, distinct thread_num): one engine context
// (= CUDA stream + workspaces) per thread_num, created on first use.
//
// Compile inside a 3DTK tree or against its headers: -I<3dtk>/include -I<this repo>/include.
#pragma once
#include <mutex>
#include <vector>

#include "b200icp.h"
#include "slam6d/searchTree.h"

class GpuSearchTree : public SearchTree {
 public:
  // same leading arguments as KDtree::KDtree(double **pts, int n, int bucketSize); the tree keeps `pts`
  // (FindClosest returns pointers into the caller's array, like the k-d tree does)
  GpuSearchTree(double** pts, int n, double max_dist_hint = 0.0, double cell_edge = 0.0, int device = 0);
  virtual ~GpuSearchTree();

  virtual double* FindClosest(double* _p, double maxdist2, int threadNum = 0) const;

  virtual void getPtPairs(std::vector<PtPair>* pairs, double* source_alignxf, double* const* q_points,
                          unsigned int startindex, unsigned int endindex, int thread_num, int rnd,
                          double max_dist_match2, double& sum, double* centroid_m, double* centroid_d);

  virtual void getPtPairs(std::vector<PtPair>* pairs, double* source_alignxf, const DataXYZ& xyz_r,
                          const DataNormal& normal_r, unsigned int startindex, unsigned int endindex,
                          int thread_num, int rnd, double max_dist_match2, double& sum, double* centroid_m,
                          double* centroid_d, PairingMode pairing_mode = CLOSEST_POINT);

  const b200icp_scan* scan() const { return scan_; }

 private:
  b200icp_ctx* context(int thread_num) const;
  void pairs_from_batch(std::vector<PtPair>* pairs, double* source_alignxf, const double* q_xyz,
                        const double* q_nrm, size_t n, int thread_num, double max_dist_match2, double& sum,
                        double* centroid_m, double* centroid_d, PairingMode pairing_mode);

  double** pts_;
  int n_;
  int device_;
  b200icp_scan* scan_;
  mutable std::mutex mu_;
  mutable std::vector<b200icp_ctx*> ctx_;   // index = thread_num
};
