// TEST INFRASTRUCTURE ONLY -- drives the reference-side adapters (3dtk_b200/host/: icp6D_gpu, GpuSearchTree) inside
// the reference's OWN classes.  oracle/_ref/libadapter3dtk_full.so = everything of libref3dtk_full.so (unmodified
// scan.cc, basicScan.cc, icp6D.cc ... + full_harness.cc) + the adapters + the product library.
//   reffa_match_gpu / reffa_do_icp_gpu   icp6D_gpu used through an icp6D* (virtual match; doICP is the base class's)
//   reffa_scan_create_gputree            a BasicScan whose search tree is a GpuSearchTree -- the one-line `case` of
//                                        INTEGRATION.md realised as an override of createSearchTreePrivate, so that
//                                        the reference's unmodified icp6D::match loop runs over the GPU tree
#include <cstring>
#include <iostream>
#include <vector>

#include "gpu_search_tree.h"
#include "icp6d_gpu.h"
#include "slam6d/basicScan.h"
#include "slam6d/icp6Dapx.h"
#include "slam6d/icp6Ddual.h"
#include "slam6d/icp6Dhelix.h"
#include "slam6d/icp6Dnapx.h"
#include "slam6d/icp6Dortho.h"
#include "slam6d/icp6Dquat.h"
#include "slam6d/icp6Dsvd.h"

namespace {

icp6Dminimizer* make_minimizer(int algo) {
  switch (algo) {
    case 1: return new icp6D_QUAT(true);
    case 2: return new icp6D_SVD(true);
    case 3: return new icp6D_ORTHO(true);
    case 4: return new icp6D_DUAL(true);
    case 5: return new icp6D_HELIX(true);
    case 6: return new icp6D_APX(true);
    case 10: return new icp6D_NAPX(true);
    default: return nullptr;
  }
}

struct CoutSilencer {
  std::streambuf* old;
  struct NullBuf : std::streambuf { int overflow(int c) override { return c; } } nb;
  CoutSilencer() : old(std::cout.rdbuf(&nb)) {}
  ~CoutSilencer() { std::cout.rdbuf(old); }
};

// BasicScan::createSearchTreePrivate (basicScan.cc:702-728) with the GPU tree as its `case`
class GpuTreeScan : public BasicScan {
 public:
  GpuTreeScan(double* rPos, double* rPosTheta, std::vector<double*> pts, double max_dist_hint)
      : BasicScan(rPos, rPosTheta, pts), hint_(max_dist_hint), rows_(nullptr) {}
  virtual ~GpuTreeScan() { delete rows_; }

 protected:
  virtual void createSearchTreePrivate() {
    DataXYZ xyz_orig(get("xyz reduced original"));
    delete rows_;
    rows_ = new PointerArray<double>(xyz_orig);     // the tree returns pointers into the scan's own array
    kd = new GpuSearchTree(rows_->get(), (int)xyz_orig.size(), hint_);
  }

 private:
  double hint_;
  PointerArray<double>* rows_;
};

struct Held {   // same layout as full_harness.cc's
  std::vector<double> xyz;
  BasicScan* scan = nullptr;
};

char g_err[512] = "";
void set_err(const char* m) { strncpy(g_err, m, sizeof g_err - 1); g_err[sizeof g_err - 1] = 0; }

}  // namespace

extern "C" {

const char* reffa_last_error() { return g_err; }

void* reffa_scan_create_gputree(const double* xyz, long n, const double rPos[3], const double rPosTheta[3],
                                double voxel, int nrpts, double max_dist_hint) {
  Held* h = new Held();
  h->xyz.assign(xyz, xyz + 3 * n);
  std::vector<double*> pts(n);
  for (long i = 0; i < n; ++i) pts[i] = &h->xyz[3 * i];
  double p[3] = {rPos[0], rPos[1], rPos[2]}, t[3] = {rPosTheta[0], rPosTheta[1], rPosTheta[2]};
  CoutSilencer quiet;
  h->scan = new GpuTreeScan(p, t, pts, max_dist_hint);
  h->scan->setReductionParameter(voxel, nrpts);
  h->scan->setSearchTreeParameter(simpleKD, 20);   // any valid type: the override above decides
  Scan::allScans.push_back(h->scan);
  return h;
}

// icp6D_gpu::match through the base-class pointer; out3 = {iterations_run, npairs_last, kernel_launches}
int reffa_match_gpu(void* prev, void* cur, int algo, int pairing_mode, double max_dist_match, int max_num_iterations,
                    double epsilonICP, int rnd, int anim, long* out3) {
  icp6Dminimizer* mini = make_minimizer(algo);
  if (!mini) return -1;
  CoutSilencer quiet;
  int it = -2;
  try {
    icp6D_gpu gpu(mini, max_dist_match, max_num_iterations, true, false, rnd, true, anim, epsilonICP, simpleKD);
    icp6D* icp = &gpu;
    it = icp->match(((Held*)prev)->scan, ((Held*)cur)->scan, (PairingMode)pairing_mode);
    if (out3) { out3[0] = gpu.last_result().iterations_run; out3[1] = (long)gpu.last_result().npairs_last;
                out3[2] = gpu.last_result().kernel_launches; }
  } catch (const std::exception& e) {
    set_err(e.what());
    it = -1000;
  }
  delete mini;
  return it;
}

// icp6D::doICP (base class, unmodified) with icp6D_gpu::match underneath
int reffa_do_icp_gpu(void** hs, int n, int algo, int pairing_mode, double max_dist_match, int max_num_iterations,
                     double epsilonICP, int rnd, int meta, int eP, int max_num_metascans) {
  icp6Dminimizer* mini = make_minimizer(algo);
  if (!mini) return -1;
  CoutSilencer quiet;
  int rc = 0;
  try {
    icp6D_gpu gpu(mini, max_dist_match, max_num_iterations, true, meta != 0, rnd, eP != 0, -1, epsilonICP, simpleKD,
                  false, false, max_num_metascans);
    std::vector<Scan*> v(n);
    for (int i = 0; i < n; ++i) v[i] = ((Held*)hs[i])->scan;
    icp6D* icp = &gpu;
    icp->doICP(v, (PairingMode)pairing_mode);
  } catch (const std::exception& e) {
    set_err(e.what());
    rc = -1000;
  }
  delete mini;
  return rc;
}

}  // extern "C"
