// TEST INFRASTRUCTURE ONLY -- extern "C" shell around the reference's OWN Scan / icp6D / graph-SLAM classes.
//
// oracle/_ref/libref3dtk_full.so links the UNMODIFIED reference sources scan.cc, basicScan.cc, metaScan.cc,
// kdMeta.cc, icp6D.cc, graph.cc, graphSlam6D.cc, lum6Deuler.cc, lum6Dquat.cc, Boctree.cc, point_type.cc,
// pointfilter.cc, allocator.cc, ann_kd.cc, io_types.cc, metrics.cc (plus everything libref3dtk.so holds), compiled
// where they lie with the header shims of oracle/shim/ (Boost thread / filesystem / interprocess / graph type,
// CXSparse cs.h, scan-server ManagedScan) and the link stand-ins of oracle/shim_impl.cc.  Nothing of the algorithms
// is restated here: scans are the reference's in-memory BasicScan (basicScan.cc:207-252, the ROS node's
// precedent src/ros/icp6Dwrapper.cc:130-147), reduction is Scan::calcReducedPoints + BOctTree, matching is
// icp6D::match / icp6D::doICP, relaxation is lum6DEuler::doGraphSlam6D, frames are what Scan::transform appends.
// Golden vectors for rows f1 / f2 / f3 and the icp6D_gpu adapter test come from here (tests/golden/make_full_golden.py).
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "slam6d/basicScan.h"
#include "slam6d/graph.h"
#include "slam6d/icp6D.h"
#include "slam6d/icp6Dapx.h"
#include "slam6d/icp6Ddual.h"
#include "slam6d/icp6Dhelix.h"
#include "slam6d/icp6Dnapx.h"
#include "slam6d/icp6Dortho.h"
#include "slam6d/icp6Dquat.h"
#include "slam6d/icp6Dsvd.h"
#include "slam6d/lum6Deuler.h"
#include "slam6d/lum6Dquat.h"
#include "slam6d/metaScan.h"
#include "slam6d/globals.icc"

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

struct CoutSilencer {   // the reference prints progress to std::cout
  std::streambuf* old;
  struct NullBuf : std::streambuf { int overflow(int c) override { return c; } } nb;
  CoutSilencer() : old(std::cout.rdbuf(&nb)) {}
  ~CoutSilencer() { std::cout.rdbuf(old); }
};

struct Held {
  std::vector<double> xyz;   // the caller's points (BasicScan copies them; kept for the pointer vector's lifetime)
  BasicScan* scan = nullptr;
};

std::vector<Scan*> as_scans(void** h, int n) {
  std::vector<Scan*> v(n);
  for (int i = 0; i < n; ++i) v[i] = ((Held*)h[i])->scan;
  return v;
}

}  // namespace

extern "C" {

// In-memory scan in its LOCAL frame with pose (rPos [cm], rPosTheta [rad]); voxel <= 0: no reduction
// (Scan::setReductionParameter / setSearchTreeParameter, scan.h:208-215).  Appended to Scan::allScans.
void* reff_scan_create(const double* xyz, long n, const double rPos[3], const double rPosTheta[3], double voxel,
                       int nrpts, int nns_method, int bucket) {
  Held* h = new Held();
  h->xyz.assign(xyz, xyz + 3 * n);
  std::vector<double*> pts(n);
  for (long i = 0; i < n; ++i) pts[i] = &h->xyz[3 * i];
  double p[3] = {rPos[0], rPos[1], rPos[2]}, t[3] = {rPosTheta[0], rPosTheta[1], rPosTheta[2]};
  CoutSilencer quiet;
  h->scan = new BasicScan(p, t, pts);
  h->scan->setReductionParameter(voxel, nrpts);
  h->scan->setSearchTreeParameter(nns_method, bucket);
  Scan::allScans.push_back(h->scan);
  return h;
}

// same, with per-point normals ("normal" field) and a normal-carrying reduction PointType (scan.cc:440-444,544-557)
void* reff_scan_create_normals(const double* xyz, const double* nrm, long n, const double rPos[3],
                               const double rPosTheta[3], double voxel, int nrpts) {
  Held* h = new Held();
  h->xyz.assign(xyz, xyz + 3 * n);
  std::vector<double*> pts(n);
  for (long i = 0; i < n; ++i) pts[i] = &h->xyz[3 * i];
  double p[3] = {rPos[0], rPos[1], rPos[2]}, t[3] = {rPosTheta[0], rPosTheta[1], rPosTheta[2]};
  CoutSilencer quiet;
  h->scan = new BasicScan(p, t, pts);
  DataNormal dn(h->scan->create("normal", sizeof(double) * 3 * (size_t)n));
  for (long i = 0; i < n; ++i) { dn[i][0] = nrm[3 * i]; dn[i][1] = nrm[3 * i + 1]; dn[i][2] = nrm[3 * i + 2]; }
  h->scan->setReductionParameter(voxel, nrpts, PointType(PointType::USE_NORMAL));
  h->scan->setSearchTreeParameter(simpleKD, 20);
  Scan::allScans.push_back(h->scan);
  return h;
}

void reff_srand(unsigned seed) { srand(seed); }

void reff_scan_free_all(void** hs, int n) {
  for (int i = 0; i < n; ++i) {
    Held* h = (Held*)hs[i];
    if (!h) continue;
    delete h->scan;
    delete h;
  }
  Scan::allScans.clear();
}

// number of rows of a DataXYZ field ("xyz", "xyz reduced", "xyz reduced original", "normal reduced" ...); copies
// up to cap rows into out (may be NULL).  Asking for a reduced field triggers the on-demand reduction.
long reff_scan_get(void* hv, const char* field, double* out, long cap) {
  Held* h = (Held*)hv;
  CoutSilencer quiet;
  DataXYZ d(h->scan->get(field));
  const long n = (long)d.size();
  if (out)
    for (long i = 0; i < n && i < cap; ++i) { out[3 * i] = d[i][0]; out[3 * i + 1] = d[i][1]; out[3 * i + 2] = d[i][2]; }
  return n;
}

void reff_scan_pose(void* hv, double transMat[16], double dalignxf[16], double rPos[3], double rPosTheta[3]) {
  Scan* s = ((Held*)hv)->scan;
  if (transMat) memcpy(transMat, s->get_transMat(), 16 * sizeof(double));
  if (dalignxf) memcpy(dalignxf, s->getDAlign(), 16 * sizeof(double));
  if (rPos) memcpy(rPos, s->get_rPos(), 3 * sizeof(double));
  if (rPosTheta) memcpy(rPosTheta, s->get_rPosTheta(), 3 * sizeof(double));
}

long reff_scan_frames(void* hv, double* mats, int* types, long cap) {
  Scan* s = ((Held*)hv)->scan;
  const long n = (long)s->getFrameCount();
  for (long i = 0; i < n && i < cap; ++i) {
    const double* m;
    Scan::AlgoType t;
    s->getFrame((size_t)i, m, t);
    if (mats) memcpy(mats + 16 * i, m, 16 * sizeof(double));
    if (types) types[i] = (int)t;
  }
  return n;
}

// Scan::transform(alignxf, type, islum) -- e.g. to move a scan before matching
void reff_scan_transform(void* hv, const double alignxf[16], int type, int islum) {
  CoutSilencer quiet;
  ((Held*)hv)->scan->transform(alignxf, (Scan::AlgoType)type, islum);
}

// icp6D::match (icp6D.cc:104-285), serial arm (library is built with OPENMP_NUM_THREADS = 1, no -fopenmp)
int reff_match(void* prev, void* cur, int algo, int pairing_mode, double max_dist_match, int max_num_iterations,
               double epsilonICP, int rnd, int nns_method) {
  icp6Dminimizer* mini = make_minimizer(algo);
  if (!mini) return -1;
  CoutSilencer quiet;
  icp6D icp(mini, max_dist_match, max_num_iterations, true, false, rnd, true, -1, epsilonICP, nns_method);
  const int it = icp.match(((Held*)prev)->scan, ((Held*)cur)->scan, (PairingMode)pairing_mode);
  delete mini;
  return it;
}

// icp6D::doICP (icp6D.cc:374-437)
int reff_do_icp(void** hs, int n, int algo, int pairing_mode, double max_dist_match, int max_num_iterations,
                double epsilonICP, int rnd, int meta, int eP, int max_num_metascans, int nns_method) {
  icp6Dminimizer* mini = make_minimizer(algo);
  if (!mini) return -1;
  CoutSilencer quiet;
  icp6D icp(mini, max_dist_match, max_num_iterations, true, meta != 0, rnd, eP != 0, -1, epsilonICP, nns_method,
            false, false, max_num_metascans);
  icp.doICP(as_scans(hs, n), (PairingMode)pairing_mode);
  delete mini;
  return 0;
}

// lum6DEuler::covarianceEuler (lum6Deuler.cc:94-260) / lum6DQuat::covarianceQuat (lum6Dquat.cc:83-...)
int reff_covariance(void* first, void* second, int quat, int nns_method, int rnd, double max_dist_match2, double* C,
                    double* CD) {
  CoutSilencer quiet;
  const int dim = quat ? 7 : 6;
  NEWMAT::Matrix Cm(dim, dim);
  NEWMAT::ColumnVector CDv(dim);
  Cm = 0.0;
  CDv = 0.0;
  if (quat) lum6DQuat::covarianceQuat(((Held*)first)->scan, ((Held*)second)->scan, nns_method, rnd, max_dist_match2, &Cm, &CDv);
  else lum6DEuler::covarianceEuler(((Held*)first)->scan, ((Held*)second)->scan, nns_method, rnd, max_dist_match2, &Cm, &CDv);
  for (int i = 0; i < dim; ++i) {
    CD[i] = CDv(i + 1);
    for (int j = 0; j < dim; ++j) C[dim * i + j] = Cm(i + 1, j + 1);
  }
  return dim;
}

// Graph(int nodes, double cldist2, int loopsize) (graph.cc:108-127) over Scan::allScans; links out as (from, to)
int reff_graph_from_poses(int nodes, double cldist2, int loopsize, int* links, int cap) {
  Graph g(nodes, cldist2, loopsize);
  const int n = g.getNrLinks();
  for (int i = 0; i < n && i < cap; ++i) { links[2 * i] = g.getLink(i, 0); links[2 * i + 1] = g.getLink(i, 1); }
  return n;
}

// lum6DEuler::doGraphSlam6D (lum6Deuler.cc:314-479) on an explicit link list; returns its return value (the
// summed pose change of the last iteration)
double reff_lum_euler(void** hs, int n, const int* links, int nlinks, int nr_it, double max_dist_match_lum,
                      double epsilon_lum, int nns_method) {
  icp6Dminimizer* mini = make_minimizer(1);
  CoutSilencer quiet;
  lum6DEuler lum(mini, max_dist_match_lum, 25.0, 50, true, false, 1, true, -1, 1e-7, nns_method, epsilon_lum);
  Graph g;
  for (int i = 0; i < nlinks; ++i) g.addLink(links[2 * i], links[2 * i + 1]);
  g.setNrScans(n);   // (addLink counts nodes as it meets them; the scan count is what the caller says, as Graph(netfile))
  const double r = lum.doGraphSlam6D(g, as_scans(hs, n), nr_it);
  delete mini;
  return r;
}

}  // extern "C"
