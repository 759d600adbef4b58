// TEST INFRASTRUCTURE ONLY -- never linked into, imported or called by the product path.
//
// ref_harness.cc: thin extern "C" shell around the *unmodified* 3DTK reference objects
// (kd.cc, searchTree.cc, icp6D{quat,svd,apx,napx}.cc, normals.cc + vendored newmat/ANN), which
// oracle/Makefile compiles straight from /root/reference into oracle/_ref/.  The harness itself
// only restates the Boost-dependent glue that cannot be compiled here:
//   * Scan::getPtPairs / getPtPairsParallel     (reference src/slam6d/scan.cc:1220-1353)
//   * Scan::transformReduced / transformMatrix  (reference src/slam6d/scan.cc:851-898)
//   * icp6D::match loop, serial and OpenMP arms (reference src/slam6d/icp6D.cc:104-285)
// Everything numerical (k-d tree build/search, SearchTree::getPtPairs, the four Align functions,
// Align_Parallel, k-NN PCA normals, M4inv / MMult / transform3) is the reference's own code.
//
// Users: tests/ (validating oracle/oracle_icp.cpp and generating tests/golden/*) and
// bench.py's cpu_baseline / --impl reference legs.

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <iostream>
#include <streambuf>

#include "slam6d/kd.h"
#include "slam6d/searchTree.h"
#include "slam6d/BruteForceNotATree.h"
#include "slam6d/icp6Dminimizer.h"
#include "slam6d/icp6Dquat.h"
#include "slam6d/icp6Dsvd.h"
#include "slam6d/icp6Dapx.h"
#include "slam6d/icp6Dortho.h"
#include "slam6d/icp6Ddual.h"
#include "slam6d/icp6Dhelix.h"
#include "slam6d/icp6Dnapx.h"
#include "slam6d/normals.h"
#include "slam6d/globals.icc"
#include "newmat/newmat.h"
#include "newmat/newmatap.h"
using namespace NEWMAT;

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct RefTree {
  std::vector<double> xyz;      // private copy: the tree keeps raw pointers into it
  std::vector<double*> rows;
  SearchTree* tree = nullptr;
  int kind = 0;                 // 0 = KDtree (nns simpleKD), 3 = BruteForceNotATree
};

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

// M4inv prints to std::cout on singular input; keep test logs clean.
struct CoutSilencer {
  std::streambuf* old;
  struct NullBuf : std::streambuf { int overflow(int c) override { return c; } } nb;
  CoutSilencer() : old(std::cout.rdbuf(&nb)) {}
  ~CoutSilencer() { std::cout.rdbuf(old); }
};

}  // namespace

extern "C" {

int ref_max_threads() {
#ifdef _OPENMP
  return std::min(omp_get_max_threads(), (int)MAX_OPENMP_NUM_THREADS);
#else
  return 1;
#endif
}

// ---- search structure (reference: basicScan.cc:702-728 factory, kd.cc, BruteForceNotATree.cc)
void* ref_tree_create(const double* xyz, long n, int nns_kind, int bucket) {
  RefTree* t = new RefTree();
  t->xyz.assign(xyz, xyz + 3 * n);
  t->rows.resize(n);
  for (long i = 0; i < n; ++i) t->rows[i] = &t->xyz[3 * i];
  t->kind = nns_kind;
  try {
    if (nns_kind == 3) t->tree = new BruteForceNotATree(t->rows.data(), (int)n);
    else t->tree = new KDtree(t->rows.data(), (int)n, bucket > 0 ? bucket : 20);
  } catch (const std::exception& e) {
    delete t;
    return nullptr;
  }
  return t;
}

void ref_tree_free(void* h) {
  RefTree* t = (RefTree*)h;
  if (!t) return;
  delete t->tree;
  delete t;
}

// index of the closest model point (into the array given to ref_tree_create) or -1.
// BruteForceNotATree copies its points, so its hit is mapped back by value search over rows.
long ref_find_closest(void* h, const double* q, double maxdist2, int thread_num) {
  RefTree* t = (RefTree*)h;
  double p[3] = {q[0], q[1], q[2]};
  double* c = t->tree->FindClosest(p, maxdist2, thread_num);
  if (!c) return -1;
  if (t->kind != 3) return (long)((c - t->xyz.data()) / 3);
  for (size_t i = 0; i < t->rows.size(); ++i)
    if (t->rows[i][0] == c[0] && t->rows[i][1] == c[1] && t->rows[i][2] == c[2]) return (long)i;
  return -2;
}

void ref_find_closest_batch(void* h, const double* q, long nq, double maxdist2, int* idx_out,
                            int nthreads) {
  RefTree* t = (RefTree*)h;
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
  if (nthreads > (int)MAX_OPENMP_NUM_THREADS) nthreads = MAX_OPENMP_NUM_THREADS;
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (long i = 0; i < nq; ++i) {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double p[3] = {q[3 * i], q[3 * i + 1], q[3 * i + 2]};
    double* c = t->tree->FindClosest(p, maxdist2, tid);
    idx_out[i] = c ? (int)((c - t->xyz.data()) / 3) : -1;
  }
}

// k nearest neighbours through KDtree::kNearestNeighbors (kd.cc:102-135); returns count written.
int ref_knn(void* h, const double* q, int k, double* out_xyz) {
  RefTree* t = (RefTree*)h;
  KDtree* kd = dynamic_cast<KDtree*>(t->tree);
  if (!kd) return -1;
  double p[3] = {q[0], q[1], q[2]};
  std::vector<Point> r = kd->kNearestNeighbors(p, k, 0);
  for (size_t i = 0; i < r.size(); ++i) {
    out_xyz[3 * i] = r[i].x; out_xyz[3 * i + 1] = r[i].y; out_xyz[3 * i + 2] = r[i].z;
  }
  return (int)r.size();
}

// ---- SearchTree::getPtPairs (searchTree.cc:92-188), the DataXYZ overload, run as-is.
// p1/p2/nrm receive up to (end-start) rows; *sum, cm[3], cd[3] are ADDED to (caller zeroes).
long ref_get_pt_pairs(void* h, const double* source_alignxf, const double* data_xyz,
                      const double* data_nrm, long start, long end, int thread_num, int rnd,
                      double maxdist2, int pairing_mode, double* p1, double* p2, double* nrm,
                      double* sum, double* cm, double* cd) {
  RefTree* t = (RefTree*)h;
  std::vector<PtPair> pairs;
  double xf[16];
  memcpy(xf, source_alignxf, sizeof xf);
  DataXYZ xyz_r(DataPointer((unsigned char*)data_xyz, 0));
  DataNormal nrm_r(DataPointer((unsigned char*)data_nrm, 0));
  CoutSilencer quiet;
  t->tree->getPtPairs(&pairs, xf, xyz_r, nrm_r, (unsigned)start, (unsigned)end, thread_num, rnd,
                      maxdist2, *sum, cm, cd, (PairingMode)pairing_mode);
  for (size_t i = 0; i < pairs.size(); ++i) {
    p1[3 * i] = pairs[i].p1.x; p1[3 * i + 1] = pairs[i].p1.y; p1[3 * i + 2] = pairs[i].p1.z;
    p2[3 * i] = pairs[i].p2.x; p2[3 * i + 1] = pairs[i].p2.y; p2[3 * i + 2] = pairs[i].p2.z;
    if (nrm) {
      nrm[3 * i] = pairs[i].p2.nx; nrm[3 * i + 1] = pairs[i].p2.ny; nrm[3 * i + 2] = pairs[i].p2.nz;
    }
  }
  return (long)pairs.size();
}

// ---- icp6Dminimizer::Align (icp6Dquat.cc:38, icp6Dsvd.cc:38, icp6Dapx.cc:35, icp6Dnapx.cc:34)
double ref_align(int algo, long n, const double* p1, const double* p2, const double* nrm,
                 const double* cm, const double* cd, double* alignxf) {
  icp6Dminimizer* m = make_minimizer(algo);
  if (!m) return -2.0;
  std::vector<PtPair> pairs(n);
  for (long i = 0; i < n; ++i) {
    double a[3] = {p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]};
    double b[3] = {p2[3 * i], p2[3 * i + 1], p2[3 * i + 2]};
    if (nrm) {
      double c[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
      pairs[i] = PtPair(a, b, c);
    } else {
      pairs[i] = PtPair(a, b);
    }
  }
  M4identity(alignxf);
  CoutSilencer quiet;
  double r = m->Align(pairs, alignxf, cm, cd);
  delete m;
  return r;
}

// ---- tiny math exports used to pin the restatement's helpers (globals.icc)
int ref_m4inv(const double* in, double* out) { CoutSilencer q; return M4inv(in, out); }
void ref_mmult(const double* a, const double* b, double* out) { MMult(a, b, out); }
void ref_euler_to_matrix4(const double* pos, const double* theta, double* out) {
  EulerToMatrix4(pos, theta, out);
}
void ref_matrix4_to_euler(const double* m, double* theta, double* pos) {
  Matrix4ToEuler(m, theta, pos);
}

// ---- icp6D::match (icp6D.cc:104-285).  `data_xyz` (and `data_nrm`) are moved in place like
// Scan::transformReduced does; `data_transmat` / `data_dalignxf` play Scan::transMat /
// Scan::dalignxf of the *data* scan; `model_dalignxf` is Source->dalignxf (scan.cc:1240).
// parallel_threads == 0 -> serial arm (icp6D.cc:224-244, the parity oracle);
// parallel_threads  > 0 -> OpenMP arm (icp6D.cc:129-222) with that many threads;
// parallel_threads  < 0 -> the serial arm's arithmetic with the neighbour search spread over -parallel_threads threads
//                          (OpenMP build only): bit-identical to the serial arm, for full-size parity checks.
// Returns the value `match` returns (the loop index at exit); rms_out/npairs_out get one entry
// per executed iteration, *iters_done their count.
int ref_match(void* model_tree, const double* model_dalignxf, double* data_xyz, double* data_nrm,
              long nd, double* data_transmat, double* data_dalignxf, int algo, int pairing_mode,
              double max_dist_match, int max_num_iterations, double epsilonICP, int rnd,
              int parallel_threads, double* rms_out, long* npairs_out, int* iters_done,
              double* ms_after_first) {
  RefTree* t = (RefTree*)model_tree;
  icp6Dminimizer* mini = make_minimizer(algo);
  *iters_done = 0;
  if (ms_after_first) *ms_after_first = 0.0;
  if (!mini) return -1;
  CoutSilencer quiet;
  const double max_dist_match2 = sqr(max_dist_match);
  double src_xf[16];
  memcpy(src_xf, model_dalignxf, sizeof src_xf);
  DataXYZ xyz_r(DataPointer((unsigned char*)data_xyz, 0));
  DataNormal nrm_r(DataPointer((unsigned char*)data_nrm, 0));
  const bool has_normals = data_nrm != nullptr;

  if (max_num_iterations == 0) { delete mini; return 0; }

  double ret = 0.0, prev_ret = 0.0, prev_prev_ret = 0.0;
  int iter = 0;
  double alignxf[16];
  unsigned long t0 = GetCurrentTimeInMilliSec();
  struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);

  for (iter = 0; iter < max_num_iterations; iter++) {
    prev_prev_ret = prev_ret;
    prev_ret = ret;
    if (iter == 1) { t0 = GetCurrentTimeInMilliSec(); clock_gettime(CLOCK_MONOTONIC, &ts0); }
    long npairs = 0;

    if (parallel_threads > 0) {
#ifdef _OPENMP
      const int T = std::min(parallel_threads, (int)MAX_OPENMP_NUM_THREADS);
      const int step = (int)ceil(nd / (double)T);
      std::vector<std::vector<PtPair> > pairs(T);
      std::vector<double> sum(T, 0.0);
      std::vector<unsigned int> n(T, 0);
      typedef double V3[3];
      typedef double V9[9];
      std::vector<double> cm_s(3 * T, 0.0), cd_s(3 * T, 0.0), Si_s(9 * T, 0.0);
      V3* centroid_m = (V3*)cm_s.data();
      V3* centroid_d = (V3*)cd_s.data();
      V9* Si = (V9*)Si_s.data();
#pragma omp parallel num_threads(T)
      {
        const int tn = omp_get_thread_num();
        // Scan::getPtPairsParallel, non-meta branch (scan.cc:1328-1352)
        long lo = (long)tn * step, hi = (tn == T - 1) ? nd : (long)step * tn + step;
        if (lo > nd) lo = nd;
        if (hi > nd) hi = nd;
        t->tree->getPtPairs(&pairs[tn], src_xf, xyz_r, nrm_r, (unsigned)lo, (unsigned)hi, tn, rnd,
                            max_dist_match2, sum[tn], centroid_m[tn], centroid_d[tn],
                            (PairingMode)pairing_mode);
        size_t sz = pairs[tn].size();
        if (sz != 0)
          for (int i = 0; i < 3; ++i) { centroid_m[tn][i] /= sz; centroid_d[tn][i] /= sz; }
        n[tn] = (unsigned)sz;
        if (algo == 1 || algo == 2) {   // icp6D.cc:170-191, formula (6)
          for (unsigned i = 0; i < n[tn]; i++) {
            const PtPair& pr = pairs[tn][i];
            double pp[3] = {pr.p1.x - centroid_m[tn][0], pr.p1.y - centroid_m[tn][1],
                            pr.p1.z - centroid_m[tn][2]};
            double qq[3] = {pr.p2.x - centroid_d[tn][0], pr.p2.y - centroid_d[tn][1],
                            pr.p2.z - centroid_d[tn][2]};
            for (int a = 0; a < 3; ++a)
              for (int b = 0; b < 3; ++b) Si[tn][3 * a + b] += pp[a] * qq[b];
          }
        }
      }
      for (int i = 0; i < T; ++i) npairs += n[i];
      if (npairs > 3) {
        if (algo == 1 || algo == 2) {
          ret = mini->Align_Parallel(T, n.data(), sum.data(), centroid_m, centroid_d, Si, alignxf);
        } else {
          // algo 6 needs a compile-time OPENMP_NUM_THREADS-sized array (icp6Dapx.cc:148); algo 10
          // has no parallel arm at all (icp6D.cc:215-218).  Not offered by this harness.
          delete mini;
          return -3;
        }
      }
#else
      delete mini;
      return -4;
#endif
    } else {
      // serial arm: Scan::getPtPairs (scan.cc:1220-1260) + Align
      double centroid_m[3] = {0, 0, 0}, centroid_d[3] = {0, 0, 0};
      std::vector<PtPair> pairs;
#ifdef _OPENMP
      if (parallel_threads < 0 && rnd <= 1) {
        // SERIAL SEMANTICS, PARALLEL SEARCH (parity oracle for full-size pairs): the reference's getPtPairs runs
        // on contiguous chunks of the data scan in -parallel_threads threads; the chunks' pair vectors are then
        // concatenated in chunk order -- the order the serial loop produces -- and the running sums of
        // getPtPairs (searchTree.cc:164-177) are re-taken over that list in order, so `pairs`, `centroid_*` and
        // `ret` hold bit for bit what one serial getPtPairs call leaves; Align below is the serial one.
        const int T = std::min(-parallel_threads, (int)MAX_OPENMP_NUM_THREADS);
        const long step = (nd + T - 1) / T;
        std::vector<std::vector<PtPair> > part(T);
#pragma omp parallel num_threads(T)
        {
          const int tn = omp_get_thread_num();
          long lo = std::min((long)tn * step, nd), hi = std::min(lo + step, nd);
          double sum_t = 0.0, cm_t[3] = {0, 0, 0}, cd_t[3] = {0, 0, 0};
          t->tree->getPtPairs(&part[tn], src_xf, xyz_r, nrm_r, (unsigned)lo, (unsigned)hi, tn, rnd,
                              max_dist_match2, sum_t, cm_t, cd_t, (PairingMode)pairing_mode);
        }
        size_t total = 0;
        for (int i = 0; i < T; ++i) total += part[i].size();
        pairs.reserve(total);
        for (int i = 0; i < T; ++i) {
          pairs.insert(pairs.end(), part[i].begin(), part[i].end());
          std::vector<PtPair>().swap(part[i]);
        }
        for (size_t i = 0; i < pairs.size(); ++i) {
          const PtPair& pr = pairs[i];
          centroid_m[0] += pr.p1.x; centroid_m[1] += pr.p1.y; centroid_m[2] += pr.p1.z;
          centroid_d[0] += pr.p2.x; centroid_d[1] += pr.p2.y; centroid_d[2] += pr.p2.z;
          double p12[3] = {pr.p1.x - pr.p2.x, pr.p1.y - pr.p2.y, pr.p1.z - pr.p2.z};
          ret += Len2(p12);
        }
      } else
#endif
      t->tree->getPtPairs(&pairs, src_xf, xyz_r, nrm_r, 0u, (unsigned)nd, 0, rnd, max_dist_match2,
                          ret, centroid_m, centroid_d, (PairingMode)pairing_mode);
      size_t sz = pairs.size();
      if (sz != 0)
        for (int i = 0; i < 3; ++i) { centroid_m[i] /= sz; centroid_d[i] /= sz; }
      npairs = (long)sz;
      if (sz > 3) ret = mini->Align(pairs, alignxf, centroid_m, centroid_d);
      else break;
    }

    rms_out[*iters_done] = ret;
    npairs_out[*iters_done] = npairs;
    (*iters_done)++;

    // CurrentScan->transform(alignxf, ...) : transformReduced + transformMatrix
    for (long i = 0; i < nd; ++i) transform3(alignxf, data_xyz + 3 * i);
    if (has_normals)
      for (long i = 0; i < nd; ++i) transform3normal(alignxf, data_nrm + 3 * i);
    double tmp[16];
    MMult(alignxf, data_transmat, tmp);
    memcpy(data_transmat, tmp, sizeof tmp);
    MMult(alignxf, data_dalignxf, tmp);
    memcpy(data_dalignxf, tmp, sizeof tmp);

    if (((fabs(ret - prev_ret) < epsilonICP) && (fabs(ret - prev_prev_ret) < epsilonICP)) ||
        (iter == max_num_iterations - 1)) {
      break;
    }
  }
  struct timespec ts1; clock_gettime(CLOCK_MONOTONIC, &ts1);
  if (ms_after_first)
    *ms_after_first = (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6;
  (void)t0;
  delete mini;
  return iter;
}

// ---- normals: exact k-NN PCA (normals.cc:220-295 + calculateNormal :518-558)
void ref_normals_knn(const double* xyz, long n, int k, const double* rPos, double* out) {
  std::vector<Point> pts, nrm;
  pts.reserve(n);
  for (long i = 0; i < n; ++i) pts.push_back(Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
  calculateNormalsKNN(nrm, pts, k, rPos);
  for (long i = 0; i < n; ++i) { out[3 * i] = nrm[i].x; out[3 * i + 1] = nrm[i].y; out[3 * i + 2] = nrm[i].z; }
}

}  // extern "C"

// ---- lum6DEuler::covarianceEuler (lum6Deuler.cc:94-260).  lum6Deuler.cc cannot be compiled here (its
// header chain pulls in CXSparse), so the body is restated over the reference's getPtPairs and the
// reference's newmat (Matrix::i()) -- a weaker pin than the functions above, and labelled as such.
extern "C" long ref_lum_link(void* h, const double* source_alignxf, const double* data_xyz, long nd,
                             double maxdist2, double* C_out, double* CD_out) {
  RefTree* t = (RefTree*)h;
  std::vector<PtPair> uk;
  double xf[16];
  memcpy(xf, source_alignxf, sizeof xf);
  DataXYZ xyz_r(DataPointer((unsigned char*)data_xyz, 0));
  DataNormal nrm_r(DataPointer((unsigned char*)0, 0));
  double dsum = 0, dcm[3] = {0, 0, 0}, dcd[3] = {0, 0, 0};
  {
    CoutSilencer quiet;
    t->tree->getPtPairs(&uk, xf, xyz_r, nrm_r, 0u, (unsigned)nd, 0, 1, maxdist2, dsum, dcm, dcd, CLOSEST_POINT);
  }
  int m = (int)uk.size();
  for (int i = 0; i < 36; ++i) C_out[i] = 0.0;
  for (int i = 0; i < 6; ++i) CD_out[i] = 0.0;
  if (m <= 2) return m;
  double x, y, z, sx, sy, sz, xy, yz, xz, ypz, xpz, xpy, dx, dy, dz, ss;
  ColumnVector D(6), MZ(6);
  Matrix MM(6, 6);
  MZ = 0.0; MM = 0.0;
  sx = sy = sz = xy = yz = xz = ypz = xpz = xpy = ss = 0.0;
  for (int j = 0; j < m; j++) {
    Point ak = uk[j].p1, bk = uk[j].p2;
    x = (ak.x + bk.x) / 2.0; y = (ak.y + bk.y) / 2.0; z = (ak.z + bk.z) / 2.0;
    dx = ak.x - bk.x; dy = ak.y - bk.y; dz = ak.z - bk.z;
    sx += x; sy += y; sz += z;
    xpy += x * x + y * y; xpz += x * x + z * z; ypz += y * y + z * z;
    xy += x * y; xz += x * z; yz += y * z;
    MZ(1) += dx; MZ(2) += dy; MZ(3) += dz;
    MZ(4) += -z * dy + y * dz; MZ(5) += -y * dx + x * dy; MZ(6) += z * dx - x * dz;
  }
  MM(1, 1) = MM(2, 2) = MM(3, 3) = m;
  MM(4, 4) = ypz; MM(5, 5) = xpy; MM(6, 6) = xpz;
  MM(1, 5) = MM(5, 1) = -sy; MM(1, 6) = MM(6, 1) = sz;
  MM(2, 4) = MM(4, 2) = -sz; MM(2, 5) = MM(5, 2) = sx;
  MM(3, 4) = MM(4, 3) = sy;  MM(3, 6) = MM(6, 3) = -sx;
  MM(4, 5) = MM(5, 4) = -xz; MM(4, 6) = MM(6, 4) = -xy; MM(5, 6) = MM(6, 5) = -yz;
  D = MM.i() * MZ;
  for (int j = 0; j < m; j++) {
    Point ak = uk[j].p1, bk = uk[j].p2;
    x = (ak.x + bk.x) / 2.0; y = (ak.y + bk.y) / 2.0; z = (ak.z + bk.z) / 2.0;
    ss += sqr(ak.x - bk.x - (D(1) - y * D(5) + z * D(6))) + sqr(ak.y - bk.y - (D(2) - z * D(4) + x * D(5))) +
          sqr(ak.z - bk.z - (D(3) + y * D(4) - x * D(6)));
  }
  ss = ss / (2 * m - 3);
  if (ss < 0.0000000000001) return m;
  ss = 1.0 / ss;
  for (int i = 0; i < 6; ++i) { CD_out[i] = MZ(i + 1) * ss; for (int k = 0; k < 6; ++k) C_out[6 * i + k] = MM(i + 1, k + 1) * ss; }
  return m;
}
