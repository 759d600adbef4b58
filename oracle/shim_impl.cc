// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref3dtk_full.so).
//
// Link-time stand-ins for the two third-party / out-of-scope dependencies of the reference sources that
// oracle/Makefile compiles unmodified:
//  * CXSparse (SuiteSparse, version unpinned by the reference: find_package(SuiteSparse), CMakeLists.txt:82):
//    cs_spalloc / cs_entry / cs_compress / cs_dropzeros / cs_cholsol / cs_qrsol / cs_spfree as called by
//    graphSlam6D::solveSparseCholesky / solveSparseQR (graphSlam6D.cc:317-372,399-419).  Published semantics
//    (T. Davis, "Direct Methods for Sparse Linear Systems", CSparse): triplet assembly sums duplicate entries,
//    cs_cholsol solves the SPD system A x = b in place and returns 1 (0 when A is not positive definite),
//    cs_qrsol solves the least-squares problem.  Here A is kept dense and factorised by a plain Cholesky (LL^T,
//    no fill-reducing permutation) / Householder-free normal equations are NOT used: cs_qrsol does Gaussian
//    elimination with partial pivoting on the square system.  Same solution up to rounding; parity at this
//    boundary is compared on poses, not on factor entries.
//  * scanio: BasicScan's file constructor pulls ScanIO::getScanIO & co.; the harness only builds scans from
//    memory (basicScan.cc:207-252, the ROS node's precedent), so these throw if ever reached.
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "cs.h"
#include "scanio/scan_io.h"
#include "scanio/helper.h"

struct cs_sparse {
  int n = 0;                         // order (max index + 1)
  std::vector<int> ti, tj;           // triplets
  std::vector<double> tx;
  std::vector<double> dense;         // row-major n x n after cs_compress
  bool compressed = false;
};

extern "C" {

cs* cs_spalloc(int, int, int, int, int) { return new cs_sparse(); }

int cs_entry(cs* T, int i, int j, double x) {
  if (!T || i < 0 || j < 0) return 0;
  T->ti.push_back(i); T->tj.push_back(j); T->tx.push_back(x);
  if (i + 1 > T->n) T->n = i + 1;
  if (j + 1 > T->n) T->n = j + 1;
  return 1;
}

cs* cs_compress(const cs* T) {
  cs* A = new cs_sparse();
  A->n = T->n;
  A->dense.assign((size_t)A->n * A->n, 0.0);
  for (size_t k = 0; k < T->tx.size(); ++k) A->dense[(size_t)T->ti[k] * A->n + T->tj[k]] += T->tx[k];
  A->compressed = true;
  return A;
}

int cs_dropzeros(cs*) { return 1; }

int cs_cholsol(int, const cs* A, double* b) {
  const int n = A->n;
  std::vector<double> L(A->dense);
  for (int j = 0; j < n; ++j) {
    double d = L[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
    if (!(d > 0.0)) return 0;
    d = std::sqrt(d);
    L[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = L[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
      L[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) {            // L y = b
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {       // L^T x = y
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  return 1;
}

int cs_qrsol(int, const cs* A, double* b) {
  const int n = A->n;
  std::vector<double> M(A->dense);
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(M[(size_t)r * n + c]) > std::fabs(M[(size_t)piv * n + c])) piv = r;
    if (M[(size_t)piv * n + c] == 0.0) return 0;
    if (piv != c) {
      for (int k = 0; k < n; ++k) std::swap(M[(size_t)c * n + k], M[(size_t)piv * n + k]);
      std::swap(b[c], b[piv]);
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = M[(size_t)r * n + c] / M[(size_t)c * n + c];
      for (int k = c; k < n; ++k) M[(size_t)r * n + k] -= f * M[(size_t)c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= M[(size_t)r * n + k] * b[k];
    b[r] = s / M[(size_t)r * n + r];
  }
  return 1;
}

cs* cs_spfree(cs* A) { delete A; return nullptr; }
int cs_print(const cs*, int) { return 1; }

}  // extern "C"

// ---- scanio stand-ins (never reached by the harness)
ScanIO* ScanIO::getScanIO(IOType) { throw std::runtime_error("oracle shim: scanio is not built (in-memory scans only)"); }
void ScanIO::clearScanIOs() {}

// scanio/helper.cc (file readers) is not built: BasicScan's frames-file pose reader is never reached
void readPoseHelper(const char*, const char*, double*, const char*, const char*) {
  throw std::runtime_error("oracle shim: scanio is not built (in-memory scans only)");
}
