// TEST INFRASTRUCTURE ONLY -- never linked into, imported or called by the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// the library built from this file.
//
// oracle_icp.cpp: plain-C++ CPU restatement of the 3DTK hot path
//   nearest-neighbour correspondence search -> rejection -> pair sums -> 6-DoF solve -> apply,
// written from the reference's behaviour (each function cites the reference lines it follows, paths
// relative to the 3DTK tree @5b570686).  PARITY PINNED: tests/test_oracle_pinning.py checks every
// function here against (a) the reference's own known-answer tests for the k-d tree
// (testing/kdtree/kdtree.cc:20-46, kdtree_indexed_random.cc:192-220) and (b) the reference objects
// themselves, compiled unmodified into oracle/_ref/libref3dtk.so by oracle/Makefile, plus the golden
// vectors under tests/golden/ that were generated from that library (tests/golden/make_golden.py).
//
// Layout notes: 4x4 matrices are column-major double[16]; points are fp64 AoS.

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

inline double sq(double v) { return v * v; }

// ------------------------------------------------------------------ 4x4 helpers (globals.icc)
// MMult, globals.icc:298-321
void mat_mul(const double* a, const double* b, double* out) {
  double r[16];
  for (int col = 0; col < 4; ++col)
    for (int row = 0; row < 4; ++row)
      r[4 * col + row] = a[row] * b[4 * col] + a[row + 4] * b[4 * col + 1] +
                         a[row + 8] * b[4 * col + 2] + a[row + 12] * b[4 * col + 3];
  memcpy(out, r, sizeof r);
}

// M3det, globals.icc:387-397 (row-major 3x3)
double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[6] * m[5]) +
         m[2] * (m[3] * m[7] - m[6] * m[4]);
}

// M4_submat / M4det / M4inv, globals.icc:718-781.  Note the reference indexes Min[si*4+sj].
void minor4(const double* m, double* out, int i, int j) {
  for (int di = 0; di < 3; ++di)
    for (int dj = 0; dj < 3; ++dj) {
      int si = di + (di >= i ? 1 : 0), sj = dj + (dj >= j ? 1 : 0);
      out[3 * di + dj] = m[4 * si + sj];
    }
}

int mat_inv(const double* m, double* out) {
  double sub[9], det = 0.0, sign = 1.0;
  for (int n = 0; n < 4; ++n, sign = -sign) {
    minor4(m, sub, 0, n);
    det += m[n] * det3(sub) * sign;
  }
  if (fabs(det) < 0.00000000000005) {
    for (int k = 0; k < 16; ++k) out[k] = (k % 5 == 0) ? 1.0 : 0.0;
    return 0;
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      int sg = 1 - ((i + j) % 2) * 2;
      minor4(m, sub, i, j);
      out[i + 4 * j] = (det3(sub) * sg) / det;
    }
  return 1;
}

// transform3 (3-argument form), globals.icc:1477-1490
inline void xf_point(const double* M, const double* p, double* q) {
  q[0] = p[0] * M[0] + p[1] * M[4] + p[2] * M[8] + M[12];
  q[1] = p[0] * M[1] + p[1] * M[5] + p[2] * M[9] + M[13];
  q[2] = p[0] * M[2] + p[1] * M[6] + p[2] * M[10] + M[14];
}
// transform3 in place, globals.icc:1454-1463 (adds the translation after the rotation sum)
inline void xf_point_inplace(const double* M, double* p) {
  double x = p[0] * M[0] + p[1] * M[4] + p[2] * M[8];
  double y = p[0] * M[1] + p[1] * M[5] + p[2] * M[9];
  double z = p[0] * M[2] + p[1] * M[6] + p[2] * M[10];
  p[0] = x + M[12]; p[1] = y + M[13]; p[2] = z + M[14];
}
// transform3normal, globals.icc:1465-1475: multiplies by the TRANSPOSE of the rotation block.
inline void xf_normal_inplace(const double* M, double* n) {
  double x = n[0] * M[0] + n[1] * M[1] + n[2] * M[2];
  double y = n[0] * M[4] + n[1] * M[5] + n[2] * M[6];
  double z = n[0] * M[8] + n[1] * M[9] + n[2] * M[10];
  n[0] = x; n[1] = y; n[2] = z;
}

// ------------------------------------------------------------------ k-d tree (kdTreeImpl.h)
// Flat-array restatement of KDTreeImpl::create (kdTreeImpl.h:82-201) and _FindClosest (:345-383):
// centroid split on the longest bbox axis, leaves of <= bucket points or when the largest
// half-extent is < 0.01, Hoare partition with the reference's loop so leaf order -- and therefore the
// winner among exactly equidistant points -- is the same.
struct KdNode {
  double center[3], half[3], splitval;
  int axis;          // -1 => leaf
  int child_lo, child_hi;
  int first, count;  // leaf range into `order`
};

struct KdTree {
  const double* pts = nullptr;
  long n = 0;
  std::vector<int> order;
  std::vector<KdNode> nodes;
  int bucket = 20;

  int build(int lo, int cnt) {
    int me = (int)nodes.size();
    nodes.emplace_back();
    double mn[3], mx[3], cen[3];
    for (int k = 0; k < 3; ++k) mn[k] = mx[k] = cen[k] = pts[3 * (long)order[lo] + k];
    for (int i = 1; i < cnt; ++i)
      for (int k = 0; k < 3; ++k) {
        double v = pts[3 * (long)order[lo + i] + k];
        mn[k] = std::min(mn[k], v);
        mx[k] = std::max(mx[k], v);
        cen[k] += v;
      }
    for (int k = 0; k < 3; ++k) cen[k] /= cnt;
    KdNode nd;
    nd.axis = -1; nd.child_lo = nd.child_hi = -1; nd.first = lo; nd.count = cnt; nd.splitval = 0;
    for (int k = 0; k < 3; ++k) { nd.center[k] = 0.5 * (mn[k] + mx[k]); nd.half[k] = 0.5 * (mx[k] - mn[k]); }
    if (cnt <= bucket) { nodes[me] = nd; return me; }
    const double dx = nd.half[0], dy = nd.half[1], dz = nd.half[2];
    int axis;
    if (dx > dy) axis = (dx > dz) ? 0 : 2;
    else axis = (dy > dz) ? 1 : 2;
    if (fabs(std::max(std::max(dx, dy), dz)) < 0.01) { nodes[me] = nd; return me; }
    nd.axis = axis;
    nd.splitval = cen[axis];
    int* left = &order[lo];
    int* right = &order[lo + cnt - 1];
    while (true) {
      while (pts[3 * (long)(*left) + axis] < nd.splitval) left++;
      while (pts[3 * (long)(*right) + axis] >= nd.splitval) right--;
      if (right < left) break;
      std::swap(*left, *right);
    }
    int nleft = (int)(left - &order[lo]);
    nodes[me] = nd;
    int a = build(lo, nleft);
    int b = build(lo + nleft, cnt - nleft);
    nodes[me].child_lo = a;
    nodes[me].child_hi = b;
    return me;
  }

  void search(int ni, const double* q, double& best_d2, int& best) const {
    const KdNode& nd = nodes[ni];
    if (nd.axis < 0) {
      for (int i = 0; i < nd.count; ++i) {
        const double* p = pts + 3 * (long)order[nd.first + i];
        double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];   // Dist2(query, point)
        double d2 = sq(dx) + sq(dy) + sq(dz);
        if (d2 < best_d2) { best_d2 = d2; best = order[nd.first + i]; }  // strict, kdTreeImpl.h:353
      }
      return;
    }
    double apx = std::max(std::max(fabs(q[0] - nd.center[0]) - nd.half[0],
                                   fabs(q[1] - nd.center[1]) - nd.half[1]),
                          fabs(q[2] - nd.center[2]) - nd.half[2]);
    if (apx >= 0 && sq(apx) >= best_d2) return;
    double myd = nd.splitval - q[nd.axis];
    if (myd >= 0.0) {
      search(nd.child_lo, q, best_d2, best);
      if (sq(myd) < best_d2) search(nd.child_hi, q, best_d2, best);
    } else {
      search(nd.child_hi, q, best_d2, best);
      if (sq(myd) < best_d2) search(nd.child_lo, q, best_d2, best);
    }
  }

  // bounded k-NN (semantics of KDtree::kNearestNeighbors, kd.cc:102-135: the k closest points,
  // nearest first); used by the normals restatement.
  void knn(int ni, const double* q, int k, std::vector<std::pair<double, int> >& heap) const {
    const KdNode& nd = nodes[ni];
    if (nd.axis < 0) {
      for (int i = 0; i < nd.count; ++i) {
        int id = order[nd.first + i];
        const double* p = pts + 3 * (long)id;
        double d2 = sq(p[0] - q[0]) + sq(p[1] - q[1]) + sq(p[2] - q[2]);
        if ((int)heap.size() < k) {
          heap.push_back(std::make_pair(d2, id));
          std::push_heap(heap.begin(), heap.end());
        } else if (d2 < heap.front().first) {
          std::pop_heap(heap.begin(), heap.end());
          heap.back() = std::make_pair(d2, id);
          std::push_heap(heap.begin(), heap.end());
        }
      }
      return;
    }
    double bound = (int)heap.size() < k ? DBL_MAX : heap.front().first;
    double apx = std::max(std::max(fabs(q[0] - nd.center[0]) - nd.half[0],
                                   fabs(q[1] - nd.center[1]) - nd.half[1]),
                          fabs(q[2] - nd.center[2]) - nd.half[2]);
    if (apx >= 0 && sq(apx) >= bound) return;
    double myd = nd.splitval - q[nd.axis];
    int first = myd >= 0.0 ? nd.child_lo : nd.child_hi;
    int second = myd >= 0.0 ? nd.child_hi : nd.child_lo;
    knn(first, q, k, heap);
    bound = (int)heap.size() < k ? DBL_MAX : heap.front().first;
    if (sq(myd) < bound) knn(second, q, k, heap);
  }
};

// ------------------------------------------------------------------ small dense numerics
// Symmetric eigen-decomposition by cyclic Jacobi (the reference calls newmat EigenValues /
// SVD; any convergent method yields the same invariant subspaces).
void jacobi_sym(int n, double* A, double* V) {  // A, V row-major n x n
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = i == j;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0, dg = 0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? dg : off) += sq(A[i * n + j]);
    if (off <= 1e-34 * dg || off == 0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = A[p * n + q];
        if (apq == 0) continue;
        double th = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
        double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; ++k) {
          double x = A[k * n + p], y = A[k * n + q];
          A[k * n + p] = c * x - s * y; A[k * n + q] = s * x + c * y;
        }
        for (int k = 0; k < n; ++k) {
          double x = A[p * n + k], y = A[q * n + k];
          A[p * n + k] = c * x - s * y; A[q * n + k] = s * x + c * y;
        }
        for (int k = 0; k < n; ++k) {
          double x = V[k * n + p], y = V[k * n + q];
          V[k * n + p] = c * x - s * y; V[k * n + q] = s * x + c * y;
        }
      }
  }
}

// SVD of a 3x3 through the eigen-decomposition of H^T H and H H^T is fragile; use one-sided Jacobi.
void svd_3x3(const double* H, double* U, double* w, double* V) {  // row-major
  double A[9];
  memcpy(A, H, sizeof A);
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0);
  for (int sweep = 0; sweep < 100; ++sweep) {
    bool any = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int k = 0; k < 3; ++k) { al += sq(A[3 * k + p]); be += sq(A[3 * k + q]); ga += A[3 * k + p] * A[3 * k + q]; }
        if (ga == 0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
        any = true;
        double ze = (be - al) / (2 * ga);
        double t = (ze >= 0 ? 1.0 : -1.0) / (fabs(ze) + sqrt(1 + ze * ze));
        double c = 1 / sqrt(1 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          double x = A[3 * k + p], y = A[3 * k + q];
          A[3 * k + p] = c * x - s * y; A[3 * k + q] = s * x + c * y;
          x = V[3 * k + p]; y = V[3 * k + q];
          V[3 * k + p] = c * x - s * y; V[3 * k + q] = s * x + c * y;
        }
      }
    if (!any) break;
  }
  int ord[3] = {0, 1, 2};
  double nr[3];
  for (int j = 0; j < 3; ++j) nr[j] = sqrt(sq(A[j]) + sq(A[3 + j]) + sq(A[6 + j]));
  std::sort(ord, ord + 3, [&](int a, int b) { return nr[a] > nr[b]; });
  double Vs[9];
  for (int j = 0; j < 3; ++j) {
    w[j] = nr[ord[j]];
    for (int k = 0; k < 3; ++k) {
      Vs[3 * k + j] = V[3 * k + ord[j]];
      U[3 * k + j] = nr[ord[j]] > 0 ? A[3 * k + ord[j]] / nr[ord[j]] : 0.0;
    }
  }
  memcpy(V, Vs, sizeof Vs);
  // rank-2 input (e.g. coplanar pairs): like a Householder SVD, return a full orthonormal U
  if (w[2] == 0.0 && w[1] > 0.0) {
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
  }
}

// ---- Horn's quartic route (icp6Dquat.cc:171-513) ----------------------------------------------
// lowest real root of x^3 + p x^2 + q x + r (icp6Dquat.cc:305-399; Littlewood's method).  The
// overflow guards of the reference only trigger beyond sqrt(DBL_MAX) and are kept in spirit.
double cubic_root(double p, double q, double r) {
  const double big = sqrt(DBL_MAX);
  if (fabs(p) > big) return -p;
  if (fabs(q) > big) return q > 0 ? -r / q : -sqrt(-q);
  if (fabs(r) > big) return -cbrt(r);
  double p3 = p / 3.0, p3s = p3 * p3;
  if (p3s > big) return -p;
  double v = r + p3 * (p3s + p3s - q);
  if (fabs(v) > big) return -p;
  double u3 = q / 3.0 - p3s;
  double w2 = (2 * u3) * (2 * u3) * u3 + v * v;
  if (w2 >= 0.0) {  // one real root
    double mc = v <= 0.0 ? (-v + sqrt(w2)) * 0.5 : (-v - sqrt(w2)) * 0.5;
    double m = cbrt(mc);
    double n = m != 0.0 ? -u3 / m : 0.0;
    return m + n - p3;
  }
  if (u3 < 0.0) {  // three real roots
    double mu = -u3, s = sqrt(mu), sc = s * mu;
    double t = -v / (sc + sc);
    double ck = cos(acos(t) / 3.0);
    if (p3 < 0.0) return (s + s) * ck - p3;
    double ss = 1.0 - ck * ck;
    if (ss < 0.0) ss = 0.0;
    return s * (-ck - sqrt(3 * ss)) - p3;
  }
  return cbrt(v) - p3;
}

// x^2 + b x + c = 0 (icp6Dquat.cc:274-302)
int quadratic_roots(double b, double c, double* r) {
  double dis = b * b - 4.0 * c;
  if (dis < 0.0) { r[0] = r[1] = 0.0; return 0; }
  double rt = sqrt(dis);
  r[0] = b > 0.0 ? (-b - rt) * 0.5 : (-b + rt) * 0.5;
  r[1] = r[0] == 0.0 ? -b : c / r[0];
  return 2;
}

// x^4 + a x^3 + b x^2 + c x + d = 0, Ferrari-Lagrange (icp6Dquat.cc:171-272)
int quartic_roots(double a, double b, double c, double d, double* rts) {
  double y = cubic_root(b, a * c - 4.0 * d, (a * a - 4.0 * b) * d + c * c);
  double esq = 0.25 * a * a - b - y;
  if (esq < 0.0) return 0;
  double fsq = 0.25 * y * y - d;
  if (fsq < 0.0) return 0;
  double ef = -(0.25 * a * y + 0.5 * c);
  double e, f;
  bool use_ef = ((a > 0.0) && (y > 0.0) && (c > 0.0)) || ((a > 0.0) && (y < 0.0) && (c < 0.0)) ||
                ((a < 0.0) && (y > 0.0) && (c < 0.0)) || ((a < 0.0) && (y < 0.0) && (c > 0.0)) ||
                (a == 0.0) || (y == 0.0) || (c == 0.0);
  if (use_ef && (b < 0.0) && (y < 0.0) && (esq > 0.0)) {
    e = sqrt(esq); f = ef / e;
  } else if (use_ef && (d < 0.0) && (fsq > 0.0)) {
    f = sqrt(fsq); e = ef / f;
  } else {
    e = sqrt(esq); f = sqrt(fsq);
    if (ef < 0.0) f = -f;
  }
  double ah = a * 0.5, g = ah - e, gg = ah + e;
  if (((b > 0.0) && (y > 0.0)) || ((b < 0.0) && (y < 0.0))) {
    if ((a > 0.0) && (e != 0.0)) g = (b + y) / gg;
    else if (e != 0.0) gg = (b + y) / g;
  }
  double h, hh;
  if ((y == 0.0) && (f == 0.0)) { h = hh = 0.0; }
  else if (((f > 0.0) && (y < 0.0)) || ((f < 0.0) && (y > 0.0))) { hh = -0.5 * y + f; h = d / hh; }
  else { h = -0.5 * y - f; hh = d / h; }
  double v1[2], v2[2];
  int n1 = quadratic_roots(gg, hh, v1), n2 = quadratic_roots(g, h, v2);
  rts[0] = v1[0]; rts[1] = v1[1];
  rts[n1] = v2[0]; rts[n1 + 1] = v2[1];
  return n1 + n2;
}

// characteristic polynomial of a symmetric 4x4: l^4 + c0 l^3 + c1 l^2 + c2 l + c3.
// The reference expands the closed form by hand (icp6Dquat.cc:468-513); we use the Faddeev-LeVerrier
// trace recurrence, which gives the same coefficients up to rounding.
void char_poly4(const double Q[4][4], double c[4]) {
  double M[4][4] = {{0}}, T[4][4];
  double coef = 1.0;
  for (int k = 1; k <= 4; ++k) {
    // M_k = Q M_{k-1} + c_{k-1} I ,  c_k = -tr(Q M_k)/k
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = 0;
        for (int l = 0; l < 4; ++l) s += Q[i][l] * M[l][j];
        T[i][j] = s + (i == j ? coef : 0.0);
      }
    memcpy(M, T, sizeof T);
    double tr = 0;
    for (int i = 0; i < 4; ++i)
      for (int l = 0; l < 4; ++l) tr += Q[i][l] * M[l][i];
    coef = -tr / k;
    c[k - 1] = coef;
  }
}

// 4x4 LU with partial pivoting + solves, as used by maxEigenVector (globals.icc:1238-1351)
bool lu4(double A[4][4], int piv[4]) {
  for (int j = 0; j < 4; ++j) {
    int jp = j;
    double t = fabs(A[j][j]);
    for (int i = j + 1; i < 4; ++i)
      if (fabs(A[i][j]) > t) { jp = i; t = fabs(A[i][j]); }
    piv[j] = jp;
    if (A[jp][j] == 0) return false;
    if (jp != j)
      for (int k = 0; k < 4; ++k) std::swap(A[j][k], A[jp][k]);
    double rc = 1.0 / A[j][j];
    for (int k = j + 1; k < 4; ++k) A[k][j] *= rc;
    for (int ii = j + 1; ii < 4; ++ii)
      for (int jj = j + 1; jj < 4; ++jj) A[ii][jj] -= A[ii][j] * A[j][jj];
  }
  return true;
}
void lu4_solve(const double A[4][4], const int piv[4], double b[4]) {
  int ii = 0;
  for (int i = 0; i < 4; ++i) {
    int ip = piv[i];
    double sum = b[ip];
    b[ip] = b[i];
    if (ii) { for (int j = ii; j <= i - 1; ++j) sum -= A[i][j] * b[j]; }
    else if (sum) ii = i;
    b[i] = sum;
  }
  for (int i = 3; i >= 0; --i) {
    double sum = b[i];
    for (int j = i + 1; j < 4; ++j) sum -= A[i][j] * b[j];
    b[i] = sum / A[i][i];
  }
}

// maxEigenVector, icp6Dquat.cc:405-463
void max_eigenvector4(const double Q[4][4], double ev[4]) {
  double c[4], rts[4] = {0, 0, 0, 0};
  char_poly4(Q, c);
  quartic_roots(c[0], c[1], c[2], c[3], rts);
  double l = rts[0];
  for (int i = 1; i < 4; ++i) if (rts[i] > l) l = rts[i];
  double N[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) N[i][j] = Q[i][j] - (i == j ? l : 0.0);
  int piv[4];
  if (!lu4(N, piv)) { ev[0] = 1; ev[1] = ev[2] = ev[3] = 0; return; }
  double best[4] = {1, 0, 0, 0};
  lu4_solve(N, piv, best);
  double len = sq(best[0]) + sq(best[1]) + sq(best[2]) + sq(best[3]);
  for (int i = 1; i < 4; ++i) {
    double cur[4] = {0, 0, 0, 0};
    cur[i] = 1;
    lu4_solve(N, piv, cur);
    double tl = sq(cur[0]) + sq(cur[1]) + sq(cur[2]) + sq(cur[3]);
    if (tl > len) { len = tl; memcpy(best, cur, sizeof cur); }
  }
  len = 1.0 / sqrt(len);
  for (int i = 0; i < 4; ++i) ev[i] = best[i] * len;
}

bool chol_solve(int n, double* A, const double* B, double* x) {  // row-major n x n; globals.icc:820-955
  std::vector<double> diag(n);
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double s = A[i * n + j];
      for (int k = i - 1; k >= 0; --k) s -= A[i * n + k] * A[j * n + k];
      if (i == j) { if (s < 1.0e-7) return false; diag[i] = sqrt(s); }
      else A[j * n + i] = s / diag[i];
    }
  for (int i = 0; i < n; ++i) {
    double s = B[i];
    for (int k = i - 1; k >= 0; --k) s -= A[i * n + k] * x[k];
    x[i] = s / diag[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * x[k];
    x[i] = s / diag[i];
  }
  return true;
}

void sines_to_matrix(const double* x, double* M) {  // icp6Dapx.cc:104-121
  double sx = x[0], cx = sqrt(1.0 - sx * sx), sy = x[1], cy = sqrt(1.0 - sy * sy), sz = x[2],
         cz = sqrt(1.0 - sz * sz);
  M[0] = cy * cz; M[1] = sx * sy * cz + cx * sz; M[2] = -cx * sy * cz + sx * sz; M[3] = 0;
  M[4] = -cy * sz; M[5] = -sx * sy * sz + cx * cz; M[6] = cx * sy * sz + sx * cz; M[7] = 0;
  M[8] = sy; M[9] = -sx * cy; M[10] = cx * cy; M[11] = 0; M[15] = 1;
}

void rot_to_matrix(const double R[3][3], const double* cm, const double* cd, double* M) {
  for (int k = 0; k < 16; ++k) M[k] = (k % 5 == 0) ? 1.0 : 0.0;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) M[4 * c + r] = R[r][c];
  for (int r = 0; r < 3; ++r)
    M[12 + r] = cm[r] - R[r][0] * cd[0] - R[r][1] * cd[1] - R[r][2] * cd[2];
}

// ------------------------------------------------------------------ the four Align functions
double align_quat(long n, const double* p1, const double* p2, const double* cm, const double* cd,
                  double* M) {  // icp6Dquat.cc:38-144
  double S[3][3] = {{0}}, sum = 0;
  for (long i = 0; i < n; ++i) {
    const double* a = p1 + 3 * i;
    const double* b = p2 + 3 * i;
    sum += sq(a[0] - b[0]) + sq(a[1] - b[1]) + sq(a[2] - b[2]);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) S[r][c] += b[r] * a[c];
  }
  double f = 1.0 / double(n);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) S[r][c] = S[r][c] * f - cd[r] * cm[c];
  double tr = S[0][0] + S[1][1] + S[2][2];
  double Q[4][4];
  Q[0][0] = tr;
  Q[0][1] = Q[1][0] = S[1][2] - S[2][1];
  Q[0][2] = Q[2][0] = S[2][0] - S[0][2];
  Q[0][3] = Q[3][0] = S[0][1] - S[1][0];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Q[i + 1][j + 1] = S[i][j] + S[j][i] - (i == j ? tr : 0);
  double q[4];
  max_eigenvector4(Q, q);
  double R[3][3];  // quaternion2matrix, icp6Dquat.cc:148-169
  R[0][0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
  R[1][1] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
  R[2][2] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
  R[0][1] = 2.0 * (q[1] * q[2] - q[0] * q[3]);
  R[1][0] = 2.0 * (q[1] * q[2] + q[0] * q[3]);
  R[0][2] = 2.0 * (q[1] * q[3] + q[0] * q[2]);
  R[2][0] = 2.0 * (q[1] * q[3] - q[0] * q[2]);
  R[1][2] = 2.0 * (q[2] * q[3] - q[0] * q[1]);
  R[2][1] = 2.0 * (q[2] * q[3] + q[0] * q[1]);
  rot_to_matrix(R, cm, cd, M);
  return sqrt(sum / n);
}

double align_svd(long n, const double* p1, const double* p2, const double* cm, const double* cd,
                 double* M) {  // icp6Dsvd.cc:38-158
  double H[9] = {0}, sum = 0;
  for (long i = 0; i < n; ++i) {
    const double* a = p1 + 3 * i;
    const double* b = p2 + 3 * i;
    sum += sq(a[0] - b[0]) + sq(a[1] - b[1]) + sq(a[2] - b[2]);
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) H[3 * j + k] += (b[j] - cd[j]) * (a[k] - cm[k]);
  }
  double U[9], V[9], w[3], R[3][3];
  svd_3x3(H, U, w, V);
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        R[i][j] = V[3 * i] * U[3 * j] + V[3 * i + 1] * U[3 * j + 1] + V[3 * i + 2] * U[3 * j + 2];
    double Rf[9] = {R[0][0], R[0][1], R[0][2], R[1][0], R[1][1], R[1][2], R[2][0], R[2][1], R[2][2]};
    if (det3(Rf) >= 0) break;
    V[2] = -V[2]; V[5] = -V[5]; V[8] = -V[8];   // flip third column, icp6Dsvd.cc:104-109
  }
  rot_to_matrix(R, cm, cd, M);
  return sqrt(sum / (double)n);
}

double align_apx(long n, const double* p1, const double* p2, const double* cm, const double* cd,
                 double* M) {  // icp6Dapx.cc:35-133
  if (n <= 3) {
    for (int k = 0; k < 16; ++k) M[k] = (k % 5 == 0) ? 1.0 : 0.0;
    return 0;
  }
  double A[9] = {0}, B[3] = {0}, sum = 0;
  for (long i = 0; i < n; ++i) {
    const double* a = p1 + 3 * i;
    const double* b = p2 + 3 * i;
    double e[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    double c[3] = {b[0] - cd[0], b[1] - cd[1], b[2] - cd[2]};
    sum += sq(e[0]) + sq(e[1]) + sq(e[2]);
    B[0] += e[2] * c[1] - e[1] * c[2];
    B[1] += e[0] * c[2] - e[2] * c[0];
    B[2] += e[1] * c[0] - e[0] * c[1];
    A[0] += sq(c[1]) + sq(c[2]);
    A[1] -= c[0] * c[1];
    A[2] -= c[0] * c[2];
    A[4] += sq(c[0]) + sq(c[2]);
    A[5] -= c[1] * c[2];
    A[8] += sq(c[0]) + sq(c[1]);
  }
  double x[3];
  if (!chol_solve(3, A, B, x)) return -1.0;
  sines_to_matrix(x, M);
  for (int r = 0; r < 3; ++r) M[12 + r] = cm[r] - M[r] * cd[0] - M[4 + r] * cd[1] - M[8 + r] * cd[2];
  return sqrt(sum / n);
}

double align_napx(long n, const double* p1, const double* p2, const double* nrm, const double* cd,
                  int weighted, double* M) {  // icp6Dnapx.cc:34-149
  double A[36] = {0}, B[6] = {0}, sum = 0;
  for (long i = 0; i < n; ++i) {
    const double* a = p1 + 3 * i;
    const double* b = p2 + 3 * i;
    const double* nn = nrm + 3 * i;
    double d = (a[0] - b[0]) * nn[0] + (a[1] - b[1]) * nn[1] + (a[2] - b[2]) * nn[2];
    double pc[3] = {b[0] - cd[0], b[1] - cd[1], b[2] - cd[2]};
    double v[6] = {pc[1] * nn[2] - pc[2] * nn[1], pc[2] * nn[0] - pc[0] * nn[2],
                   pc[0] * nn[1] - pc[1] * nn[0], nn[0], nn[1], nn[2]};
    sum += d * d;
    for (int r = 0; r < 6; ++r) {
      B[r] += weighted ? d * v[r] : v[r];   // as shipped the residual factor is absent (:69-74)
      for (int c = r; c < 6; ++c) A[6 * r + c] += v[r] * v[c];
    }
  }
  double x[6];
  if (!chol_solve(6, A, B, x)) return -1.0;
  sines_to_matrix(x, M);
  for (int r = 0; r < 3; ++r)
    M[12 + r] = x[3 + r] + cd[r] - M[r] * cd[0] - M[4 + r] * cd[1] - M[8 + r] * cd[2];
  return sqrt(sum / n);
}

}  // namespace

// icp6D_ORTHO::Align (icp6Dortho.cc:41-153): literal per-pair walk in absolute coordinates
double align_ortho(long n, const double* p1, const double* p2, const double* cm, const double* cd, double* M) {
  double sum = 0, H[9] = {0}, HH[9], V[9];
  for (long i = 0; i < n; ++i) {
    double m[3], d[3];
    for (int k = 0; k < 3; ++k) { m[k] = p1[3 * i + k] - cm[k]; d[k] = p2[3 * i + k] - cd[k]; }
    sum += sq(p1[3 * i] - p2[3 * i]) + sq(p1[3 * i + 1] - p2[3 * i + 1]) + sq(p1[3 * i + 2] - p2[3 * i + 2]);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) H[3 * a + b] += m[a] * d[b];
  }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) HH[3 * a + b] = H[a] * H[b] + H[3 + a] * H[3 + b] + H[6 + a] * H[6 + b];
  jacobi_sym(3, HH, V);
  double W[9] = {0};
  for (int k = 0; k < 3; ++k)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) W[3 * a + b] += V[3 * a + k] * V[3 * b + k] / sqrt(HH[4 * k]);
  double R[3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) R[a][b] = H[3 * a] * W[b] + H[3 * a + 1] * W[3 + b] + H[3 * a + 2] * W[6 + b];
  rot_to_matrix(R, cm, cd, M);
  return sqrt(sum / (double)n);
}

// icp6D_DUAL::Align (icp6Ddual.cc:41-150): C1, C2 accumulated pair by pair exactly as written there
double align_dual(long n, const double* p1, const double* p2, double* M) {
  double sum = 0, C1[16] = {0}, C2[16] = {0};
  for (long i = 0; i < n; ++i) {
    const double* m = p1 + 3 * i;
    const double* d = p2 + 3 * i;
    sum += sq(m[0] - d[0]) + sq(m[1] - d[1]) + sq(m[2] - d[2]);
    double Cm[3][3] = {{0, -m[2], m[1]}, {m[2], 0, -m[0]}, {-m[1], m[0], 0}};
    double Cd[3][3] = {{0, -d[2], d[1]}, {d[2], 0, -d[0]}, {-d[1], d[0], 0}};
    C1[0] += m[0] * d[0] + m[1] * d[1] + m[2] * d[2];
    for (int j = 0; j < 3; ++j) {
      double mtCd = m[0] * Cd[0][j] + m[1] * Cd[1][j] + m[2] * Cd[2][j];
      double Cmd = Cm[j][0] * d[0] + Cm[j][1] * d[1] + Cm[j][2] * d[2];
      C1[1 + j] += -mtCd;
      C1[4 * (1 + j)] += -Cmd;
      C2[1 + j] += -d[j] + m[j];
      C2[4 * (1 + j)] += d[j] - m[j];
      for (int k = 0; k < 3; ++k) {
        double CmCd = Cm[j][0] * Cd[0][k] + Cm[j][1] * Cd[1][k] + Cm[j][2] * Cd[2][k];
        C1[4 * (1 + j) + 1 + k] += m[j] * d[k] + CmCd;
        C2[4 * (1 + j) + 1 + k] += -Cd[j][k] - Cm[j][k];
      }
    }
  }
  for (int k = 0; k < 16; ++k) { C1[k] *= -2; C2[k] *= 2; }
  double A[16], V[16];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double t = 0;
      for (int k = 0; k < 4; ++k) t += C2[4 * k + a] * C2[4 * k + b];
      A[4 * a + b] = (t * 1.0 / (2 * n) - C1[4 * a + b] - C1[4 * b + a]) * 0.5;
    }
  jacobi_sym(4, A, V);
  int best = 0;   // SVD of a symmetric matrix: column 1 of U belongs to the largest |eigenvalue|
  for (int k = 1; k < 4; ++k) if (fabs(A[5 * k]) > fabs(A[5 * best])) best = k;
  double qd[4] = {V[best], V[4 + best], V[8 + best], V[12 + best]};
  double q[3] = {qd[1], qd[2], qd[3]};
  double s[4];
  for (int a = 0; a < 4; ++a) {
    double t = 0;
    for (int k = 0; k < 4; ++k) t += C2[4 * a + k] * qd[k];
    s[a] = t * (-1.0) / (2 * n);
  }
  double Cq[3][3] = {{0, -q[2], q[1]}, {q[2], 0, -q[0]}, {-q[1], q[0], 0}};
  double tr[3];
  for (int a = 0; a < 3; ++a) {
    tr[a] = -q[a] * s[0];
    for (int b = 0; b < 3; ++b) tr[a] += ((a == b ? qd[0] : 0.0) + Cq[a][b]) * s[1 + b];
  }
  double qq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  for (int k = 0; k < 16; ++k) M[k] = 0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      M[4 * b + a] = (a == b ? qd[0] * qd[0] - qq : 0.0) + q[a] * q[b] * 2 + Cq[a][b] * qd[0] * 2;
  M[12] = tr[0]; M[13] = tr[1]; M[14] = tr[2]; M[15] = 1;
  return sqrt(sum / (double)n);
}

static bool solve6(double A[6][6], double* b);   // defined with the LUM link below

// icp6D_HELIX::Align + computeRt (icp6Dhelix.cc:48-205): sums over the pairs in absolute coordinates as written
// there; the 6x6 inverse (newmat .i()) by Gaussian elimination with partial pivoting
double align_helix(long n, const double* p1, const double* p2, double* M) {
  double B[6][3] = {{0}}, bd[6] = {0}, sum = 0;
  for (long i = 0; i < n; ++i) {
    double x = p2[3 * i], y = p2[3 * i + 1], z = p2[3 * i + 2];
    B[4][0] += -z; B[3][1] += z; B[5][0] += y; B[3][2] += -y; B[4][2] += x; B[5][1] += -x;
    B[0][0] += z * z + y * y; B[1][0] += y * -x; B[2][0] += -z * x;
    B[1][1] += z * z + x * x; B[2][1] += z * -y; B[2][2] += x * x + y * y;
    double dx = x - p1[3 * i], dy = y - p1[3 * i + 1], dz = z - p1[3 * i + 2];
    bd[0] += -z * dy + y * dz; bd[1] += z * dx - x * dz; bd[2] += -y * dx + x * dy;
    bd[3] += dx; bd[4] += dy; bd[5] += dz;
    sum += dx * dx + dy * dy + dz * dz;
  }
  double A[6][6] = {{0}};
  A[3][3] = A[4][4] = A[5][5] = (double)n;
  A[0][4] = A[4][0] = B[4][0]; A[1][3] = A[3][1] = B[3][1]; A[0][5] = A[5][0] = B[5][0];
  A[2][3] = A[3][2] = B[3][2]; A[2][4] = A[4][2] = B[4][2]; A[1][5] = A[5][1] = B[5][1];
  A[0][1] = A[1][0] = B[1][0]; A[0][2] = A[2][0] = B[2][0]; A[1][2] = A[2][1] = B[2][1];
  A[0][0] = B[0][0]; A[1][1] = B[1][1]; A[2][2] = B[2][2];
  solve6(A, bd);
  double c[3] = {-bd[0], -bd[1], -bd[2]}, cs[3] = {-bd[3], -bd[4], -bd[5]};
  double cl = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  double chk = c[0] * cs[0] + c[1] * cs[1] + c[2] * cs[2];
  double angle = atan(cl);
  double g[3] = {c[0] / cl, c[1] / cl, c[2] / cl};
  double sa = sin(-angle / 2), b0 = cos(-angle / 2), b1 = g[0] * sa, b2 = g[1] * sa, b3 = g[2] * sa;
  double R[3][3];
  R[0][0] = b0 * b0 + b1 * b1 - b2 * b2 - b3 * b3; R[0][1] = 2 * (b1 * b2 + b0 * b3); R[0][2] = 2 * (b1 * b3 - b0 * b2);
  R[1][0] = 2 * (b1 * b2 - b0 * b3); R[1][1] = b0 * b0 - b1 * b1 + b2 * b2 - b3 * b3; R[1][2] = 2 * (b2 * b3 + b0 * b1);
  R[2][0] = 2 * (b1 * b3 + b0 * b2); R[2][1] = 2 * (b2 * b3 - b0 * b1); R[2][2] = b0 * b0 - b1 * b1 - b2 * b2 + b3 * b3;
  double nn = b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) R[a][b] /= nn;
  double skew = chk / (cl * cl);
  double gs[3] = {(cs[0] - c[0] * skew) / cl, (cs[1] - c[1] * skew) / cl, (cs[2] - c[2] * skew) / cl};
  double pt[3] = {g[1] * gs[2] - g[2] * gs[1], g[2] * gs[0] - g[0] * gs[2], g[0] * gs[1] - g[1] * gs[0]};
  for (int k = 0; k < 16; ++k) M[k] = 0;
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) M[4 * b + a] = R[a][b];
  for (int a = 0; a < 3; ++a)
    M[12 + a] = -(R[a][0] * pt[0] + R[a][1] * pt[1] + R[a][2] * pt[2]) + g[a] * (skew * angle) + pt[a];
  M[15] = 1;
  return sqrt(sum / (double)n);
}

extern "C" {

// ---- math pins
int orc_m4inv(const double* in, double* out) { return mat_inv(in, out); }
void orc_mmult(const double* a, const double* b, double* out) { mat_mul(a, b, out); }
void orc_euler_to_matrix4(const double* pos, const double* th, double* M) {  // globals.icc:501-531
  double sx = sin(th[0]), cx = cos(th[0]), sy = sin(th[1]), cy = cos(th[1]), sz = sin(th[2]),
         cz = cos(th[2]);
  M[0] = cy * cz; M[1] = sx * sy * cz + cx * sz; M[2] = -cx * sy * cz + sx * sz; M[3] = 0.0;
  M[4] = -cy * sz; M[5] = -sx * sy * sz + cx * cz; M[6] = cx * sy * sz + sx * cz; M[7] = 0.0;
  M[8] = sy; M[9] = -sx * cy; M[10] = cx * cy; M[11] = 0.0;
  M[12] = pos[0]; M[13] = pos[1]; M[14] = pos[2]; M[15] = 1;
}

// ---- search structure
void* orc_tree_create(const double* xyz, long n, int bucket) {
  if (n <= 0) return nullptr;   // kdTreeImpl.h:86-88 throws on zero points
  KdTree* t = new KdTree();
  t->pts = xyz;   // like the reference, the tree refers to the caller's array
  t->n = n;
  t->bucket = bucket > 0 ? bucket : 20;
  t->order.resize(n);
  for (long i = 0; i < n; ++i) t->order[i] = (int)i;
  t->nodes.reserve(n / 4 + 16);
  t->build(0, (int)n);
  return t;
}
void orc_tree_free(void* h) { delete (KdTree*)h; }

// KDtree::FindClosest, kd.cc:78-87
long orc_find_closest(void* h, const double* q, double maxdist2) {
  KdTree* t = (KdTree*)h;
  double best = maxdist2;
  int id = -1;
  t->search(0, q, best, id);
  return id;
}

void orc_find_closest_batch(void* h, const double* q, long nq, double maxdist2, int* idx,
                            double* d2) {
  KdTree* t = (KdTree*)h;
#pragma omp parallel for schedule(dynamic, 1024)
  for (long i = 0; i < nq; ++i) {
    double best = maxdist2;
    int id = -1;
    t->search(0, q + 3 * i, best, id);
    idx[i] = id;
    if (d2) d2[i] = id >= 0 ? best : -1.0;
  }
}

// brute force with the acceptance rule of the k-d tree (strict <) -- the differential oracle of
// testing/kdtree/kdtree_indexed_random.cc:14-26
long orc_brute_closest(const double* xyz, long n, const double* q, double maxdist2) {
  long best = -1;
  for (long i = 0; i < n; ++i) {
    const double* p = xyz + 3 * i;
    double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    double d2 = sq(dx) + sq(dy) + sq(dz);
    if (d2 < maxdist2) { maxdist2 = d2; best = i; }
  }
  return best;
}

// ---- SearchTree::getPtPairs, searchTree.cc:92-188 (rnd <= 1 only: the reference's sampling uses
// the global std::rand stream and is not reproducible)
long orc_get_pt_pairs(void* h, const double* source_alignxf, const double* data_xyz,
                      const double* data_nrm, long start, long end, double maxdist2,
                      int pairing_mode, double* p1, double* p2, double* nrm, int* idx, double* sum,
                      double* cm, double* cd) {
  KdTree* t = (KdTree*)h;
  double inv[16];
  mat_inv(source_alignxf, inv);
  long np = 0;
  for (long i = start; i < end; ++i) {
    double tq[3] = {data_xyz[3 * i], data_xyz[3 * i + 1], data_xyz[3 * i + 2]}, s[3], nn[3] = {0, 0, 0};
    xf_point(inv, tq, s);
    if (pairing_mode != 0) {
      nn[0] = data_nrm[3 * i]; nn[1] = data_nrm[3 * i + 1]; nn[2] = data_nrm[3 * i + 2];
      double l = sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);   // Normalize3
      nn[0] /= l; nn[1] /= l; nn[2] /= l;
    }
    double best = maxdist2;
    int id = -1;
    t->search(0, s, best, id);
    if (id < 0) continue;
    xf_point(source_alignxf, t->pts + 3 * (long)id, s);
    if (pairing_mode == 2) {   // CLOSEST_PLANE_SIMPLE: s <- (n.(s-t)) n + t, searchTree.cc:149-162
      double tmp[3] = {s[0] - tq[0], s[1] - tq[1], s[2] - tq[2]};
      double dot = nn[0] * tmp[0] + nn[1] * tmp[1] + nn[2] * tmp[2];
      for (int k = 0; k < 3; ++k) s[k] = nn[k] * dot + tq[k];
    }
    for (int k = 0; k < 3; ++k) { cm[k] += s[k]; cd[k] += tq[k]; }
    *sum += sq(s[0] - tq[0]) + sq(s[1] - tq[1]) + sq(s[2] - tq[2]);
    for (int k = 0; k < 3; ++k) { p1[3 * np + k] = s[k]; p2[3 * np + k] = tq[k]; if (nrm) nrm[3 * np + k] = nn[k]; }
    if (idx) idx[np] = id;
    ++np;
  }
  return np;
}

double orc_align(int algo, long n, const double* p1, const double* p2, const double* nrm,
                 const double* cm, const double* cd, int napx_weighted, double* alignxf) {
  switch (algo) {
    case 1: return align_quat(n, p1, p2, cm, cd, alignxf);
    case 2: return align_svd(n, p1, p2, cm, cd, alignxf);
    case 3: return align_ortho(n, p1, p2, cm, cd, alignxf);
    case 4: return align_dual(n, p1, p2, alignxf);
    case 5: return align_helix(n, p1, p2, alignxf);
    case 6: return align_apx(n, p1, p2, cm, cd, alignxf);
    case 10: return align_napx(n, p1, p2, nrm, cd, napx_weighted, alignxf);
  }
  return -2.0;
}

// ---- icp6D::match, serial arm (icp6D.cc:104-285 with Scan::getPtPairs scan.cc:1220-1260 and
// Scan::transform scan.cc:851-898).  data_xyz / data_nrm are moved in place.
int orc_match(void* model_tree, const double* model_dalignxf, double* data_xyz, double* data_nrm,
              long nd, double* data_transmat, double* data_dalignxf, int algo, int pairing_mode,
              double max_dist_match, int max_iter, double eps, int napx_weighted, double* rms_out,
              long* npairs_out, int* iters_done) {
  *iters_done = 0;
  if (max_iter == 0) return 0;
  const double md2 = sq(max_dist_match);
  std::vector<double> p1(3 * nd), p2(3 * nd), pn(3 * nd);
  double ret = 0, prev = 0, pprev = 0, alignxf[16];
  for (int k = 0; k < 16; ++k) alignxf[k] = (k % 5 == 0) ? 1.0 : 0.0;
  int iter = 0;
  for (iter = 0; iter < max_iter; ++iter) {
    pprev = prev;
    prev = ret;
    double cm[3] = {0, 0, 0}, cd[3] = {0, 0, 0};
    long np = orc_get_pt_pairs(model_tree, model_dalignxf, data_xyz, data_nrm, 0, nd, md2,
                               pairing_mode, p1.data(), p2.data(), pn.data(), nullptr, &ret, cm, cd);
    if (np != 0)
      for (int k = 0; k < 3; ++k) { cm[k] /= np; cd[k] /= np; }
    if (np > 3) ret = orc_align(algo, np, p1.data(), p2.data(), pn.data(), cm, cd, napx_weighted, alignxf);
    else break;
    rms_out[*iters_done] = ret;
    npairs_out[*iters_done] = np;
    ++*iters_done;
    for (long i = 0; i < nd; ++i) xf_point_inplace(alignxf, data_xyz + 3 * i);
    if (data_nrm)
      for (long i = 0; i < nd; ++i) xf_normal_inplace(alignxf, data_nrm + 3 * i);
    mat_mul(alignxf, data_transmat, data_transmat);
    mat_mul(alignxf, data_dalignxf, data_dalignxf);
    if ((fabs(ret - prev) < eps && fabs(ret - pprev) < eps) || iter == max_iter - 1) break;
  }
  return iter;
}

// ---- normals: calculateNormalsKNN + calculateNormal (normals.cc:220-295, :518-558)
void orc_normals_knn(const double* xyz, long n, int k, const double* rPos, double* out) {
  KdTree* t = (KdTree*)orc_tree_create(xyz, n, 20);
#pragma omp parallel for schedule(dynamic, 256)
  for (long i = 0; i < n; ++i) {
    const double* p = xyz + 3 * i;
    std::vector<std::pair<double, int> > heap;
    heap.reserve(k + 1);
    t->knn(0, p, k, heap);
    int m = (int)heap.size();
    double mean[3] = {0, 0, 0};
    for (int j = 0; j < m; ++j)
      for (int c = 0; c < 3; ++c) mean[c] += xyz[3 * (long)heap[j].second + c];
    for (int c = 0; c < 3; ++c) mean[c] /= m;
    double C[9] = {0}, V[9];
    for (int j = 0; j < m; ++j) {
      double d[3];
      for (int c = 0; c < 3; ++c) d[c] = xyz[3 * (long)heap[j].second + c] - mean[c];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) C[3 * a + b] += d[a] * d[b];
    }
    for (int a = 0; a < 9; ++a) C[a] *= 1.0 / m;
    jacobi_sym(3, C, V);
    int lo = 0;
    for (int a = 1; a < 3; ++a) if (C[4 * a] < C[4 * lo]) lo = a;
    double nv[3] = {V[lo], V[3 + lo], V[6 + lo]};
    double pv[3] = {p[0] - rPos[0], p[1] - rPos[1], p[2] - rPos[2]};
    double pl = sqrt(sq(pv[0]) + sq(pv[1]) + sq(pv[2]));
    double ang = (nv[0] * pv[0] + nv[1] * pv[1] + nv[2] * pv[2]) / pl;
    if (ang < 0) { nv[0] = -nv[0]; nv[1] = -nv[1]; nv[2] = -nv[2]; }
    double nl = sqrt(sq(nv[0]) + sq(nv[1]) + sq(nv[2]));
    for (int c = 0; c < 3; ++c) out[3 * i + c] = nv[c] / nl;
  }
  orc_tree_free(t);
}

}  // extern "C"

// ---- lum6DEuler::covarianceEuler (lum6Deuler.cc:94-260): per-link 6x6 information matrix C = MM/s^2 and
// vector CD = MZ/s^2 from the pairs of (first = model tree, second = data).  The 6x6 inverse the
// reference takes with newmat (MM.i(), an LU factorisation) is done here by Gaussian elimination with
// partial pivoting.  PARITY NOTE: lum6Deuler.cc itself cannot be compiled here (its header chain needs
// CXSparse), so this function is pinned against the harness' restatement that uses the reference's
// own newmat for the inverse, not against covarianceEuler directly.
static bool solve6(double A[6][6], double* b) {
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (A[piv][c] == 0.0) return false;
    if (piv != c) { for (int k = 0; k < 6; ++k) std::swap(A[c][k], A[piv][k]); std::swap(b[c], b[piv]); }
    for (int r = c + 1; r < 6; ++r) {
      double f = A[r][c] / A[c][c];
      for (int k = c; k < 6; ++k) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  for (int r = 5; r >= 0; --r) {
    double t = b[r];
    for (int k = r + 1; k < 6; ++k) t -= A[r][k] * b[k];
    b[r] = t / A[r][r];
  }
  return true;
}

extern "C" long orc_lum_link(void* model_tree, const double* model_dalignxf, const double* data_xyz, long nd,
                             double maxdist2, double* C, double* CD) {
  std::vector<double> p1(3 * nd), p2(3 * nd);
  double dummy_sum = 0, cm[3] = {0, 0, 0}, cd[3] = {0, 0, 0};
  long m = orc_get_pt_pairs(model_tree, model_dalignxf, data_xyz, nullptr, 0, nd, maxdist2, 0, p1.data(),
                            p2.data(), nullptr, nullptr, &dummy_sum, cm, cd);
  for (int i = 0; i < 36; ++i) C[i] = 0.0;
  for (int i = 0; i < 6; ++i) CD[i] = 0.0;
  if (m <= 2) return m;
  double sx = 0, sy = 0, sz = 0, xy = 0, yz = 0, xz = 0, ypz = 0, xpz = 0, xpy = 0, MZ[6] = {0, 0, 0, 0, 0, 0};
  for (long j = 0; j < m; ++j) {
    const double* a = &p1[3 * j];
    const double* b = &p2[3 * j];
    double x = (a[0] + b[0]) / 2.0, y = (a[1] + b[1]) / 2.0, z = (a[2] + b[2]) / 2.0;
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    sx += x; sy += y; sz += z;
    xpy += x * x + y * y; xpz += x * x + z * z; ypz += y * y + z * z;
    xy += x * y; xz += x * z; yz += y * z;
    MZ[0] += dx; MZ[1] += dy; MZ[2] += dz;
    MZ[3] += -z * dy + y * dz; MZ[4] += -y * dx + x * dy; MZ[5] += z * dx - x * dz;
  }
  double MM[6][6] = {{0}};
  MM[0][0] = MM[1][1] = MM[2][2] = (double)m;
  MM[3][3] = ypz; MM[4][4] = xpy; MM[5][5] = xpz;
  MM[0][4] = MM[4][0] = -sy; MM[0][5] = MM[5][0] = sz;
  MM[1][3] = MM[3][1] = -sz; MM[1][4] = MM[4][1] = sx;
  MM[2][3] = MM[3][2] = sy;  MM[2][5] = MM[5][2] = -sx;
  MM[3][4] = MM[4][3] = -xz; MM[3][5] = MM[5][3] = -xy; MM[4][5] = MM[5][4] = -yz;
  double A[6][6], D[6];
  memcpy(A, MM, sizeof A);
  memcpy(D, MZ, sizeof D);
  if (!solve6(A, D)) return m;
  double ss = 0;
  for (long j = 0; j < m; ++j) {
    const double* a = &p1[3 * j];
    const double* b = &p2[3 * j];
    double x = (a[0] + b[0]) / 2.0, y = (a[1] + b[1]) / 2.0, z = (a[2] + b[2]) / 2.0;
    ss += sq(a[0] - b[0] - (D[0] - y * D[4] + z * D[5])) + sq(a[1] - b[1] - (D[1] - z * D[3] + x * D[4])) +
          sq(a[2] - b[2] - (D[2] + y * D[3] - x * D[5]));
  }
  ss = ss / (2 * m - 3);
  if (ss < 0.0000000000001) return m;   // identical clouds: C = CD = 0 (lum6Deuler.cc:219-231)
  ss = 1.0 / ss;
  for (int i = 0; i < 6; ++i) { CD[i] = MZ[i] * ss; for (int k = 0; k < 6; ++k) C[6 * i + k] = MM[i][k] * ss; }
  return m;
}

// ---- Graph / doGraphSlam6D (TEST INFRASTRUCTURE, like everything in this file) ------------------------------
// Graph::Graph(int nodes, double cldist2, int loopsize), graph.cc:108-127
extern "C" int orc_graph_from_poses(const double* rpos, int n, double cldist2, int loopsize, int* links, int cap) {
  int m = 0;
  for (int i = 0; i + 1 < n; ++i) { if (m < cap) { links[2 * m] = i; links[2 * m + 1] = i + 1; } ++m; }
  for (int j = 0; j < n; ++j)
    for (int k = j + 1; k < n; ++k) {
      double d2 = sq(rpos[3 * j] - rpos[3 * k]) + sq(rpos[3 * j + 1] - rpos[3 * k + 1]) + sq(rpos[3 * j + 2] - rpos[3 * k + 2]);
      if (abs(k - j) > loopsize && d2 < cldist2) { if (m < cap) { links[2 * m] = j; links[2 * m + 1] = k; } ++m; }
    }
  return m;
}

// Matrix4ToEuler, globals.icc:540-578
extern "C" void orc_matrix4_to_euler(const double* a, double* th, double* pos) {
  th[1] = a[0] > 0.0 ? asin(a[8]) : M_PI - asin(a[8]);
  double C = cos(th[1]);
  if (fabs(C) > 0.005) { th[0] = atan2(-a[9] / C, a[10] / C); th[2] = atan2(-a[4] / C, a[0] / C); }
  else { th[0] = 0.0; th[2] = atan2(a[1], a[5]); }
  if (pos) { pos[0] = a[12]; pos[1] = a[13]; pos[2] = a[14]; }
}

// general dense solve by LU with partial pivoting (deliberately NOT the Cholesky route the product takes:
// G is SPD, both must give the same X up to rounding)
static bool lu_solve(int n, std::vector<double>& A, std::vector<double>& b) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (fabs(A[(size_t)r * n + c]) > fabs(A[(size_t)piv * n + c])) piv = r;
    if (A[(size_t)piv * n + c] == 0.0) return false;
    if (piv != c) { for (int k = 0; k < n; ++k) std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]); std::swap(b[c], b[piv]); }
    for (int r = c + 1; r < n; ++r) {
      double f = A[(size_t)r * n + c] / A[(size_t)c * n + c];
      if (f == 0.0) continue;
      for (int k = c; k < n; ++k) A[(size_t)r * n + k] -= f * A[(size_t)c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double t = b[r];
    for (int k = r + 1; k < n; ++k) t -= A[(size_t)r * n + k] * b[k];
    b[r] = t / A[(size_t)r * n + r];
  }
  return true;
}

// lum6DEuler::doGraphSlam6D (lum6Deuler.cc:314-479) with FillGB3D (:265-304) and Scan::transformToEuler
// (scan.cc:1061-1083) over in-memory scans.  xyz_all: the scans' "xyz reduced original" rows back to back
// (offsets[n+1]); transmats / dalignxfs: 16 doubles per scan, in/out.  Like the reference, the "xyz reduced"
// working copy of every scan is MOVED by each transform (the product never moves points).
// PARITY NOTE: lum6Deuler.cc does not compile here (CXSparse); the linear solve is the reference's SPD system
// solved densely.  G_out/B_out (optional) return the system of the first iteration.
extern "C" int orc_lum_graph_slam(int n_scans, const double* xyz_all, const long* offsets, const int* links,
                                  int n_links, double maxdist2, int nr_it, double eps, double* transmats,
                                  double* dalignxfs, double* ret_out, double* G_out, double* B_out) {
  if (n_scans <= 0) return -1;
  std::vector<void*> trees(n_scans);
  std::vector<std::vector<double>> cur(n_scans);
  for (int i = 0; i < n_scans; ++i) {
    long n = offsets[i + 1] - offsets[i];
    trees[i] = orc_tree_create(xyz_all + 3 * offsets[i], n, 20);
    cur[i].assign(xyz_all + 3 * offsets[i], xyz_all + 3 * offsets[i + 1]);
    for (long k = 0; k < n; ++k) xf_point_inplace(dalignxfs + 16 * i, &cur[i][3 * k]);
  }
  double ret = DBL_MAX;
  int it = 0;
  const int dim = 6 * (n_scans - 1);
  for (; it < nr_it && ret > eps && dim > 0; ++it) {
    std::vector<double> G((size_t)dim * dim, 0.0), B(dim, 0.0);
    for (int l = 0; l < n_links; ++l) {
      int first = links[2 * l], second = links[2 * l + 1], a = first - 1, b = second - 1;
      double C[36], CD[6];
      orc_lum_link(trees[first], dalignxfs + 16 * first, cur[second].data(), (long)(cur[second].size() / 3),
                   maxdist2, C, CD);
      for (int r = 0; r < 6; ++r) {
        if (a >= 0) B[6 * a + r] += CD[r];
        if (b >= 0) B[6 * b + r] -= CD[r];
        for (int c = 0; c < 6; ++c) {
          if (a >= 0) G[(size_t)(6 * a + r) * dim + 6 * a + c] += C[6 * r + c];
          if (b >= 0) G[(size_t)(6 * b + r) * dim + 6 * b + c] += C[6 * r + c];
          if (a >= 0 && b >= 0) {
            G[(size_t)(6 * a + r) * dim + 6 * b + c] -= C[6 * r + c];
            G[(size_t)(6 * b + r) * dim + 6 * a + c] -= C[6 * r + c];
          }
        }
      }
    }
    if (it == 0 && G_out) memcpy(G_out, G.data(), G.size() * sizeof(double));
    if (it == 0 && B_out) memcpy(B_out, B.data(), B.size() * sizeof(double));
    std::vector<double> X = B, A = G;
    if (!lu_solve(dim, A, X)) return -2;
    double sum_position_diff = 0.0;
    for (int i = 1; i < n_scans; ++i) {
      double* T = transmats + 16 * i;
      double* dal = dalignxfs + 16 * i;
      double th[3], pos[3];
      orc_matrix4_to_euler(T, th, pos);
      double xa = pos[0], ya = pos[1], za = pos[2];
      double ctx = cos(th[0]), stx = sin(th[0]), cty = cos(th[1]), sty = sin(th[1]);
      double Ha[6][6] = {{0}};
      for (int k = 0; k < 6; ++k) Ha[k][k] = 1.0;
      Ha[0][4] = -za * ctx + ya * stx;       Ha[0][5] = ya * cty * ctx + za * stx * cty;
      Ha[1][3] = za;  Ha[1][4] = -xa * stx;  Ha[1][5] = -xa * ctx * cty + za * sty;
      Ha[2][3] = -ya; Ha[2][4] = xa * ctx;   Ha[2][5] = -xa * cty * stx - ya * sty;
      Ha[3][5] = sty;
      Ha[4][4] = stx; Ha[4][5] = ctx * cty;
      Ha[5][4] = ctx; Ha[5][5] = -stx * cty;
      double result[6];
      for (int k = 0; k < 6; ++k) result[k] = X[6 * (i - 1) + k];
      if (!solve6(Ha, result)) return -3;        // result = Ha^-1 * Xtmp
      double npos[3], nth[3];
      for (int k = 0; k < 3; ++k) { npos[k] = pos[k] - result[k]; nth[k] = th[k] - result[k + 3]; }
      double tinv[16], alignxf[16], tmp[16];
      mat_inv(T, tinv);
      orc_euler_to_matrix4(npos, nth, alignxf);
      const double* steps[2] = {tinv, alignxf};
      for (int s = 0; s < 2; ++s) {            // Scan::transform twice: points, transMat, dalignxf
        for (size_t k = 0; k < cur[i].size() / 3; ++k) xf_point_inplace(steps[s], &cur[i][3 * k]);
        mat_mul(steps[s], T, tmp); memcpy(T, tmp, sizeof tmp);
        mat_mul(steps[s], dal, tmp); memcpy(dal, tmp, sizeof tmp);
      }
      sum_position_diff += sqrt(sq(result[0]) + sq(result[1]) + sq(result[2]));
    }
    ret = sum_position_diff / (double)n_scans;
  }
  for (int i = 0; i < n_scans; ++i) orc_tree_free(trees[i]);
  if (ret_out) *ret_out = ret;
  return it;
}
