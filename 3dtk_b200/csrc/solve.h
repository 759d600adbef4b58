// solve.h -- the O(1) 6-DoF solve of one ICP iteration, from pair MOMENTS instead of a pair list.
//
// The reference minimizers walk a std::vector<PtPair> (208 B / pair):
//   icp6D_QUAT::Align  src/slam6d/icp6Dquat.cc:38-144    (Horn, unit quaternion)
//   icp6D_SVD::Align   src/slam6d/icp6Dsvd.cc:38-158     (Arun, SVD of the cross-covariance)
//   icp6D_APX::Align   src/slam6d/icp6Dapx.cc:35-133     (small-angle, 3x3 Cholesky)
//   icp6D_NAPX::Align  src/slam6d/icp6Dnapx.cc:34-149    (point-to-plane small-angle, 6x6 Cholesky)
// Here the correspondence kernel reduces every accepted pair into a fixed set of fp64 sums taken
// about a shift origin `o` (so products stay small), and these functions turn the sums into the
// same alignxf / RMS.  All four results are translation-covariant, which is what makes the shift
// legal: centred second moments do not depend on o, and t is rebuilt from the un-shifted centroids.
//
// Compiled for host (b200icp_align_pairs, CPU tests) and device (solve kernel): no libc++ types.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b200 {

// ---- moment layouts ---------------------------------------------------------------------------
// Point-to-point family (QUAT, SVD, APX).  p1 = model-side point, p2 = data-side point, primes are
// coordinates minus the shift origin.
enum : int {
  MP_N = 0,      // number of pairs
  MP_D2 = 1,     // sum |p1 - p2|^2
  MP_M = 2,      // [3] sum p1'
  MP_D = 5,      // [3] sum p2'
  MP_DM = 8,     // [9] sum p2'_i * p1'_j, row-major i*3+j   (S of icp6Dquat.cc:63-71)
  MP_DD = 17,    // [6] sum p2' p2'^T : xx xy xz yy yz zz
  NS_P2P = 23
};
// Point-to-plane (NAPX).  n = unit normal carried by the pair, a = p2' x n, d = (p1 - p2) . n
enum : int {
  MN_N = 0,
  MN_D2 = 1,     // sum d^2
  MN_M = 2,      // [3] sum p1'
  MN_D = 5,      // [3] sum p2'
  MN_A = 8,      // [3] sum a
  MN_NRM = 11,   // [3] sum n
  MN_AA = 14,    // [6] sum a a^T  (xx xy xz yy yz zz)
  MN_AN = 20,    // [9] sum a_i n_j
  MN_NN = 29,    // [6] sum n n^T
  MN_DA = 35,    // [3] sum d * a      (only used by the least-squares variant)
  MN_DN = 38,    // [3] sum d * n
  MN_PD2 = 41,   // sum |p1 - p2|^2 (what getPtPairs reports as `sum`)
  NS_NAPX = 44
};
constexpr int NS_MAX = 44;

B2_HD int sym6(int i, int j) {  // index into a packed symmetric 3x3 (xx xy xz yy yz zz)
  if (i > j) { int t = i; i = j; j = t; }
  return i == 0 ? j : (i == 1 ? 2 + j : 5);
}

// ---- 4x4 column-major helpers (globals.icc:282-321, :761-781 semantics) ------------------------
B2_HD void m4_identity(double* M) {
  for (int i = 0; i < 16; ++i) M[i] = 0.0;
  M[0] = M[5] = M[10] = M[15] = 1.0;
}

B2_HD void m4_mul(const double* A, const double* B, double* C) {  // C = A*B, C may not alias
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      // same operand order as MMult: A[r]*B[4c] + A[r+4]*B[4c+1] + A[r+8]*B[4c+2] + A[r+12]*B[4c+3]
      C[4 * c + r] = A[r] * B[4 * c] + A[r + 4] * B[4 * c + 1] + A[r + 8] * B[4 * c + 2] +
                     A[r + 12] * B[4 * c + 3];
    }
}

B2_HD double det3(double a, double b, double c, double d, double e, double f, double g, double h,
                  double i) {
  return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// General 4x4 inverse by cofactors; returns 0 (and identity) when |det| < 5e-14 like M4inv.
B2_HD int m4_inverse(const double* M, double* out) {
  double cof[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s[9];
      int k = 0;
      for (int rr = 0; rr < 4; ++rr) {
        if (rr == r) continue;
        for (int cc = 0; cc < 4; ++cc) {
          if (cc == c) continue;
          s[k++] = M[4 * cc + rr];
        }
      }
      double d = det3(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8]);
      cof[4 * c + r] = ((r + c) & 1) ? -d : d;   // cofactor of element (r,c)
    }
  double det = M[0] * cof[0] + M[4] * cof[4] + M[8] * cof[8] + M[12] * cof[12];  // expand row 0
  if (fabs(det) < 0.00000000000005) { m4_identity(out); return 0; }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[4 * c + r] = cof[4 * r + c] / det;  // inverse = adj^T / det
  return 1;
}

B2_HD void xf_point(const double* M, const double* p, double* q) {  // transform3, 3-arg form
  q[0] = p[0] * M[0] + p[1] * M[4] + p[2] * M[8] + M[12];
  q[1] = p[0] * M[1] + p[1] * M[5] + p[2] * M[9] + M[13];
  q[2] = p[0] * M[2] + p[1] * M[6] + p[2] * M[10] + M[14];
}

// ---- small dense kernels ------------------------------------------------------------------------
// Cyclic Jacobi eigen-decomposition of a symmetric NxN (N = 3 or 4).  A is destroyed (diagonal ->
// eigenvalues), V columns -> eigenvectors.
template <int N>
B2_HD void jacobi_eig(double A[N][N], double V[N][N]) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        if (i == j) diag += A[i][i] * A[i][i];
        else off += A[i][j] * A[i][j];
      }
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        double apq = A[p][q];
        if (apq == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {  // A <- A J
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {  // A <- J^T A
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

// One-sided (Hestenes) Jacobi SVD of a 3x3: H = U diag(w) V^T, w sorted descending.
B2_HD void svd3(const double H[3][3], double U[3][3], double w[3], double V[3][3]) {
  double A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { A[i][j] = H[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) {
          alpha += A[k][p] * A[k][p];
          beta += A[k][q] * A[k][q];
          gamma += A[k][p] * A[k][q];
        }
        if (gamma == 0.0 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          double x = A[k][p], y = A[k][q];
          A[k][p] = c * x - s * y;
          A[k][q] = s * x + c * y;
          x = V[k][p]; y = V[k][q];
          V[k][p] = c * x - s * y;
          V[k][q] = s * x + c * y;
        }
      }
    if (!rotated) break;
  }
  int ord[3] = {0, 1, 2};
  double nrm[3];
  for (int j = 0; j < 3; ++j) nrm[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (nrm[ord[j]] > nrm[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
  double Vs[3][3];
  for (int j = 0; j < 3; ++j) {
    int s = ord[j];
    w[j] = nrm[s];
    for (int k = 0; k < 3; ++k) {
      Vs[k][j] = V[k][s];
      U[k][j] = nrm[s] > 0.0 ? A[k][s] / nrm[s] : 0.0;
    }
  }
  // rank-deficient input: complete U to an orthonormal basis so that V U^T is still orthogonal
  if (w[2] <= 1e-300 * (w[0] > 0 ? w[0] : 1.0) || w[2] == 0.0) {
    if (w[1] > 0.0) {
      U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
      U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
      U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = Vs[i][j];
}

template <int N>
B2_HD bool cholesky_solve(double A[N][N], const double* B, double* x) {
  // Numerical-Recipes style choldc/cholsl with the reference's 1e-7 pivot floor
  // (globals.icc:820-868, :907-955).  Upper triangle of A is read, lower overwritten.
  double diag[N];
  for (int i = 0; i < N; ++i)
    for (int j = i; j < N; ++j) {
      double sum = A[i][j];
      for (int k = i - 1; k >= 0; --k) sum -= A[i][k] * A[j][k];
      if (i == j) {
        if (sum < 1.0e-7) return false;
        diag[i] = sqrt(sum);
      } else {
        A[j][i] = sum / diag[i];
      }
    }
  for (int i = 0; i < N; ++i) {
    double sum = B[i];
    for (int k = i - 1; k >= 0; --k) sum -= A[i][k] * x[k];
    x[i] = sum / diag[i];
  }
  for (int i = N - 1; i >= 0; --i) {
    double sum = x[i];
    for (int k = i + 1; k < N; ++k) sum -= A[k][i] * x[k];
    x[i] = sum / diag[i];
  }
  return true;
}

// Eigenvector of the LARGEST eigenvalue of a symmetric 4x4 (Horn's N matrix, icp6Dquat.cc:405-463).
// The reference takes lambda_max from the characteristic quartic (Ferrari) and the vector from an LU solve of
// (Q - lambda I); here: quartic coefficients by the Faddeev-LeVerrier trace recurrence, lambda_max by Newton
// from an upper bound (monotone for a polynomial with only real roots), vector = largest column of the
// adjugate of (Q - lambda I).  A handful of dependent divisions instead of ~40 Jacobi rotations: this runs
// on ONE thread at the end of every ICP iteration.
B2_HD void sym4_max_eigvec(const double Q[4][4], double v[4]) {
  double c[5] = {1.0, 0.0, 0.0, 0.0, 0.0};   // x^4 + c1 x^3 + c2 x^2 + c3 x + c4
  double M[4][4], T[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) M[i][j] = 0.0;
  double fro = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) fro += Q[i][j] * Q[i][j];
  for (int k = 1; k <= 4; ++k) {   // M_k = Q M_{k-1} + c_{k-1} I ; c_k = -tr(Q M_k)/k
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double t = Q[i][0] * M[0][j] + Q[i][1] * M[1][j] + Q[i][2] * M[2][j] + Q[i][3] * M[3][j];
        T[i][j] = t + (i == j ? c[k - 1] : 0.0);
      }
    double tr = 0.0;
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < 4; ++j) M[i][j] = T[i][j];
      tr += Q[i][0] * T[0][i] + Q[i][1] * T[1][i] + Q[i][2] * T[2][i] + Q[i][3] * T[3][i];
    }
    c[k] = -tr / (double)k;
  }
  double x = sqrt(fro) * (1.0 + 1e-12) + 1e-300;   // >= spectral radius >= lambda_max
  for (int it = 0; it < 100; ++it) {
    const double p = (((x + c[1]) * x + c[2]) * x + c[3]) * x + c[4];
    const double dp = ((4.0 * x + 3.0 * c[1]) * x + 2.0 * c[2]) * x + c[3];
    if (!(dp > 0.0)) break;
    const double xn = x - p / dp;
    if (!(xn < x)) break;          // converged to rounding (Newton approaches from the right)
    const bool tiny = (x - xn) <= 4e-16 * fabs(x);
    x = xn;
    if (tiny) break;
  }
  double A[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) A[i][j] = Q[i][j] - (i == j ? x : 0.0);
  // cofactor C(r,cidx) of the symmetric A
  auto cof = [&](int r, int cc) {
    int ri[3], ci[3], a = 0, b = 0;
    for (int i = 0; i < 4; ++i) { if (i != r) ri[a++] = i; if (i != cc) ci[b++] = i; }
    const double d = det3(A[ri[0]][ci[0]], A[ri[0]][ci[1]], A[ri[0]][ci[2]], A[ri[1]][ci[0]], A[ri[1]][ci[1]],
                          A[ri[1]][ci[2]], A[ri[2]][ci[0]], A[ri[2]][ci[1]], A[ri[2]][ci[2]]);
    return ((r + cc) & 1) ? -d : d;
  };
  double dg[4];
  int best = 0;
  for (int i = 0; i < 4; ++i) { dg[i] = cof(i, i); if (fabs(dg[i]) > fabs(dg[best])) best = i; }
  for (int i = 0; i < 4; ++i) v[i] = i == best ? dg[i] : cof(i, best);
  const double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  if (!(n2 > 0.0)) { v[0] = 1.0; v[1] = v[2] = v[3] = 0.0; }   // Q == lambda I: any rotation is optimal
}

// rotation from the three small-angle sines (icp6Dapx.cc:104-121 == EulerToMatrix4 with sines given)
B2_HD void rot_from_sines(const double* x, double* M) {
  double sx = x[0], cx = sqrt(1.0 - sx * sx);
  double sy = x[1], cy = sqrt(1.0 - sy * sy);
  double sz = x[2], cz = sqrt(1.0 - sz * sz);
  M[0] = cy * cz;
  M[1] = sx * sy * cz + cx * sz;
  M[2] = -cx * sy * cz + sx * sz;
  M[3] = 0;
  M[4] = -cy * sz;
  M[5] = -sx * sy * sz + cx * cz;
  M[6] = cx * sy * sz + sx * cz;
  M[7] = 0;
  M[8] = sy;
  M[9] = -sx * cy;
  M[10] = cx * cy;
  M[11] = 0;
  M[15] = 1;
}

B2_HD void set_translation_from_centroids(double* M, const double* cm, const double* cd) {
  // t = cm - R cd   (icp6Dquat.cc:135-141)
  M[12] = cm[0] - M[0] * cd[0] - M[4] * cd[1] - M[8] * cd[2];
  M[13] = cm[1] - M[1] * cd[0] - M[5] * cd[1] - M[9] * cd[2];
  M[14] = cm[2] - M[2] * cd[0] - M[6] * cd[1] - M[10] * cd[2];
}

// ---- the minimizers ---------------------------------------------------------------------------
// All take moments `mom` about shift origin `o` and write a column-major alignxf.  Return the RMS the
// reference's Align returns (sqrt(sum/n)), or -1.0 when a Cholesky pivot check fails.

// Horn: S -> 4x4 N -> eigenvector of the largest eigenvalue (icp6Dquat.cc:86-141).  The reference gets
// lambda_max from the characteristic quartic (Ferrari) and the vector from an LU solve; a Jacobi
// eigen-decomposition of the same symmetric matrix yields the same unit quaternion up to sign.
B2_HD double solve_quat(const double* mom, const double* o, double* alignxf) {
  const double n = mom[MP_N];
  const double inv = 1.0 / n;
  double cm[3], cd[3];
  for (int i = 0; i < 3; ++i) { cm[i] = mom[MP_M + i] * inv; cd[i] = mom[MP_D + i] * inv; }
  double S[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) S[i][j] = mom[MP_DM + 3 * i + j] * inv - cd[i] * cm[j];
  double trace = S[0][0] + S[1][1] + S[2][2];
  double Q[4][4];
  Q[0][0] = trace;
  Q[0][1] = Q[1][0] = S[1][2] - S[2][1];
  Q[0][2] = Q[2][0] = S[2][0] - S[0][2];
  Q[0][3] = Q[3][0] = S[0][1] - S[1][0];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Q[i + 1][j + 1] = S[i][j] + S[j][i] - (i == j ? trace : 0.0);
  double qv[4];
  sym4_max_eigvec(Q, qv);
  double q0 = qv[0], q1 = qv[1], q2 = qv[2], q3 = qv[3];
  double ql = 1.0 / sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  q0 *= ql; q1 *= ql; q2 *= ql; q3 *= ql;
  // quaternion -> rotation (icp6Dquat.cc:148-169)
  double q00 = q0 * q0, q11 = q1 * q1, q22 = q2 * q2, q33 = q3 * q3;
  double q03 = q0 * q3, q13 = q1 * q3, q23 = q2 * q3, q02 = q0 * q2, q12 = q1 * q2, q01 = q0 * q1;
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
  m4_identity(alignxf);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) alignxf[4 * c + r] = R[r][c];
  double cmo[3] = {cm[0] + o[0], cm[1] + o[1], cm[2] + o[2]};
  double cdo[3] = {cd[0] + o[0], cd[1] + o[1], cd[2] + o[2]};
  set_translation_from_centroids(alignxf, cmo, cdo);
  return sqrt(mom[MP_D2] * inv);
}

// Arun: H = sum d' m'^T (centred), H = U L V^T, R = V U^T, reflection fix (icp6Dsvd.cc:79-115).
B2_HD double solve_svd(const double* mom, const double* o, double* alignxf) {
  const double n = mom[MP_N];
  const double inv = 1.0 / n;
  double cm[3], cd[3];
  for (int i = 0; i < 3; ++i) { cm[i] = mom[MP_M + i] * inv; cd[i] = mom[MP_D + i] * inv; }
  double H[3][3], U[3][3], V[3][3], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i][j] = mom[MP_DM + 3 * i + j] - n * cd[i] * cm[j];
  svd3(H, U, w, V);
  double R[3][3];
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i][j] = V[i][0] * U[j][0] + V[i][1] * U[j][1] + V[i][2] * U[j][2];
    double det = det3(R[0][0], R[0][1], R[0][2], R[1][0], R[1][1], R[1][2], R[2][0], R[2][1], R[2][2]);
    if (det >= 0.0) break;
    V[0][2] = -V[0][2]; V[1][2] = -V[1][2]; V[2][2] = -V[2][2];
  }
  m4_identity(alignxf);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) alignxf[4 * c + r] = R[r][c];
  double cmo[3] = {cm[0] + o[0], cm[1] + o[1], cm[2] + o[2]};
  double cdo[3] = {cd[0] + o[0], cd[1] + o[1], cd[2] + o[2]};
  set_translation_from_centroids(alignxf, cmo, cdo);
  return sqrt(mom[MP_D2] * inv);
}

// ORTHO (icp6Dortho.cc:41-153, Horn/Hilden/Negahdaripour orthonormal matrices): H = sum m' d'^T (centred),
// R = H (H^T H)^(-1/2) through the eigen-decomposition of the symmetric H^T H, t = cm - R cd.
B2_HD double solve_ortho(const double* mom, const double* o, double* alignxf) {
  const double n = mom[MP_N];
  const double inv = 1.0 / n;
  double cm[3], cd[3];
  for (int i = 0; i < 3; ++i) { cm[i] = mom[MP_M + i] * inv; cd[i] = mom[MP_D + i] * inv; }
  double H[3][3], HH[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i][j] = mom[MP_DM + 3 * j + i] - n * cm[i] * cd[j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) HH[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];
  jacobi_eig<3>(HH, V);
  double W[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // sum_k v_k v_k^T / sqrt(lambda_k)
  for (int k = 0; k < 3; ++k) {
    const double f = 1.0 / sqrt(HH[k][k]);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) W[i][j] += V[i][k] * V[j][k] * f;
  }
  m4_identity(alignxf);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) alignxf[4 * c + r] = H[r][0] * W[0][c] + H[r][1] * W[1][c] + H[r][2] * W[2][c];
  double cmo[3] = {cm[0] + o[0], cm[1] + o[1], cm[2] + o[2]};
  double cdo[3] = {cd[0] + o[0], cd[1] + o[1], cd[2] + o[2]};
  set_translation_from_centroids(alignxf, cmo, cdo);
  return sqrt(mom[MP_D2] * inv);
}

// DUAL (icp6Ddual.cc:41-150, Walker/Shao/Volz dual quaternions).  Every entry of the reference's C1 / C2 is
// linear in M = sum m d^T and in sum m, sum d (Cm d = m x d, Cm Cd = d m^T - (m.d) I), so the per-pair walk
// collapses to the moments.  Evaluated in the shifted frame (origin o) and mapped back: t = t' + o - R o -- the
// least-squares optimum is the same, the reference evaluates it with absolute coordinates.
// The rotation quaternion is the eigenvector of A with the largest |eigenvalue| (the reference takes column 1 of
// newmat's SVD of the symmetric A, i.e. the largest singular value).
B2_HD double solve_dual(const double* mom, const double* o, double* alignxf) {
  const double n = mom[MP_N];
  const double inv = 1.0 / n;
  double M[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i][j] = mom[MP_DM + 3 * j + i];      // sum m_i d_j
  const double tr = M[0][0] + M[1][1] + M[2][2];
  const double cr[3] = {M[1][2] - M[2][1], M[2][0] - M[0][2], M[0][1] - M[1][0]};   // sum m x d
  double C1[4][4], C2[4][4];
  C1[0][0] = tr;
  for (int i = 0; i < 3; ++i) { C1[0][i + 1] = -cr[i]; C1[i + 1][0] = -cr[i]; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C1[i + 1][j + 1] = M[i][j] + M[j][i] - (i == j ? tr : 0.0);
  const double dm[3] = {mom[MP_M] - mom[MP_D], mom[MP_M + 1] - mom[MP_D + 1], mom[MP_M + 2] - mom[MP_D + 2]};
  const double sp[3] = {mom[MP_M] + mom[MP_D], mom[MP_M + 1] + mom[MP_D + 1], mom[MP_M + 2] + mom[MP_D + 2]};
  C2[0][0] = 0.0;
  for (int i = 0; i < 3; ++i) { C2[0][i + 1] = dm[i]; C2[i + 1][0] = -dm[i]; }
  // -(Cd + Cm) = -[sp]x
  C2[1][1] = 0.0;    C2[1][2] = sp[2];  C2[1][3] = -sp[1];
  C2[2][1] = -sp[2]; C2[2][2] = 0.0;    C2[2][3] = sp[0];
  C2[3][1] = sp[1];  C2[3][2] = -sp[0]; C2[3][3] = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { C1[i][j] *= -2.0; C2[i][j] *= 2.0; }
  double A[4][4], V[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double t = 0.0;
      for (int k = 0; k < 4; ++k) t += C2[k][i] * C2[k][j];
      A[i][j] = (t * (0.5 * inv) - C1[i][j] - C1[j][i]) * 0.5;
    }
  jacobi_eig<4>(A, V);
  int best = 0;
  for (int k = 1; k < 4; ++k) if (fabs(A[k][k]) > fabs(A[best][best])) best = k;
  const double q0 = V[0][best], q[3] = {V[1][best], V[2][best], V[3][best]};
  const double qd[4] = {q0, q[0], q[1], q[2]};
  double sv[4];
  for (int i = 0; i < 4; ++i) {
    double t = 0.0;
    for (int k = 0; k < 4; ++k) t += C2[i][k] * qd[k];
    sv[i] = -t * (0.5 * inv);
  }
  // p = Q s with Q = [[q0, q^T], [-q, q0 I + Cq]]; translation = p[1..3]
  const double tp[3] = {-q[0] * sv[0] + q0 * sv[1] - q[2] * sv[2] + q[1] * sv[3],
                        -q[1] * sv[0] + q[2] * sv[1] + q0 * sv[2] - q[0] * sv[3],
                        -q[2] * sv[0] - q[1] * sv[1] + q[0] * sv[2] + q0 * sv[3]};
  const double qq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  double R[3][3];
  const double Cq[3][3] = {{0, -q[2], q[1]}, {q[2], 0, -q[0]}, {-q[1], q[0], 0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = (i == j ? q0 * q0 - qq : 0.0) + 2.0 * q[i] * q[j] + 2.0 * q0 * Cq[i][j];
  m4_identity(alignxf);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) alignxf[4 * c + r] = R[r][c];
  for (int r = 0; r < 3; ++r)
    alignxf[12 + r] = tp[r] + o[r] - (R[r][0] * o[0] + R[r][1] * o[1] + R[r][2] * o[2]);
  return sqrt(mom[MP_D2] * inv);
}

// Gaussian elimination with partial pivoting, A x = b in place (x returned in b); false when singular.
template <int N>
B2_HD bool gauss_solve(double A[N][N], double* b) {
  for (int c = 0; c < N; ++c) {
    int piv = c;
    for (int r = c + 1; r < N; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (A[piv][c] == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < N; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
      const double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int k = c; k < N; ++k) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  for (int r = N - 1; r >= 0; --r) {
    double t = b[r];
    for (int k = r + 1; k < N; ++k) t -= A[r][k] * b[k];
    b[r] = t / A[r][r];
  }
  return true;
}

// HELIX (icp6Dhelix.cc:48-205, Pottmann/Leopoldseder/Hofer): least squares over the linearised motion
// x -> x + cs + c x x, then the helical motion with that axis, angle atan|c| and pitch (computeRt).  The 6x6
// system only needs sum p2, sum p2 p2^T, sum p2 x p1 and sum (p2 - p1): all in the pair moments.  Solved in the
// shifted frame (the family of linearised motions does not depend on the origin, the helical motion is a
// geometric object) and mapped back: t = t' + o - R o.
B2_HD double solve_helix(const double* mom, const double* o, double* alignxf) {
  const double n = mom[MP_N];
  const double* D = mom + MP_D;
  const double* DM = mom + MP_DM;      // sum b_i a_j, b = p2 - o, a = p1 - o
  const double xx = mom[MP_DD], xy = mom[MP_DD + 1], xz = mom[MP_DD + 2], yy = mom[MP_DD + 3], yz = mom[MP_DD + 4],
               zz = mom[MP_DD + 5];
  double B[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) B[i][j] = 0.0;
  B[0][0] = zz + yy; B[1][1] = zz + xx; B[2][2] = xx + yy;
  B[0][1] = B[1][0] = -xy; B[0][2] = B[2][0] = -xz; B[1][2] = B[2][1] = -yz;
  B[3][3] = B[4][4] = B[5][5] = n;
  B[0][4] = B[4][0] = -D[2]; B[1][3] = B[3][1] = D[2];
  B[0][5] = B[5][0] = D[1];  B[2][3] = B[3][2] = -D[1];
  B[2][4] = B[4][2] = D[0];  B[1][5] = B[5][1] = -D[0];
  double x[6] = {DM[7] - DM[5], DM[2] - DM[6], DM[3] - DM[1],
                 mom[MP_D] - mom[MP_M], mom[MP_D + 1] - mom[MP_M + 1], mom[MP_D + 2] - mom[MP_M + 2]};
  gauss_solve<6>(B, x);
  const double c[3] = {-x[0], -x[1], -x[2]}, cs[3] = {-x[3], -x[4], -x[5]};
  const double cl = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  const double chk = c[0] * cs[0] + c[1] * cs[1] + c[2] * cs[2];
  const double angle = atan(cl);
  const double g[3] = {c[0] / cl, c[1] / cl, c[2] / cl};
  const double sa = sin(-angle / 2), b0 = cos(-angle / 2), b1 = g[0] * sa, b2 = g[1] * sa, b3 = g[2] * sa;
  double R[3][3];
  R[0][0] = b0 * b0 + b1 * b1 - b2 * b2 - b3 * b3; R[0][1] = 2 * (b1 * b2 + b0 * b3); R[0][2] = 2 * (b1 * b3 - b0 * b2);
  R[1][0] = 2 * (b1 * b2 - b0 * b3); R[1][1] = b0 * b0 - b1 * b1 + b2 * b2 - b3 * b3; R[1][2] = 2 * (b2 * b3 + b0 * b1);
  R[2][0] = 2 * (b1 * b3 + b0 * b2); R[2][1] = 2 * (b2 * b3 - b0 * b1); R[2][2] = b0 * b0 - b1 * b1 - b2 * b2 + b3 * b3;
  const double nn = b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] /= nn;
  const double skew = chk / (cl * cl);
  const double gs[3] = {(cs[0] - c[0] * skew) / cl, (cs[1] - c[1] * skew) / cl, (cs[2] - c[2] * skew) / cl};
  const double pt[3] = {g[1] * gs[2] - g[2] * gs[1], g[2] * gs[0] - g[0] * gs[2], g[0] * gs[1] - g[1] * gs[0]};
  m4_identity(alignxf);
  for (int cc = 0; cc < 3; ++cc)
    for (int r = 0; r < 3; ++r) alignxf[4 * cc + r] = R[r][cc];
  for (int r = 0; r < 3; ++r) {
    const double tp = -(R[r][0] * pt[0] + R[r][1] * pt[1] + R[r][2] * pt[2]) + g[r] * (skew * angle) + pt[r];
    alignxf[12 + r] = tp + o[r] - (R[r][0] * o[0] + R[r][1] * o[1] + R[r][2] * o[2]);
  }
  return sqrt(mom[MP_D2] / n);
}

// APX (icp6Dapx.cc:35-133): A x = B over centred data-side second moments.
B2_HD double solve_apx(const double* mom, const double* o, double* alignxf) {
  const double n = mom[MP_N];
  if (n <= 3.0) { m4_identity(alignxf); return 0.0; }
  const double inv = 1.0 / n;
  double cm[3], cd[3];
  for (int i = 0; i < 3; ++i) { cm[i] = mom[MP_M + i] * inv; cd[i] = mom[MP_D + i] * inv; }
  // C[a][b] = sum (p2-cd)_a (p2-cd)_b ;  G[a][b] = sum (p1-p2)_a (p2-cd)_b
  double C[3][3], G[3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      C[a][b] = mom[MP_DD + sym6(a, b)] - n * cd[a] * cd[b];
      double s12a = mom[MP_M + a] - mom[MP_D + a];  // sum (p1-p2)_a
      G[a][b] = (mom[MP_DM + 3 * b + a] - mom[MP_DD + sym6(a, b)]) - cd[b] * s12a;
    }
  double A[3][3] = {{C[1][1] + C[2][2], -C[0][1], -C[0][2]},
                    {0.0, C[0][0] + C[2][2], -C[1][2]},
                    {0.0, 0.0, C[0][0] + C[1][1]}};
  double B[3] = {G[2][1] - G[1][2], G[0][2] - G[2][0], G[1][0] - G[0][1]};
  double x[3];
  if (!cholesky_solve<3>(A, B, x)) return -1.0;
  rot_from_sines(x, alignxf);
  double cmo[3] = {cm[0] + o[0], cm[1] + o[1], cm[2] + o[2]};
  double cdo[3] = {cd[0] + o[0], cd[1] + o[1], cd[2] + o[2]};
  set_translation_from_centroids(alignxf, cmo, cdo);
  return sqrt(mom[MP_D2] * inv);
}

// NAPX (icp6Dnapx.cc:34-149).  weighted == 0 reproduces the shipped right-hand side B = sum [c; n];
// weighted == 1 is the least-squares form B = sum d [c; n].
B2_HD double solve_napx(const double* mom, const double* o, int weighted, double* alignxf) {
  const double n = mom[MN_N];
  const double inv = 1.0 / n;
  double cd[3];
  for (int i = 0; i < 3; ++i) cd[i] = mom[MN_D + i] * inv;
  // K = [cd]x : (K v) = cd x v
  double K[3][3] = {{0, -cd[2], cd[1]}, {cd[2], 0, -cd[0]}, {-cd[1], cd[0], 0}};
  double AA[3][3], AN[3][3], NN[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      AA[i][j] = mom[MN_AA + sym6(i, j)];
      NN[i][j] = mom[MN_NN + sym6(i, j)];
      AN[i][j] = mom[MN_AN + 3 * i + j];
    }
  // KN = K NN ; CN = AN - K NN ; CC = AA - AN K^T - K AN^T + K NN K^T
  double KN[3][3], CN[3][3], CC[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      KN[i][j] = K[i][0] * NN[0][j] + K[i][1] * NN[1][j] + K[i][2] * NN[2][j];
      CN[i][j] = AN[i][j] - KN[i][j];
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double anKt = AN[i][0] * K[j][0] + AN[i][1] * K[j][1] + AN[i][2] * K[j][2];
      double kanT = K[i][0] * AN[j][0] + K[i][1] * AN[j][1] + K[i][2] * AN[j][2];
      double knkT = KN[i][0] * K[j][0] + KN[i][1] * K[j][1] + KN[i][2] * K[j][2];
      CC[i][j] = AA[i][j] - anKt - kanT + knkT;
    }
  double A[6][6], B[6], x[6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) A[i][j] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      if (j >= i) { A[i][j] = CC[i][j]; A[3 + i][3 + j] = NN[i][j]; }
      A[i][3 + j] = CN[i][j];
    }
  const double* sa = mom + (weighted ? MN_DA : MN_A);
  const double* sn = mom + (weighted ? MN_DN : MN_NRM);
  B[0] = sa[0] - (cd[1] * sn[2] - cd[2] * sn[1]);
  B[1] = sa[1] - (cd[2] * sn[0] - cd[0] * sn[2]);
  B[2] = sa[2] - (cd[0] * sn[1] - cd[1] * sn[0]);
  B[3] = sn[0]; B[4] = sn[1]; B[5] = sn[2];
  if (!cholesky_solve<6>(A, B, x)) return -1.0;
  rot_from_sines(x, alignxf);
  double cdo[3] = {cd[0] + o[0], cd[1] + o[1], cd[2] + o[2]};
  alignxf[12] = x[3] + cdo[0] - alignxf[0] * cdo[0] - alignxf[4] * cdo[1] - alignxf[8] * cdo[2];
  alignxf[13] = x[4] + cdo[1] - alignxf[1] * cdo[0] - alignxf[5] * cdo[1] - alignxf[9] * cdo[2];
  alignxf[14] = x[5] + cdo[2] - alignxf[2] * cdo[0] - alignxf[6] * cdo[1] - alignxf[10] * cdo[2];
  return sqrt(mom[MN_D2] * inv);
}

B2_HD int moment_count(int algo) { return algo == 10 ? (int)NS_NAPX : (int)NS_P2P; }

B2_HD double solve_any(int algo, const double* mom, const double* o, int napx_weighted,
                       double* alignxf) {
  switch (algo) {
    case 1: return solve_quat(mom, o, alignxf);
    case 2: return solve_svd(mom, o, alignxf);
    case 3: return solve_ortho(mom, o, alignxf);
    case 4: return solve_dual(mom, o, alignxf);
    case 5: return solve_helix(mom, o, alignxf);
    case 6: return solve_apx(mom, o, alignxf);
    case 10: return solve_napx(mom, o, napx_weighted, alignxf);
    default: return -2.0;
  }
}

// Per-pair accumulation, shared by the host path (b200icp_align_pairs) and the kernels.
// dd: also sum p2' p2'^T (MP_DD) -- only the HELIX and APX minimizers read it
template <class Acc>
B2_HD void accumulate_p2p(Acc&& acc, const double* p1, const double* p2, const double* o, bool dd = true) {
  double a[3] = {p1[0] - o[0], p1[1] - o[1], p1[2] - o[2]};
  double b[3] = {p2[0] - o[0], p2[1] - o[1], p2[2] - o[2]};
  double e0 = p1[0] - p2[0], e1 = p1[1] - p2[1], e2 = p1[2] - p2[2];
  acc[MP_N] += 1.0;
  acc[MP_D2] += e0 * e0 + e1 * e1 + e2 * e2;
  for (int i = 0; i < 3; ++i) { acc[MP_M + i] += a[i]; acc[MP_D + i] += b[i]; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) acc[MP_DM + 3 * i + j] += b[i] * a[j];
  if (dd) {
    acc[MP_DD + 0] += b[0] * b[0]; acc[MP_DD + 1] += b[0] * b[1]; acc[MP_DD + 2] += b[0] * b[2];
    acc[MP_DD + 3] += b[1] * b[1]; acc[MP_DD + 4] += b[1] * b[2]; acc[MP_DD + 5] += b[2] * b[2];
  }
}

// minimizers whose solve reads MP_DD
B2_HD bool algo_needs_dd(int algo) { return algo == 5 || algo == 6; }

template <class Acc>
B2_HD void accumulate_napx(Acc&& acc, const double* p1, const double* p2, const double* nrm,
                           const double* o) {
  double b[3] = {p2[0] - o[0], p2[1] - o[1], p2[2] - o[2]};
  double e0 = p1[0] - p2[0], e1 = p1[1] - p2[1], e2 = p1[2] - p2[2];
  double d = e0 * nrm[0] + e1 * nrm[1] + e2 * nrm[2];
  double a[3] = {b[1] * nrm[2] - b[2] * nrm[1], b[2] * nrm[0] - b[0] * nrm[2],
                 b[0] * nrm[1] - b[1] * nrm[0]};
  acc[MN_N] += 1.0;
  acc[MN_D2] += d * d;
  acc[MN_PD2] += e0 * e0 + e1 * e1 + e2 * e2;
  for (int i = 0; i < 3; ++i) {
    acc[MN_M + i] += p1[i] - o[i];
    acc[MN_D + i] += b[i];
    acc[MN_A + i] += a[i];
    acc[MN_NRM + i] += nrm[i];
    acc[MN_DA + i] += d * a[i];
    acc[MN_DN + i] += d * nrm[i];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) acc[MN_AN + 3 * i + j] += a[i] * nrm[j];
  acc[MN_AA + 0] += a[0] * a[0]; acc[MN_AA + 1] += a[0] * a[1]; acc[MN_AA + 2] += a[0] * a[2];
  acc[MN_AA + 3] += a[1] * a[1]; acc[MN_AA + 4] += a[1] * a[2]; acc[MN_AA + 5] += a[2] * a[2];
  acc[MN_NN + 0] += nrm[0] * nrm[0]; acc[MN_NN + 1] += nrm[0] * nrm[1]; acc[MN_NN + 2] += nrm[0] * nrm[2];
  acc[MN_NN + 3] += nrm[1] * nrm[1]; acc[MN_NN + 4] += nrm[1] * nrm[2]; acc[MN_NN + 5] += nrm[2] * nrm[2];
}

}  // namespace b200
