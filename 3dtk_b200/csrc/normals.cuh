// normals.cuh -- exact k-NN PCA normals on the same uniform grid.
//
// Replaces calculateNormalsKNN + calculateNormal (reference src/slam6d/normals.cc:220-295, :518-558):
// for every point, its k nearest neighbours (itself included, as KDtree::kNearestNeighbors returns it
// at distance 0), mean, covariance 1/k X^T X, eigenvector of the smallest eigenvalue, flipped so that
// n . (p - rPos) >= 0, unit length.
// One thread per point (points are cell-sorted, so a warp walks neighbouring cells together).  Cube
// shells of cells are visited outwards; the search ends when k hits are inside the radius the finished
// shells guarantee.  Candidates are screened in fp32 against the current k-th distance (with the same
// error bound as nn_search.cuh) and ranked by their exact fp64 distance.
#pragma once
#include "nn_search.cuh"

namespace b200 {

template <int K>
struct KnnList {
  double d2[K];
  int id[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; ++i) { d2[i] = 1.0e300; id[i] = -1; }
  }
  // keep ascending order; (d2, id) lexicographic so the result does not depend on visiting order
  __device__ __forceinline__ void insert(double v, int j) {
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      const bool here = v < d2[i] || (v == d2[i] && j < id[i]);
      const bool above = v < d2[i - 1] || (v == d2[i - 1] && j < id[i - 1]);
      if (here) {
        d2[i] = above ? d2[i - 1] : v;
        id[i] = above ? id[i - 1] : j;
      }
    }
    if (v < d2[0] || (v == d2[0] && j < id[0])) { d2[0] = v; id[0] = j; }
  }
};

template <int K>
__device__ __forceinline__ void knn_scan_range(const GridDev& g, unsigned beg, unsigned end, float qx,
                                               float qy, float qz, double sx, double sy, double sz,
                                               float e, int k, KnnList<K>& L, float& thr) {
  for (unsigned j = beg; j < end; ++j) {
    const float4 p = __ldg(g.p32 + j);
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    if (d2 < thr) {
      const double2 pa = __ldg(reinterpret_cast<const double2*>(g.p64 + j));
      const double pz = __ldg(reinterpret_cast<const double*>(g.p64 + j) + 2);
      const double ex = pa.x - sx, ey = pa.y - sy, ez = pz - sz;
      const double v = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
      // rank by original row on exact ties (ids here are sorted positions; rows via p.w)
      L.insert(v, (int)j);
      const double kth = L.d2[k - 1];
      thr = kth >= 1.0e299 ? 3.0e38f : filter_bound<true>(kth, e);
    }
  }
}

// out: 3 doubles per point in ORIGINAL row order (perm given), or -- perm == nullptr -- a double4 per point in the
// scan's own cell-sorted order (written straight into the scan's "normal reduced" array)
template <int K>
__global__ void __launch_bounds__(128)
normals_knn_kernel(GridDev g, const uint32_t* __restrict__ perm, int k, double rx, double ry, double rz,
                   double* __restrict__ out) {
  const uint32_t j0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (j0 >= g.n) return;
  const double4 P = g.p64[j0];
  const double sx = P.x, sy = P.y, sz = P.z;
  const float qx = (float)(sx - g.c[0]), qy = (float)(sy - g.c[1]), qz = (float)(sz - g.c[2]);
  const float e = 1.25e-7f * (2.f * g.bmax) + 1e-30f;
  const int cx = cell_coord(sx, g.g0[0], g.inv_h), cy = cell_coord(sy, g.g0[1], g.inv_h),
            cz = cell_coord(sz, g.g0[2], g.inv_h);
  const int m = min(k, (int)g.n);
  KnnList<K> L;
  L.init();
  float thr = 3.0e38f;
  const int rmax = max(max(g.nx, g.ny), g.nz);
  for (int r = 0; r <= rmax; ++r) {
    for (int dz = -r; dz <= r; ++dz) {
      const int z = cz + dz;
      if ((unsigned)z >= (unsigned)g.nz) continue;
      for (int dy = -r; dy <= r; ++dy) {
        const int y = cy + dy;
        if ((unsigned)y >= (unsigned)g.ny) continue;
        const size_t row = ((size_t)z * g.ny + y) * g.nx;
        if (max(abs(dy), abs(dz)) == r) {
          const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
          if (x0 <= x1)
            knn_scan_range<K>(g, __ldg(g.cell_start + row + x0), __ldg(g.cell_start + row + x1 + 1), qx, qy,
                              qz, sx, sy, sz, e, m, L, thr);
        } else {
          const int xa = cx - r, xb = cx + r;
          if ((unsigned)xa < (unsigned)g.nx)
            knn_scan_range<K>(g, __ldg(g.cell_start + row + xa), __ldg(g.cell_start + row + xa + 1), qx, qy,
                              qz, sx, sy, sz, e, m, L, thr);
          if ((unsigned)xb < (unsigned)g.nx)
            knn_scan_range<K>(g, __ldg(g.cell_start + row + xb), __ldg(g.cell_start + row + xb + 1), qx, qy,
                              qz, sx, sy, sz, e, m, L, thr);
        }
      }
    }
    const double lim = (double)r * g.h;
    if (L.id[m - 1] >= 0 && L.d2[m - 1] <= lim * lim * (1.0 - 1e-9)) break;
  }
  // ---- PCA (normals.cc:518-558)
  double mean[3] = {0, 0, 0};
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < K; ++i)
    if (i < m && L.id[i] >= 0) {
      const double4 q = g.p64[L.id[i]];
      mean[0] += q.x; mean[1] += q.y; mean[2] += q.z;
      ++cnt;
    }
  const double inv = 1.0 / (double)cnt;
  mean[0] *= inv; mean[1] *= inv; mean[2] *= inv;
  double Cm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, V[3][3];
#pragma unroll
  for (int i = 0; i < K; ++i)
    if (i < m && L.id[i] >= 0) {
      const double4 q = g.p64[L.id[i]];
      const double d[3] = {q.x - mean[0], q.y - mean[1], q.z - mean[2]};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) Cm[a][b] += d[a] * d[b];
    }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) Cm[a][b] *= inv;
  jacobi_eig<3>(Cm, V);
  int lo = 0;
  if (Cm[1][1] < Cm[lo][lo]) lo = 1;
  if (Cm[2][2] < Cm[lo][lo]) lo = 2;
  double n[3] = {V[0][lo], V[1][lo], V[2][lo]};
  const double px = sx - rx, py = sy - ry, pz = sz - rz;
  if (n[0] * px + n[1] * py + n[2] * pz < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  const double nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  if (perm) {
    const size_t dst = perm[j0];
    out[3 * dst] = n[0] / nl; out[3 * dst + 1] = n[1] / nl; out[3 * dst + 2] = n[2] / nl;
  } else {
    reinterpret_cast<double4*>(out)[j0] = make_double4(n[0] / nl, n[1] / nl, n[2] / nl, 0.0);
  }
}

}  // namespace b200
