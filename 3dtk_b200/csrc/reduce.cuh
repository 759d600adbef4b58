// reduce.cuh -- octree voxel reduction (the step right before the hot path): centre, average
// and one-random-point-per-voxel extraction.
//
// Replaces Scan::calcReducedPoints + BOctTree construction + GetOctTreeCenter for `-r <voxel>` with the default
// `-O 0` (reference src/slam6d/scan.cc:560-601; include/slam6d/Boctree.h:224-270 root cube, :612-656 child
// centres, :1353-1355 child index, :1164-1195 leaf rule, :928-949 centre extraction).
// The octree is never built: a point's path from the root is a sequence of 3-bit child indices
// (bit k set iff !(p[k] < centre[k]): Scan::calcReducedPoints builds the tree through the T** constructor, whose
// partition keeps `p < split` on the lower side, Boctree.h:268,1784-1815 -- a point ON a splitting plane goes to the
// upper child; pinned by a known-answer test against the compiled reference), i.e. a Morton-like key whose numeric order IS the reference's
// depth-first output order.  Keys are computed with the reference's own centre arithmetic
// (centre +- size/2.0, size halved per level) so points on a splitting plane fall on the same side, then
// radix-sorted and made unique; the leaf-cube centre is re-derived from the key by the same walk.
// Average / random extraction (GetOctTreeAvg Boctree.h:951-983, GetOctTreeRandom :985-1019, called from
// scan.cc:585-601 for `-O -1` / `-O 1`): a stable sort of (key, row) groups the points of a voxel (in input order);
// one thread per voxel walks them like the reference's loop: sequential fp64 sums divided by the count (all
// attributes, i.e. xyz and, with a normal-carrying PointType, the normals: scan.cc:544-557,652-676), or point number
// (int)(length * rand() / (RAND_MAX + 1.0)) with the k-th voxel in depth-first order consuming the k-th value of the
// C library's rand() stream (globals.icc:607-610).  The reference keeps a leaf's points in the order its unstable
// (Hoare) partitions leave them, not in input order: averages agree to rounding (~1e-15), the random mode picks the
// same NUMBER in every voxel but counts it through a differently ordered list -- another point of the same voxel.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include "common.cuh"

namespace b200 {

struct OctRoot {
  double c[3];
  double size;   // half-size of the root cube: max half-extent + 1.0
  int levels;    // number of subdivisions until the child half-size is <= voxel (>= 1)
};

__global__ void oct_key_kernel(const double* __restrict__ xyz, size_t n, OctRoot root,
                               unsigned long long* __restrict__ keys) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
  double cx = root.c[0], cy = root.c[1], cz = root.c[2], s = root.size;
  unsigned long long key = 0;
  for (int l = 0; l < root.levels; ++l) {
    const unsigned bx = !(px < cx), by = !(py < cy), bz = !(pz < cz);
    const double hs = s / 2.0;
    cx = bx ? cx + hs : cx - hs;
    cy = by ? cy + hs : cy - hs;
    cz = bz ? cz + hs : cz - hs;
    s = hs;
    key = (key << 3) | (unsigned long long)(bx | (by << 1) | (bz << 2));
  }
  keys[i] = key;
}

__global__ void oct_centre_kernel(const unsigned long long* __restrict__ keys, size_t m, OctRoot root,
                                  double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long key = keys[i];
  double cx = root.c[0], cy = root.c[1], cz = root.c[2], s = root.size;
  for (int l = root.levels - 1; l >= 0; --l) {
    const unsigned c = (unsigned)(key >> (3 * l)) & 7u;
    const double hs = s / 2.0;
    cx = (c & 1u) ? cx + hs : cx - hs;
    cy = (c & 2u) ? cy + hs : cy - hs;
    cz = (c & 4u) ? cz + hs : cz - hs;
    s = hs;
  }
  out[3 * i] = cx; out[3 * i + 1] = cy; out[3 * i + 2] = cz;
}

__global__ void iota_kernel(uint32_t* __restrict__ v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

// flags[i] = 1 where a new voxel starts in the sorted key array
__global__ void oct_heads_kernel(const unsigned long long* __restrict__ keys, size_t n, unsigned char* __restrict__ flags) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per voxel: heads[v] .. heads[v+1] (or n) are its points in input order (rows[] = original row)
// MODE -1: average (GetOctTreeAvg), MODE 1: one random point (GetOctTreeRandom(c)), rnd[v] = v-th rand() value
template <int MODE>
__global__ void oct_extract_kernel(const double* __restrict__ xyz, const double* __restrict__ nrm,
                                   const uint32_t* __restrict__ rows, const uint32_t* __restrict__ heads, size_t m,
                                   size_t n, const int* __restrict__ rnd, double* __restrict__ out_xyz,
                                   double* __restrict__ out_nrm) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= m) return;
  const size_t a = heads[v], b = v + 1 < m ? heads[v + 1] : n;
  const unsigned length = (unsigned)(b - a);
  if (MODE == 1) {
    // inline int rand(int rnd) { return (int)((double)rnd * (double)std::rand() / (RAND_MAX + 1.0)); }
    const int pick = (int)((double)(int)length * (double)rnd[v] / (2147483647.0 + 1.0));
    const size_t r = rows[a + (size_t)pick];
    for (int k = 0; k < 3; ++k) out_xyz[3 * v + k] = xyz[3 * r + k];
    if (nrm && out_nrm)
      for (int k = 0; k < 3; ++k) out_nrm[3 * v + k] = nrm[3 * r + k];
    return;
  }
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (size_t j = a; j < b; ++j) {
    const size_t r = rows[j];
    s[0] += xyz[3 * r]; s[1] += xyz[3 * r + 1]; s[2] += xyz[3 * r + 2];
    if (nrm) { s[3] += nrm[3 * r]; s[4] += nrm[3 * r + 1]; s[5] += nrm[3 * r + 2]; }
  }
  for (int k = 0; k < 3; ++k) out_xyz[3 * v + k] = s[k] / length;     // avgp[j] /= length (unsigned -> double)
  if (nrm && out_nrm)
    for (int k = 0; k < 3; ++k) out_nrm[3 * v + k] = s[3 + k] / length;
}

}  // namespace b200
