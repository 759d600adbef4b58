// reduce.cuh -- octree voxel-CENTRE reduction (the step right before the hot path).
//
// Replaces Scan::calcReducedPoints + BOctTree construction + GetOctTreeCenter for `-r <voxel>` with the default
// `-O 0` (reference src/slam6d/scan.cc:560-601; include/slam6d/Boctree.h:224-270 root cube, :612-656 child
// centres, :1353-1355 child index, :1164-1195 leaf rule, :928-949 centre extraction).
// The octree is never built: a point's path from the root is a sequence of 3-bit child indices
// (bit k set iff p[k] > centre[k], strict), i.e. a Morton-like key whose numeric order IS the reference's
// depth-first output order.  Keys are computed with the reference's own centre arithmetic
// (centre +- size/2.0, size halved per level) so points on a splitting plane fall on the same side, then
// radix-sorted and made unique; the leaf-cube centre is re-derived from the key by the same walk.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
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
    const unsigned bx = px > cx, by = py > cy, bz = pz > cz;
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

}  // namespace b200
