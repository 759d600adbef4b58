// grid_build.cuh -- builds the radix-binned uniform grid of a scan on the device.
//
// Replaces KDTreeImpl::create (reference include/slam6d/kdTreeImpl.h:82-201), i.e. what
// Scan::createSearchTree (src/slam6d/scan.cc:285-306) triggers once per scan.  Steps:
//   bbox reduce -> choose the cell edge from the measured cell occupancy -> per-point cell id ->
//   stable radix sort of (cell id, row) [cub::DeviceRadixSort: library plumbing, runs once per scan,
//   not in the per-iteration path] -> per-cell counts -> exclusive scan -> gather fp64 / fp32x4 copies.
// Sorting by (cell, original row) makes the layout -- and every fp64 sum taken over it -- reproducible
// from run to run.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cfloat>
#include "common.cuh"

namespace b200 {

constexpr uint64_t kCellCap = 1ull << 24;  // dense cell-table cap (64 MB of uint32)

__global__ void bbox_partial_kernel(const double* __restrict__ xyz, size_t n, double* __restrict__ part) {
  // part[block][6] = min xyz, max xyz
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double v = xyz[3 * i + k];
      lo[k] = fmin(lo[k], v);
      hi[k] = fmax(hi[k], v);
    }
  }
  __shared__ double s[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], m));
      hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], m));
    }
  if (lane == 0)
    for (int k = 0; k < 3; ++k) { s[k][warp] = lo[k]; s[3 + k][warp] = hi[k]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int k = threadIdx.x;
    double v = s[k][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = k < 3 ? fmin(v, s[k][w]) : fmax(v, s[k][w]);
    part[6 * blockIdx.x + k] = v;
  }
}

__global__ void bbox_final_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= 6) return;
  double v = part[k];
  for (int b = 1; b < nblocks; ++b) v = k < 3 ? fmin(v, part[6 * b + k]) : fmax(v, part[6 * b + k]);
  out[k] = v;
}

struct GridShape {
  double g0[3];
  double h, inv_h;
  int nx, ny, nz;
};

__device__ __forceinline__ uint32_t point_cell(const GridShape& gs, double x, double y, double z) {
  int ix = (int)floor((x - gs.g0[0]) * gs.inv_h);
  int iy = (int)floor((y - gs.g0[1]) * gs.inv_h);
  int iz = (int)floor((z - gs.g0[2]) * gs.inv_h);
  ix = min(max(ix, 0), gs.nx - 1);
  iy = min(max(iy, 0), gs.ny - 1);
  iz = min(max(iz, 0), gs.nz - 1);
  return (uint32_t)(((size_t)iz * gs.ny + iy) * gs.nx + ix);
}

// keys[i] = cell id, vals[i] = i, counts[cell]++ ; keys/vals may be null for the occupancy trial
__global__ void cell_key_kernel(const double* __restrict__ xyz, size_t n, GridShape gs,
                                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                uint32_t* __restrict__ counts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = point_cell(gs, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  if (keys) { keys[i] = c; vals[i] = (uint32_t)i; }
  atomicAdd(counts + c, 1u);
}

__global__ void count_occupied_kernel(const uint32_t* __restrict__ counts, size_t ncells,
                                      unsigned long long* __restrict__ out) {
  unsigned long long local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += (size_t)gridDim.x * blockDim.x)
    local += counts[i] != 0u;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) local += __shfl_xor_sync(0xffffffffu, local, m);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

__global__ void gather_kernel(const double* __restrict__ xyz, const double* __restrict__ nrm_in, size_t n,
                              const uint32_t* __restrict__ perm, double cx, double cy, double cz,
                              float4* __restrict__ p32, double4* __restrict__ p64,
                              double4* __restrict__ nrm_out, float* __restrict__ bmax_part) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float m = 0.f;
  if (j < n) {
    const uint32_t src = perm[j];
    const double x = xyz[3 * (size_t)src], y = xyz[3 * (size_t)src + 1], z = xyz[3 * (size_t)src + 2];
    p64[j] = make_double4(x, y, z, 0.0);
    const float fx = (float)(x - cx), fy = (float)(y - cy), fz = (float)(z - cz);
    p32[j] = make_float4(fx, fy, fz, __uint_as_float(src));
    m = fmaxf(fmaxf(fabsf(fx), fabsf(fy)), fabsf(fz));
    if (nrm_in)
      nrm_out[j] = make_double4(nrm_in[3 * (size_t)src], nrm_in[3 * (size_t)src + 1], nrm_in[3 * (size_t)src + 2], 0.0);
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s[w]);
    bmax_part[blockIdx.x] = m;
  }
}

}  // namespace b200
