// Microbenchmark of the serial part of one ICP iteration (one thread, like the last block of icp_iter_kernel):
// cycles of solve_quat and its pieces.  nvcc -arch=sm_100a -O3 -I3dtk_b200/csrc tools/bench_src/solve_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "solve.h"
using namespace b200;
__global__ void k(const double* mom_in, const double* o_in, double* out, long long* cyc) {
  __shared__ double mom[NS_MAX], o[3];
  if (threadIdx.x < NS_P2P) mom[threadIdx.x] = mom_in[threadIdx.x];
  if (threadIdx.x < 3) o[threadIdx.x] = o_in[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    double alignxf[16];
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      double ret = solve_any(1, mom, o, 0, alignxf);
      long long t1 = clock64();
      double T[16], X[16];
      m4_identity(X);
      m4_mul(alignxf, X, T);
      long long t2 = clock64();
      cyc[3 * rep] = t1 - t0; cyc[3 * rep + 1] = t2 - t1;
      out[0] = ret;
      for (int i = 0; i < 16; ++i) out[1 + i] = T[i];
      // pieces: eigenvector only
      double Q[4][4];
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Q[i][j] = mom[MP_DM + (i + j) % 9] * (i == j ? 2.0 : 0.3) / mom[0];
      for (int i = 0; i < 4; ++i) for (int j = 0; j < i; ++j) Q[i][j] = Q[j][i];
      double v[4];
      long long t3 = clock64();
      sym4_max_eigvec(Q, v);
      long long t4 = clock64();
      cyc[3 * rep + 2] = t4 - t3;
      out[17] = v[0];
    }
  }
}
int main() {
  double mom[NS_MAX] = {0}, o[3] = {10, 20, 30};
  srand(1);
  struct A { double* p; double& operator[](int k) { return p[k]; } } acc{mom};
  for (int i = 0; i < 100000; ++i) {
    double p1[3], p2[3];
    for (int k = 0; k < 3; ++k) { p1[k] = 2000.0 * rand() / RAND_MAX - 1000.0; }
    // small rotation about z + translation + noise
    double c = cos(0.01), s = sin(0.01);
    p2[0] = c * p1[0] - s * p1[1] + 3.0 + 0.5 * rand() / RAND_MAX;
    p2[1] = s * p1[0] + c * p1[1] - 2.0 + 0.5 * rand() / RAND_MAX;
    p2[2] = p1[2] + 1.0 + 0.5 * rand() / RAND_MAX;
    accumulate_p2p(acc, p1, p2, o);
  }
  double *dm, *dout, *dov; long long* dc;
  cudaMalloc(&dm, sizeof mom); cudaMalloc(&dov, sizeof o); cudaMalloc(&dout, 32 * 8); cudaMalloc(&dc, 16 * 8);
  cudaMemcpy(dm, mom, sizeof mom, cudaMemcpyHostToDevice); cudaMemcpy(dov, o, sizeof o, cudaMemcpyHostToDevice);
  k<<<1, 64>>>(dm, dov, dout, dc);
  long long cyc[9]; double out[18];
  cudaMemcpy(cyc, dc, sizeof cyc, cudaMemcpyDeviceToHost);
  printf("first launch : solve_any %lld cycles (rep 0), %lld (rep 1)\n", cyc[0], cyc[3]);
  // evict the L2 (code included) by streaming 512 MB, then run the same kernel again: rep 0 now fetches its code from DRAM
  char* big; cudaMalloc(&big, 512u << 20);
  for (int t = 0; t < 3; ++t) {
    cudaMemset(big, t, 512u << 20);
    k<<<1, 64>>>(dm, dov, dout, dc);
    cudaMemcpy(cyc, dc, sizeof cyc, cudaMemcpyDeviceToHost);
    printf("after L2 flush: solve_any %lld cycles (rep 0), %lld (rep 1)\n", cyc[0], cyc[3]);
  }
  k<<<1, 64>>>(dm, dov, dout, dc);
  cudaMemcpy(cyc, dc, sizeof cyc, cudaMemcpyDeviceToHost); cudaMemcpy(out, dout, sizeof out, cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  for (int r = 0; r < 3; ++r) printf("rep %d: solve_any(QUAT) %lld cycles, m4_mul %lld, sym4_max_eigvec %lld\n", r, cyc[3 * r], cyc[3 * r + 1], cyc[3 * r + 2]);
  printf("ret %.12f  t = %.6f %.6f %.6f\n", out[0], out[13], out[14], out[15]);
  return 0;
}
