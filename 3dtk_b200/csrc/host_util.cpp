// host_util.cpp -- host-only entry points of the C ABI: pair-list Align (moments + solve.h), the
// synthetic scene generator of SURVEY.md section 8d, and the 4x4 helpers bindings need to build poses
// the way the reference does.  No CUDA calls here; usable (and tested) on a machine without a GPU.
#include "../../include/b200icp.h"

#include <cmath>
#include <cstring>
#include <random>
#include <vector>

#include "solve.h"
#include "synth_scene.h"

using namespace b200;

extern "C" {

int b200icp_align_pairs(int algo, size_t n, const double* p1, const double* p2, const double* nrm,
                        const double centroid_m[3], const double centroid_d[3], double alignxf[16],
                        double* rms_out) {
  if (!p1 || !p2 || !alignxf) return B200ICP_EINVAL;
  if (algo != 1 && algo != 2 && algo != 3 && algo != 4 && algo != 5 && algo != 6 && algo != 10) return B200ICP_EINVAL;
  if (algo == 10 && !nrm) return B200ICP_EINVAL;
  if (n == 0) return B200ICP_EEMPTY;
  // Shift origin: the data-side centroid the caller already holds (any point near the cloud works).
  double o[3] = {0, 0, 0};
  if (centroid_d) memcpy(o, centroid_d, sizeof o);
  double mom[NS_MAX];
  for (int k = 0; k < NS_MAX; ++k) mom[k] = 0.0;
  if (algo == 10)
    for (size_t i = 0; i < n; ++i) accumulate_napx(mom, p1 + 3 * i, p2 + 3 * i, nrm + 3 * i, o);
  else
    for (size_t i = 0; i < n; ++i) accumulate_p2p(mom, p1 + 3 * i, p2 + 3 * i, o);
  (void)centroid_m;  // the centroids are recomputed from the same pairs (they are sums over them)
  double r = solve_any(algo, mom, o, 0, alignxf);
  if (rms_out) *rms_out = r;
  return B200ICP_OK;
}

void b200icp_euler_to_matrix4(const double rPos[3], const double rPosTheta[3], double out[16]) {
  b200::euler_to_matrix4(rPos, rPosTheta, out);
}

int b200icp_m4inv(const double in[16], double out[16]) { return m4_inverse(in, out); }

void b200icp_mmult(const double a[16], const double b[16], double out[16]) {
  double t[16];
  m4_mul(a, b, t);
  memcpy(out, t, sizeof t);
}

void b200icp_transform_points(const double xf[16], double* xyz, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    double q[3];
    xf_point(xf, xyz + 3 * i, q);
    xyz[3 * i] = q[0]; xyz[3 * i + 1] = q[1]; xyz[3 * i + 2] = q[2];
  }
}

int b200icp_synth_scene(uint64_t geom_seed, uint64_t sample_seed, size_t n, double noise_sigma,
                        double* xyz_out) {
  return b200::synth_scene(geom_seed, sample_seed, n, noise_sigma, xyz_out) == 0 ? B200ICP_OK : B200ICP_EINVAL;
}

}  // extern "C"
