// TEST / BENCH INFRASTRUCTURE ONLY -- builds oracle/_build/libscenegen.so.
//
// The synthetic scan pair of SURVEY.md section 8d for processes that must not load the product library (the
// reference arm of bench.py): same generator source as b200icp_synth_scene (3dtk_b200/csrc/synth_scene.h, header
// only), same pose arithmetic (EulerToMatrix4, M4inv by cofactors, transform3 -- solve.h restates
// include/slam6d/globals.icc:501-531,761-781,1454-1490), so both arms of the bench see bit-identical arrays.
#include <cstring>
#include "../3dtk_b200/csrc/solve.h"
#include "../3dtk_b200/csrc/synth_scene.h"

extern "C" {

int scenegen_scene(uint64_t geom_seed, uint64_t sample_seed, size_t n, double sigma, double* xyz_out) {
  return b200::synth_scene(geom_seed, sample_seed, n, sigma, xyz_out);
}

// model = scene(geom, seed_model) at the identity pose; data = scene(geom, seed_data) moved by the INVERSE of the
// pose (rPos [cm], rPosTheta [deg]); P_out (may be NULL) = the pose, i.e. the transform ICP should recover.
int scenegen_pair(uint64_t geom_seed, uint64_t seed_model, uint64_t seed_data, size_t n, double sigma,
                  const double rPos[3], const double rPosThetaDeg[3], double* model_out, double* data_out,
                  double* P_out) {
  if (b200::synth_scene(geom_seed, seed_model, n, sigma, model_out) != 0) return -1;
  if (b200::synth_scene(geom_seed, seed_data, n, sigma, data_out) != 0) return -1;
  const double d2r = 0.017453292519943295769236907684886;   // numpy.deg2rad's constant (pi / 180)
  const double th[3] = {rPosThetaDeg[0] * d2r, rPosThetaDeg[1] * d2r, rPosThetaDeg[2] * d2r};
  double P[16], Pinv[16];
  b200::euler_to_matrix4(rPos, th, P);
  if (!b200::m4_inverse(P, Pinv)) return -2;
  for (size_t i = 0; i < n; ++i) {
    double q[3];
    b200::xf_point(Pinv, data_out + 3 * i, q);
    data_out[3 * i] = q[0]; data_out[3 * i + 1] = q[1]; data_out[3 * i + 2] = q[2];
  }
  if (P_out) memcpy(P_out, P, sizeof P);
  return 0;
}

}  // extern "C"
