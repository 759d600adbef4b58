// synth_scene.h -- the synthetic scene generator of SURVEY.md section 8d (header-only, no dependencies).
//
// Test / bench INPUT data, not part of the ICP path: a 2000 x 300 x 1000 cm room (inner faces) with four interior wall
// panels and twenty boxes placed from `geom_seed`, sampled area-proportionally from `sample_seed` with additive
// N(0, sigma^2) noise per coordinate; fp64 AoS.  Fully specified (std::mt19937_64 + Box-Muller), so every
// translation unit that includes it produces the same arrays: the product library exports it as
// b200icp_synth_scene, and oracle/scene_gen.cpp builds it into a library of its own so that the reference arm of
// bench.py can make its inputs without loading the product.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <random>
#include <vector>

namespace b200 {
namespace scene_detail {
struct Rect { double o[3], u[3], v[3], area; };

inline double u01(std::mt19937_64& g) { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }

void add_rect(std::vector<Rect>& rs, double ox, double oy, double oz, double ux, double uy, double uz,
              double vx, double vy, double vz) {
  Rect r = {{ox, oy, oz}, {ux, uy, uz}, {vx, vy, vz}, 0.0};
  const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
  r.area = std::sqrt(cx * cx + cy * cy + cz * cz);
  rs.push_back(r);
}

void add_box(std::vector<Rect>& rs, double x0, double y0, double z0, double sx, double sy, double sz) {
  add_rect(rs, x0, y0 + sy, z0, sx, 0, 0, 0, 0, sz);       // top
  add_rect(rs, x0, y0, z0, sx, 0, 0, 0, sy, 0);            // z = z0
  add_rect(rs, x0, y0, z0 + sz, sx, 0, 0, 0, sy, 0);       // z = z0+sz
  add_rect(rs, x0, y0, z0, 0, 0, sz, 0, sy, 0);            // x = x0
  add_rect(rs, x0 + sx, y0, z0, 0, 0, sz, 0, sy, 0);       // x = x0+sx
}

}  // namespace scene_detail

inline int synth_scene(uint64_t geom_seed, uint64_t sample_seed, size_t n, double noise_sigma, double* xyz_out) {
  using namespace scene_detail;
  if (!xyz_out && n) return -1;
  std::vector<Rect> rs;
  // room: x in [-1000,1000], y (up) in [0,300], z in [-500,500]; inner faces
  add_rect(rs, -1000, 0, -500, 2000, 0, 0, 0, 0, 1000);    // floor
  add_rect(rs, -1000, 300, -500, 2000, 0, 0, 0, 0, 1000);  // ceiling
  add_rect(rs, -1000, 0, -500, 2000, 0, 0, 0, 300, 0);     // z = -500
  add_rect(rs, -1000, 0, 500, 2000, 0, 0, 0, 300, 0);      // z = +500
  add_rect(rs, -1000, 0, -500, 0, 0, 1000, 0, 300, 0);     // x = -1000
  add_rect(rs, 1000, 0, -500, 0, 0, 1000, 0, 300, 0);      // x = +1000
  std::mt19937_64 gg(geom_seed);
  for (int w = 0; w < 4; ++w) {  // interior wall panels, alternating orientation
    const double len = 250.0 + 350.0 * u01(gg), hgt = 200.0 + 100.0 * u01(gg);
    const double px = -850.0 + 1700.0 * u01(gg), pz = -420.0 + 840.0 * u01(gg);
    if (w & 1) add_rect(rs, px, 0, std::min(pz, 500.0 - len), 0, 0, len, 0, hgt, 0);
    else add_rect(rs, std::min(px, 1000.0 - len), 0, pz, len, 0, 0, 0, hgt, 0);
  }
  for (int b = 0; b < 20; ++b) {  // axis-aligned boxes standing on the floor
    const double sx = 50.0 + 100.0 * u01(gg), sy = 50.0 + 100.0 * u01(gg), sz = 50.0 + 100.0 * u01(gg);
    const double x0 = -950.0 + (1900.0 - sx) * u01(gg), z0 = -470.0 + (940.0 - sz) * u01(gg);
    add_box(rs, x0, 0.0, z0, sx, sy, sz);
  }
  std::vector<double> cum(rs.size());
  double total = 0.0;
  for (size_t i = 0; i < rs.size(); ++i) { total += rs[i].area; cum[i] = total; }
  std::mt19937_64 gs(sample_seed);
  const double two_pi = 6.283185307179586476925286766559;
  for (size_t i = 0; i < n; ++i) {
    const double pick = u01(gs) * total;
    size_t lo = 0, hi = rs.size() - 1;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (cum[mid] > pick) hi = mid; else lo = mid + 1; }
    const Rect& r = rs[lo];
    const double a = u01(gs), b = u01(gs);
    // Box-Muller, three normals from two pairs (fully specified, independent of libstdc++'s
    // std::normal_distribution)
    const double r1 = std::sqrt(-2.0 * std::log(1.0 - u01(gs))), t1 = two_pi * u01(gs);
    const double r2 = std::sqrt(-2.0 * std::log(1.0 - u01(gs))), t2 = two_pi * u01(gs);
    const double nx = r1 * std::cos(t1), ny = r1 * std::sin(t1), nz = r2 * std::cos(t2);
    xyz_out[3 * i + 0] = r.o[0] + a * r.u[0] + b * r.v[0] + noise_sigma * nx;
    xyz_out[3 * i + 1] = r.o[1] + a * r.u[1] + b * r.v[1] + noise_sigma * ny;
    xyz_out[3 * i + 2] = r.o[2] + a * r.u[2] + b * r.v[2] + noise_sigma * nz;
  }
  return 0;
}


// EulerToMatrix4 (reference include/slam6d/globals.icc:501-531): column-major pose from position + Euler angles
inline void euler_to_matrix4(const double rPos[3], const double rPosTheta[3], double out[16]) {
  const double sx = sin(rPosTheta[0]), cx = cos(rPosTheta[0]);
  const double sy = sin(rPosTheta[1]), cy = cos(rPosTheta[1]);
  const double sz = sin(rPosTheta[2]), cz = cos(rPosTheta[2]);
  out[0] = cy * cz;
  out[1] = sx * sy * cz + cx * sz;
  out[2] = -cx * sy * cz + sx * sz;
  out[3] = 0.0;
  out[4] = -cy * sz;
  out[5] = -sx * sy * sz + cx * cz;
  out[6] = cx * sy * sz + sx * cz;
  out[7] = 0.0;
  out[8] = sy;
  out[9] = -sx * cy;
  out[10] = cx * cy;
  out[11] = 0.0;
  out[12] = rPos[0];
  out[13] = rPos[1];
  out[14] = rPos[2];
  out[15] = 1.0;
}

}  // namespace b200
