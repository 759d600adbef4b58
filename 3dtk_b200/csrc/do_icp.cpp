// do_icp.cpp -- icp6D::doICP (reference src/slam6d/icp6D.cc:374-437): the sequential driver around
// b200icp_match, with Scan::mergeCoordinatesWithRoboterPosition (scan.cc:826-833) and the metascan option
// (MetaScan, metaScan.cc; max_num_metascans window icp6D.cc:424-431).  Host logic over the public C ABI only.
#include "../../include/b200icp.h"

#include <cstring>
#include <vector>

extern "C" int b200icp_set_error_(int code, const char* msg);

extern "C" int b200icp_do_icp(b200icp_ctx* ctx, b200icp_scan* const* scans, int n_scans,
                              const b200icp_match_params* params, int extrapolate_pose, int meta,
                              int max_num_metascans, const double* transMatOrg, int* iterations_out,
                              b200icp_frames* frames) {
  if (!ctx || !scans || !params || n_scans < 0) return b200icp_set_error_(B200ICP_EINVAL, "do_icp: bad argument");
  std::vector<double> org((size_t)16 * (n_scans > 0 ? n_scans : 1));
  for (int i = 0; i < n_scans; ++i) {
    if (!scans[i]) return b200icp_set_error_(B200ICP_EINVAL, "do_icp: NULL scan");
    if (transMatOrg) memcpy(&org[(size_t)16 * i], transMatOrg + (size_t)16 * i, 16 * sizeof(double));
    else b200icp_scan_get_pose(scans[i], &org[(size_t)16 * i], nullptr);
  }
  // Scan::transform(.., ICP, 0) of scan i (scan.cc:955-983): a frame for EVERY scan, from their current poses;
  // pose_i (may be NULL) overrides scan i's transMat (a pose of the match log, not the final one)
  std::vector<double> all((size_t)16 * (n_scans > 0 ? n_scans : 1));
  auto push_frames = [&](int i, const double* pose_i) {
    if (!frames) return;
    for (int k = 0; k < n_scans; ++k) b200icp_scan_get_pose(scans[k], &all[(size_t)16 * k], nullptr);
    if (pose_i) memcpy(&all[(size_t)16 * i], pose_i, 16 * sizeof(double));
    b200icp_frames_transform(frames, i, all.data(), B200ICP_FRAME_ICP, 0);
  };
  std::vector<const b200icp_scan*> meta_scans;
  b200icp_scan* metascan = nullptr;
  int rc = B200ICP_OK;
  for (int i = 0; i < n_scans && rc == B200ICP_OK; ++i) {
    if (iterations_out) iterations_out[i] = 0;
    if (i > 0) {
      if (extrapolate_pose) {
        double prevT[16], inv[16], delta[16];
        b200icp_scan_get_pose(scans[i - 1], prevT, nullptr);
        if (!b200icp_m4inv(&org[(size_t)16 * (i - 1)], inv)) {
          rc = b200icp_set_error_(B200ICP_ESTATE, "do_icp: singular transMatOrg");
          break;
        }
        b200icp_mmult(prevT, inv, delta);
        b200icp_scan_transform(scans[i], delta);
      }
      push_frames(i, nullptr);                       // icp6D.cc:109  transform(id, ICP, 0)
      b200icp_match_result res;
      rc = b200icp_match(ctx, meta ? metascan : scans[i - 1], scans[i], params, nullptr, nullptr, &res);
      if (rc != B200ICP_OK) break;
      if (iterations_out) iterations_out[i] = res.iterations;
      if (frames && res.iterations_run > 0) {
        double first[16];
        b200icp_last_poses(ctx, 1, first);
        push_frames(i, first);                       // icp6D.cc:258-260  iteration 0 (anim = -1)
        // the end pose is written when the loop ends by convergence or by the iteration limit (icp6D.cc:266-279),
        // not when it stops for lack of pairs (:231-236)
        if (res.iterations_run == res.iterations + 1) push_frames(i, nullptr);
      }
    }
    if (meta) {
      if (metascan) { b200icp_scan_destroy(ctx, metascan); metascan = nullptr; }
      if (i != n_scans - 1) {
        meta_scans.push_back(scans[i]);
        if (max_num_metascans > 0)
          while ((int)meta_scans.size() > max_num_metascans) meta_scans.erase(meta_scans.begin());
        rc = b200icp_metascan_create(ctx, meta_scans.data(), (int)meta_scans.size(), 0.0, params->max_dist_match,
                                     &metascan);
      }
    }
  }
  if (metascan) b200icp_scan_destroy(ctx, metascan);
  return rc;
}
