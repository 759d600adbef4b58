// icp6d_gpu.h -- reference-side adapter: 3DTK's icp6D with match() running on the B200 engine.
//
// Drop-in for icp6D wherever the reference instantiates it (src/slam6d/slam6D.cc:737,770,810,825;
// graphSlam6D's constructor graphSlam6D.cc:64-65): same constructor arguments, `match` is the virtual
// icp6D::match (include/slam6d/icp6D.h:51-53, src/slam6d/icp6D.cc:104-285).  doICP, metascans, pose extrapolation,
// frames: the unmodified base class keeps driving them (icp6D::doICP calls the virtual match).
//
//   model   PreviousScan's "xyz reduced original" (what its search tree is built over, basicScan.cc:702-709) with
//           Source->dalignxf, or -- for a MetaScan -- the members' current points (kdMeta.cc:34-72); uploaded and
//           binned once per Scan object and kept (a scan's original points never change)
//   data    CurrentScan's current "xyz reduced" (+ "normal reduced" for CLOSEST_PLANE_SIMPLE), uploaded per match
//   loop    b200icp_match: every iteration of the loop on the device in the reference's arithmetic order
//   result  replayed on the host through Scan::transform with the reference's own frame rule (icp6D.cc:258-279):
//           CurrentScan's points, transMat, dalignxf and every scan's frames end up as the CPU loop leaves them
//           (iterations that write no frame are composed into one Scan::transform call)
//
// Minimizers: the algorithm id of the icp6Dminimizer selects the on-device solve (1 QUAT, 2 SVD, 3 ORTHO, 4 DUAL,
// 5 HELIX, 6 APX, 10 NAPX); anything else throws std::runtime_error -- there is no CPU fallback in this class.
// Compile inside a 3DTK tree or against its headers: -I<3dtk>/include -I<this repo>/include.
#pragma once
#include <map>

#include "b200icp.h"
#include "slam6d/icp6D.h"

class icp6D_gpu : public icp6D {
 public:
  icp6D_gpu(icp6Dminimizer* my_icp6Dminimizer, double max_dist_match = 25.0, int max_num_iterations = 50,
            bool quiet = false, bool meta = false, int rnd = 1, bool eP = true, int anim = -1,
            double epsilonICP = 0.0000001, int nns_method = simpleKD, bool cuda_enabled = false,
            bool cad_matching = false, int max_num_metascans = -1, int device = 0);
  virtual ~icp6D_gpu();

  virtual int match(Scan* PreviousScan, Scan* CurrentScan, PairingMode pairing_mode = CLOSEST_POINT);

  // device copies are keyed by Scan*; call when a Scan object is deleted while this matcher lives on
  void forget(Scan* scan);
  // what the last match did on the device
  const b200icp_match_result& last_result() const { return last_; }

 private:
  struct Model { b200icp_scan* scan; const void* data; size_t n; };
  b200icp_scan* model_of(Scan* s);
  int algo_id() const;

  b200icp_ctx* ctx_;
  std::map<Scan*, Model> models_;
  b200icp_match_result last_;
};
