// icp6d_gpu.cc -- see icp6d_gpu.h.  Errors follow the reference's style: std::runtime_error (basicScan.cc:723-726).
#include "icp6d_gpu.h"

#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "slam6d/globals.icc"
#include "slam6d/icp6Dortho.h"
#include "slam6d/metaScan.h"

namespace {
[[noreturn]] void raise(const char* what) {
  throw std::runtime_error(std::string("icp6D_gpu: ") + what + ": " + b200icp_last_error());
}
}  // namespace

icp6D_gpu::icp6D_gpu(icp6Dminimizer* my_icp6Dminimizer, double max_dist_match, int max_num_iterations, bool quiet,
                     bool meta, int rnd, bool eP, int anim, double epsilonICP, int nns_method, bool cuda_enabled,
                     bool cad_matching, int max_num_metascans, int device)
    : icp6D(my_icp6Dminimizer, max_dist_match, max_num_iterations, quiet, meta, rnd, eP, anim, epsilonICP, nns_method,
            cuda_enabled, cad_matching, max_num_metascans),
      ctx_(nullptr) {
  memset(&last_, 0, sizeof last_);
  if (b200icp_create(device, &ctx_) != B200ICP_OK) raise("create");
}

icp6D_gpu::~icp6D_gpu() {
  for (auto& kv : models_) b200icp_scan_destroy(ctx_, kv.second.scan);
  b200icp_destroy(ctx_);
}

void icp6D_gpu::forget(Scan* scan) {
  auto it = models_.find(scan);
  if (it == models_.end()) return;
  b200icp_scan_destroy(ctx_, it->second.scan);
  models_.erase(it);
}

int icp6D_gpu::algo_id() const {
  const int id = my_icp6Dminimizer->getAlgorithmID();
  switch (id) {
    case 1: case 2: case 4: case 5: case 6: case 10: return id;
    case 3:   // icp6D_ORTHO and icp6D_LUMEULER share the id (icp6Dortho.h:34, icp6Dlumeuler.h:33)
      if (dynamic_cast<icp6D_ORTHO*>(my_icp6Dminimizer)) return 3;
      break;
    default: break;
  }
  throw std::runtime_error("icp6D_gpu: minimizer with algorithm id " + std::to_string(id) +
                           " is not on the accelerated path (no CPU fallback in this class)");
}

// device copy of what PreviousScan's search tree is built over: "xyz reduced original" (basicScan.cc:704-709)
b200icp_scan* icp6D_gpu::model_of(Scan* s) {
  DataXYZ xyz(s->get("xyz reduced original"));
  const size_t n = xyz.size();
  if (n == 0) throw std::runtime_error("icp6D_gpu: model scan has no reduced points");
  const void* data = &xyz[0][0];
  auto it = models_.find(s);
  if (it != models_.end() && it->second.data == data && it->second.n == n) return it->second.scan;
  if (it != models_.end()) forget(s);
  b200icp_scan* dev = nullptr;
  if (b200icp_scan_create(ctx_, &xyz[0][0], nullptr, n, 0.0, sqrt(max_dist_match2), &dev) != B200ICP_OK)
    raise("scan_create (model)");
  models_[s] = Model{dev, data, n};
  return dev;
}

int icp6D_gpu::match(Scan* PreviousScan, Scan* CurrentScan, PairingMode pairing_mode) {
  double id[16];
  M4identity(id);
  CurrentScan->transform(id, Scan::ICP, 0);   // write end pose (icp6D.cc:109)
  if (max_num_iterations == 0) return 0;
  if (pairing_mode != CLOSEST_POINT && pairing_mode != CLOSEST_PLANE_SIMPLE)
    throw std::runtime_error("icp6D_gpu: pairing mode not on the accelerated path");
  const int algo = algo_id();
  long time = GetCurrentTimeInMilliSec();

  // ---- model
  b200icp_scan* model = nullptr;
  bool own_model = false;
  MetaScan* meta_prev = dynamic_cast<MetaScan*>(PreviousScan);
  if (meta_prev) {
    // KDtreeMetaManaged searches the members' CURRENT points (kdMeta.cc:34-72): one grid over all of them
    std::vector<double> all;
    for (size_t k = 0; k < meta_prev->size(); ++k) {
      DataXYZ xyz(meta_prev->getScan(k)->get("xyz reduced"));
      if (xyz.size()) all.insert(all.end(), &xyz[0][0], &xyz[0][0] + 3 * xyz.size());
    }
    if (all.empty()) throw std::runtime_error("icp6D_gpu: empty metascan");
    if (b200icp_scan_create(ctx_, all.data(), nullptr, all.size() / 3, 0.0, sqrt(max_dist_match2), &model) != B200ICP_OK)
      raise("scan_create (metascan)");
    own_model = true;
  } else {
    model = model_of(PreviousScan);
    // Source->dalignxf (scan.cc:1240): the model grid stays in the frame it was built in
    if (b200icp_scan_set_pose(model, PreviousScan->get_transMat(), PreviousScan->getDAlign()) != B200ICP_OK)
      raise("scan_set_pose");
  }

  // ---- data: current "xyz reduced" (+ normals), pose = the scan's current transMat, dalignxf restarts at identity
  DataXYZ cur(CurrentScan->get("xyz reduced"));
  const size_t nd = cur.size();
  b200icp_scan* data = nullptr;
  int iter = 0;
  if (nd > 0) {
    const double* nrm = nullptr;
    DataNormal* dn = nullptr;
    if (pairing_mode == CLOSEST_PLANE_SIMPLE) {
      dn = new DataNormal(CurrentScan->get("normal reduced"));
      if (dn->size() != nd) { delete dn; throw std::runtime_error("icp6D_gpu: CLOSEST_PLANE_SIMPLE needs reduced normals"); }
      nrm = &(*dn)[0][0];
    }
    const int rc = b200icp_scan_create(ctx_, &cur[0][0], nrm, nd, 0.0, sqrt(max_dist_match2), &data);
    delete dn;
    if (rc != B200ICP_OK) { if (own_model) b200icp_scan_destroy(ctx_, model); raise("scan_create (data)"); }
    double T0[16];
    memcpy(T0, CurrentScan->get_transMat(), sizeof T0);
    b200icp_scan_set_pose(data, T0, id);

    b200icp_match_params prm;
    memset(&prm, 0, sizeof prm);
    prm.algo = algo;
    prm.pairing_mode = pairing_mode == CLOSEST_PLANE_SIMPLE ? B200ICP_CLOSEST_PLANE_SIMPLE : B200ICP_CLOSEST_POINT;
    prm.max_dist_match = sqrt(max_dist_match2);
    prm.max_num_iterations = max_num_iterations;
    prm.epsilon_icp = epsilonICP;
    prm.rnd = rnd;
    prm.exact = 1;
    if (b200icp_match(ctx_, model, data, &prm, nullptr, nullptr, &last_) != B200ICP_OK) {
      b200icp_scan_destroy(ctx_, data);
      if (own_model) b200icp_scan_destroy(ctx_, model);
      raise("match");
    }
    iter = last_.iterations;
    nr_pointPair = (int)last_.npairs_last;

    // ---- replay on the host scan with the frame rule of icp6D.cc:258-264: the transform of an iteration that writes
    // a frame (iteration 0 and every anim'th one) is applied by its own Scan::transform(.., ICP, 0); the transforms
    // of the iterations in between write no frame (islum = -1 in the reference), so they are composed and applied by
    // ONE Scan::transform(.., ICP, -1) -- same frames, same final transMat / dalignxf / points up to the rounding of
    // the composition, and the host walks the scan's points 3 times per match instead of once per iteration
    const int ran = last_.iterations_run;
    std::vector<double> poses(16 * (size_t)(ran > 0 ? ran : 1));
    if (ran > 0 && b200icp_last_poses(ctx_, ran, poses.data()) < ran) raise("last_poses");
    double prev[16], inv[16], alignxf[16], pending[16], tmp[16];
    bool has_pending = false;
    M4identity(pending);
    memcpy(prev, T0, sizeof prev);
    for (int k = 0; k < ran; ++k) {
      M4inv(prev, inv);
      MMult(&poses[16 * (size_t)k], inv, alignxf);            // alignxf_k = T_k * T_{k-1}^-1
      const bool frame = (k == 0 && anim != -2) || (anim > 0 && k % anim == 0);
      if (frame) {
        if (has_pending) { CurrentScan->transform(pending, Scan::ICP, -1); M4identity(pending); has_pending = false; }
        CurrentScan->transform(alignxf, Scan::ICP, 0);
      } else {
        MMult(alignxf, pending, tmp);
        memcpy(pending, tmp, sizeof pending);
        has_pending = true;
      }
      memcpy(prev, &poses[16 * (size_t)k], sizeof prev);
    }
    if (has_pending) CurrentScan->transform(pending, Scan::ICP, -1);
    // the loop ends through its convergence test / iteration cap with an end-pose frame, or leaves early
    // ("do we have enough point pairs?", :233-241) without one
    if (ran > 0 && ran == iter + 1) CurrentScan->transform(id, Scan::ICP, anim == -2 ? -1 : 0);
    b200icp_scan_destroy(ctx_, data);
  }
  if (own_model) b200icp_scan_destroy(ctx_, model);

  long endtime = GetCurrentTimeInMilliSec() - time;
  cout << "TIME  " << endtime << "   ITER " << iter << endl;   // as icp6D::match (icp6D.cc:282-283)
  return iter;
}
