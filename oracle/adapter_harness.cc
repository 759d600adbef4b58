// TEST INFRASTRUCTURE ONLY.  extern "C" shell around the reference-side adapter
// (3dtk_b200/host/gpu_search_tree.{h,cc}) so tests can drive it with ctypes.  Linked with the reference's
// own compiled searchTree.o: `use_base_loop` runs the UNMODIFIED SearchTree::getPtPairs batch loop
// (src/slam6d/searchTree.cc:92-188) on top of the adapter's virtual FindClosest; otherwise the adapter's
// batched override runs.  Either way the pairs must equal what the reference's KDtree produces.
#include <cstring>
#include <vector>

#include "gpu_search_tree.h"

namespace {
struct Holder {
  std::vector<double> xyz;
  std::vector<double*> rows;
  GpuSearchTree* tree = nullptr;
};
}  // namespace

extern "C" {

void* adp_tree_create(const double* xyz, long n, double max_dist_hint, char* err, int errlen) {
  Holder* h = new Holder();
  h->xyz.assign(xyz, xyz + 3 * n);
  h->rows.resize(n);
  for (long i = 0; i < n; ++i) h->rows[i] = &h->xyz[3 * i];
  try {
    h->tree = new GpuSearchTree(h->rows.data(), (int)n, max_dist_hint);
  } catch (const std::exception& e) {
    if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    delete h;
    return nullptr;
  }
  return h;
}

void adp_tree_free(void* p) {
  Holder* h = (Holder*)p;
  if (!h) return;
  delete h->tree;
  delete h;
}

long adp_find_closest(void* p, const double* q, double maxdist2, int thread_num) {
  Holder* h = (Holder*)p;
  double qq[3] = {q[0], q[1], q[2]};
  SearchTree* t = h->tree;   // through the base-class interface, as Scan code would call it
  double* c = t->FindClosest(qq, maxdist2, thread_num);
  return c ? (long)((c - h->xyz.data()) / 3) : -1;
}

long adp_get_pt_pairs(void* p, const double* source_alignxf, const double* data_xyz, const double* data_nrm,
                      long start, long end, int thread_num, int rnd, double maxdist2, int pairing_mode,
                      int use_base_loop, double* p1, double* p2, double* nrm, double* sum, double* cm,
                      double* cd) {
  Holder* h = (Holder*)p;
  std::vector<PtPair> pairs;
  double xf[16];
  memcpy(xf, source_alignxf, sizeof xf);
  DataXYZ xyz_r(DataPointer((unsigned char*)data_xyz, 0));
  DataNormal nrm_r(DataPointer((unsigned char*)data_nrm, 0));
  SearchTree* t = h->tree;
  if (use_base_loop)
    t->SearchTree::getPtPairs(&pairs, xf, xyz_r, nrm_r, (unsigned)start, (unsigned)end, thread_num, rnd,
                              maxdist2, *sum, cm, cd, (PairingMode)pairing_mode);
  else
    t->getPtPairs(&pairs, xf, xyz_r, nrm_r, (unsigned)start, (unsigned)end, thread_num, rnd, maxdist2, *sum,
                  cm, cd, (PairingMode)pairing_mode);
  for (size_t i = 0; i < pairs.size(); ++i) {
    p1[3 * i] = pairs[i].p1.x; p1[3 * i + 1] = pairs[i].p1.y; p1[3 * i + 2] = pairs[i].p1.z;
    p2[3 * i] = pairs[i].p2.x; p2[3 * i + 1] = pairs[i].p2.y; p2[3 * i + 2] = pairs[i].p2.z;
    if (nrm) { nrm[3 * i] = pairs[i].p2.nx; nrm[3 * i + 1] = pairs[i].p2.ny; nrm[3 * i + 2] = pairs[i].p2.nz; }
  }
  return (long)pairs.size();
}

}  // extern "C"
