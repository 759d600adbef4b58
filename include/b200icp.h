/* b200icp.h -- C ABI of the B200-native ICP correspondence-and-alignment engine.
 *
 * This is the drop-in boundary for ONE path of JMUWRobotics/3DTK: per-iteration nearest-neighbour
 * correspondence search + distance rejection + covariance / residual accumulation + 6-DoF solve.
 * Every entry point names the reference interface it replaces (paths relative to the 3DTK tree,
 * commit 5b570686).  The reference-side C++ adapter (class GpuSearchTree : public SearchTree,
 * class icp6D_gpu : public icp6D) that binds these symbols is shown in INTEGRATION.md and shipped
 * in 3dtk_b200/host/.
 *
 * Conventions (identical to the reference, include/slam6d/globals.icc:298-321, :1454-1490):
 *   - 4x4 matrices are double[16] in column-major OpenGL order, translation in [12..14];
 *   - points / normals are fp64 AoS, 3 doubles per element (DataXYZ / DataNormal);
 *   - all functions return 0 on success, a negative B200ICP_E* code otherwise, never throw;
 *     b200icp_last_error() gives the message of the calling thread's last failure;
 *   - indices returned refer to the caller's original point order (row of the uploaded array).
 *   - The library needs a CUDA device (sm_100a).  There is NO CPU fallback: without a usable GPU
 *     b200icp_create() fails with B200ICP_ENODEV and nothing else can be called.
 */
#ifndef B200ICP_H_
#define B200ICP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ICP_OK 0
#define B200ICP_EINVAL (-1)  /* bad argument                                             */
#define B200ICP_ENODEV (-2)  /* no CUDA device / wrong architecture                       */
#define B200ICP_ECUDA (-3)   /* CUDA runtime error (message in b200icp_last_error)        */
#define B200ICP_ENOMEM (-4)  /* host or device allocation failed                          */
#define B200ICP_EEMPTY (-5)  /* zero points (reference: kdTreeImpl.h:86-88 throws)        */
#define B200ICP_ESTATE (-6)  /* object used in a state that does not allow the call       */

/* icp6Dminimizer::getAlgorithmID() values (include/slam6d/icp6D{quat,svd,apx,napx}.h) */
#define B200ICP_ALGO_QUAT 1
#define B200ICP_ALGO_SVD 2
#define B200ICP_ALGO_ORTHO 3  /* icp6D_ORTHO (src/slam6d/icp6Dortho.cc), same pair moments as QUAT/SVD */
#define B200ICP_ALGO_DUAL 4   /* icp6D_DUAL  (src/slam6d/icp6Ddual.cc),  same pair moments             */
#define B200ICP_ALGO_HELIX 5  /* icp6D_HELIX (src/slam6d/icp6Dhelix.cc), same pair moments             */
#define B200ICP_ALGO_APX 6
#define B200ICP_ALGO_NAPX 10

/* PairingMode (include/slam6d/pairingMode.h:4-8); mode 1 (along-normal) is not on the path */
#define B200ICP_CLOSEST_POINT 0
#define B200ICP_CLOSEST_PLANE_SIMPLE 2

typedef struct b200icp_ctx b200icp_ctx;   /* one per host thread / CUDA stream            */
typedef struct b200icp_scan b200icp_scan; /* device-resident scan: points + search grid   */
typedef struct b200icp_frames b200icp_frames;   /* per-scan frame lists, see "scan files and frames" */

/* ---- context ------------------------------------------------------------------------------ */
int b200icp_create(int device, b200icp_ctx** out);
void b200icp_destroy(b200icp_ctx* ctx);
/* Run all subsequent work of this context on `cuda_stream` (a cudaStream_t; NULL = own stream). */
int b200icp_set_stream(b200icp_ctx* ctx, void* cuda_stream);
int b200icp_synchronize(b200icp_ctx* ctx);
const char* b200icp_last_error(void);
/* Version / build info string ("b200icp <ver> sm_100a ..."). */
const char* b200icp_version(void);

/* ---- scans ---------------------------------------------------------------------------------
 * Replaces: Scan::createSearchTree / BasicScan::createSearchTreePrivate (src/slam6d/scan.cc:285-306,
 * basicScan.cc:702-728) + KDtree::KDtree (src/slam6d/kd.cc:46-49).  `xyz` is the scan's
 * "xyz reduced original" array (points in the frame the tree is built in); `normals` (may be NULL)
 * its "normal reduced".  The arrays are copied; the caller keeps ownership.
 * cell_edge <= 0 picks the grid cell edge from the point density; max_dist_hint (> 0) is the search
 * radius the scan will mostly be queried with (icp6D's max_dist_match), used only for that choice.
 * A scan can act as model (Source) and as data (Target) of a match.
 * transMat / dalignxf start as identity (basicScan.cc:293); see b200icp_scan_set_pose. */
int b200icp_scan_create(b200icp_ctx* ctx, const double* xyz, const double* normals, size_t n,
                        double cell_edge, double max_dist_hint, b200icp_scan** out);
/* Same, but xyz / normals are DEVICE pointers (fp64 AoS) already resident in HBM. */
int b200icp_scan_create_device(b200icp_ctx* ctx, const double* d_xyz, const double* d_normals,
                               size_t n, double cell_edge, double max_dist_hint,
                               b200icp_scan** out);
void b200icp_scan_destroy(b200icp_ctx* ctx, b200icp_scan* scan);
size_t b200icp_scan_size(const b200icp_scan* scan);
/* grid facts: dims[3], cell edge, number of cells, occupied cells */
int b200icp_scan_grid_info(const b200icp_scan* scan, int dims[3], double* cell_edge,
                           uint64_t* n_cells, uint64_t* n_occupied);
/* Scan::transMat / Scan::dalignxf (scan.cc:878-898).  Either pointer may be NULL. */
int b200icp_scan_get_pose(const b200icp_scan* scan, double transMat[16], double dalignxf[16]);
int b200icp_scan_set_pose(b200icp_scan* scan, const double transMat[16], const double dalignxf[16]);
/* Scan::transform's bookkeeping (scan.cc:851-898) without touching points: transMat <- alignxf*transMat,
 * dalignxf <- alignxf*dalignxf, normal map by the transform3normal rule.  The kernels apply dalignxf on load. */
int b200icp_scan_transform(b200icp_scan* scan, const double alignxf[16]);
/* MetaScan + KDtreeMetaManaged (src/slam6d/metaScan.cc:27-69, kdMeta.cc:34-72): ONE search structure over the
 * CURRENT positions ("xyz reduced") of all member scans, member after member in their original row order,
 * with identity pose.  Built on the device (export of every member through its dalignxf, then the usual
 * grid build); the members are not modified and may be destroyed afterwards. */
int b200icp_metascan_create(b200icp_ctx* ctx, const b200icp_scan* const* scans, int n_scans, double cell_edge,
                            double max_dist_hint, b200icp_scan** out);
/* Current "xyz reduced" (= dalignxf * original) and "normal reduced", original row order;
 * what Scan::transformReduced (scan.cc:851-875) leaves in the arrays.  nrm_out may be NULL. */
int b200icp_scan_download(b200icp_ctx* ctx, const b200icp_scan* scan, double* xyz_out,
                          double* nrm_out);

/* ---- API-compatible search path --------------------------------------------------------------
 * Replaces: KDtree::FindClosest (src/slam6d/kd.cc:78-87; include/slam6d/searchTree.h:81).
 * Exact fp64 nearest neighbour with the k-d tree's strict rule d^2 < maxdist2
 * (kdTreeImpl.h:353; testing/kdtree/kdtree.cc:20-35).  *idx_out = row of the model point, or -1.
 * Exact-distance ties resolve to the lowest row (the k-d tree's choice depends on build order). */
int b200icp_find_closest(b200icp_ctx* ctx, const b200icp_scan* model, const double p[3],
                         double maxdist2, int64_t* idx_out);

/* Replaces: the batch loop SearchTree::getPtPairs (src/slam6d/searchTree.cc:92-188) minus PtPair
 * materialisation.  For i in [0,n): t = q_xyz[i]; s = inv(source_alignxf) * t; NN of s in the model
 * grid within maxdist2; idx_out[i] = model row or -1; d2_out[i] (optional) = squared distance in the
 * tree frame.  sums_out (optional, 8 doubles, ASSIGNED): {npairs, sum |p1-p2|^2, centroid_m[3] (sum,
 * not divided), centroid_d[3] (sum)} exactly as getPtPairs accumulates them, including the
 * CLOSEST_PLANE_SIMPLE projection (searchTree.cc:149-162) when pairing_mode == 2 (q_nrm required).
 * All pointers are HOST pointers; the H2D / D2H copies are part of the call. */
int b200icp_nn_batch(b200icp_ctx* ctx, const b200icp_scan* model, const double* q_xyz,
                     const double* q_nrm, size_t n, const double source_alignxf[16],
                     double maxdist2, int pairing_mode, int32_t* idx_out, double* d2_out,
                     double sums_out[8]);
/* Same with DEVICE pointers (q_xyz, q_nrm, idx_out, d2_out on device; sums_out on host). */
int b200icp_nn_batch_device(b200icp_ctx* ctx, const b200icp_scan* model, const double* d_q_xyz,
                            const double* d_q_nrm, size_t n, const double source_alignxf[16],
                            double maxdist2, int pairing_mode, int32_t* d_idx_out,
                            double* d_d2_out, double sums_out[8]);

/* ---- 6-DoF solve from pair moments (host; also what the device solve kernel runs) -------------
 * Replaces: icp6D_QUAT/SVD/APX/NAPX::Align (src/slam6d/icp6Dquat.cc:38-144, icp6Dsvd.cc:38-158,
 * icp6Dapx.cc:35-133, icp6Dnapx.cc:34-149) operating on explicit pairs: p1 = model-side point,
 * p2 = data-side point, nrm = per-pair unit normal (NAPX only, else NULL).  centroid_m/centroid_d
 * as passed by icp6D::match.  Returns the RMS the reference returns through *rms_out
 * (-1.0 when the Cholesky factorisation fails, icp6Dapx.cc:97-100).  Pure host arithmetic over
 * the moment sums -- used by the adapter's Align override and by CPU-side tests. */
int b200icp_align_pairs(int algo, size_t n, const double* p1, const double* p2, const double* nrm,
                        const double centroid_m[3], const double centroid_d[3],
                        double alignxf[16], double* rms_out);

/* ---- fused match -------------------------------------------------------------------------------
 * Replaces: icp6D::match (src/slam6d/icp6D.cc:104-285) -- Scan::getPtPairs + Align + Scan::transform
 * per iteration, convergence test icp6D.cc:266-268 -- entirely on the device. */
typedef struct b200icp_match_params {
  int algo;               /* B200ICP_ALGO_*                                                  */
  int pairing_mode;       /* B200ICP_CLOSEST_POINT | B200ICP_CLOSEST_PLANE_SIMPLE            */
  double max_dist_match;  /* NOT squared (icp6D ctor squares it, icp6D.cc:80)                */
  int max_num_iterations; /* icp6D::max_num_iterations                                       */
  double epsilon_icp;     /* icp6D::epsilonICP                                               */
  int rnd;                /* <=1: every point; >1: deterministic 1/rnd subsample             */
  int exact;              /* 1: fp32 filter + fp64 verified NN (default); 0: fp32 decisions  */
  int profile;            /* 1: record per-launch CUDA-event timings (b200icp_match_profile) */
  int napx_weighted;      /* 0: B as shipped (icp6Dnapx.cc:69-74); 1: least-squares B += d*[c;n] */
  int sharded;            /* 1: `data` holds only this rank's slice of the scan; moments are summed over the
                             connected ranks inside the kernel (b200icp_comm_*); all ranks call collectively */
  int reserved[2];
} b200icp_match_params;

typedef struct b200icp_match_result {
  int iterations;       /* value icp6D::match returns (loop index at exit)                   */
  int iterations_run;   /* iterations that produced a transform (entries in rms / npairs)    */
  double rms_last;
  uint64_t npairs_last;
  uint64_t queries;     /* data points searched per iteration                                */
  double nn_kernel_ms;  /* profile=1: mean device time of one correspondence kernel launch   */
  double solve_kernel_ms;
  uint32_t kernel_launches; /* kernels launched by this call                                 */
  uint32_t stage2_queries_last; /* queries that needed the wide (ring) search, last iteration */
} b200icp_match_result;

/* model->dalignxf is used as Source->dalignxf; data's transMat / dalignxf are updated in place
 * (b200icp_scan_get_pose).  rms_per_iter / npairs_per_iter (optional) receive
 * max_num_iterations entries at most. */
int b200icp_match(b200icp_ctx* ctx, const b200icp_scan* model, b200icp_scan* data,
                  const b200icp_match_params* params, double* rms_per_iter,
                  uint64_t* npairs_per_iter, b200icp_match_result* result);

/* ---- icp6D::doICP (src/slam6d/icp6D.cc:374-437): sequential matching of a scan sequence --------------------
 * For i = 1..n-1: optional odometry extrapolation (Scan::mergeCoordinatesWithRoboterPosition, scan.cc:826-833:
 * scan i is moved by transMat_{i-1} * inv(transMatOrg_{i-1})), then match(previous or metascan, current).
 * transMatOrg = 16 doubles per scan: the pose each scan was loaded with (NULL: the transMat at entry).
 * meta != 0 matches against a metascan of all scans processed so far (the last max_num_metascans when > 0),
 * rebuilt after every scan like the reference.  iterations_out (may be NULL) receives n entries (entry 0 = 0).
 * frames (may be NULL) records what Scan::transform would have pushed with the default anim = -1: per match the
 * start pose, the pose after iteration 0 and the end pose, for every scan (icp6D.cc:109, :258-279). */
int b200icp_do_icp(b200icp_ctx* ctx, b200icp_scan* const* scans, int n_scans, const b200icp_match_params* params,
                   int extrapolate_pose, int meta, int max_num_metascans, const double* transMatOrg,
                   int* iterations_out, b200icp_frames* frames);

/* ---- query-sharded match across GPUs (the reference's pICP split, scan.cc:1335-1342) --------
 * Every rank holds the whole model scan and a contiguous slice of the data scan.  The per-iteration sum of the
 * pair moments over all ranks is FUSED into the iteration kernel: each rank stores its moments into every
 * rank's mailbox over NVLink (peer-mapped memory), raises a flag, sums the rows in rank order and runs the same
 * solve -- one launch per iteration, no NCCL call, identical loop state on every rank.
 *   b200icp_comm_create       allocates this rank's mailbox; ipc_handle_out (64 bytes, may be NULL) receives its
 *                             cudaIpcMemHandle for other PROCESSES
 *   b200icp_comm_connect_ipc  all_handles = world x 64 bytes in rank order (exchange them with any host
 *                             transport, e.g. torch.distributed.all_gather)
 *   b200icp_comm_connect_local  same-process variant: contexts on (different) devices with peer access
 * Then call b200icp_match with params.sharded = 1 on every rank. */
#define B200ICP_COMM_HANDLE_BYTES 64
int b200icp_comm_create(b200icp_ctx* ctx, int rank, int world, void* ipc_handle_out);
int b200icp_comm_connect_ipc(b200icp_ctx* ctx, int world, const void* all_handles);
int b200icp_comm_connect_local(b200icp_ctx* ctx, int world, b200icp_ctx* const* all_ctx);
void* b200icp_comm_mailbox(b200icp_ctx* ctx);
int b200icp_comm_destroy(b200icp_ctx* ctx);

/* transMat of the data scan after every iteration of the context's last b200icp_match that produced a transform
 * (16 doubles each, column-major like the reference): what Scan::transform would have pushed as frames
 * (icp6D.cc:258-264, scan.cc:955-983).  Returns the number of iterations recorded (may exceed cap). */
int b200icp_last_poses(b200icp_ctx* ctx, int cap, double* transmats);

/* Per-iteration record of the context's last b200icp_match: device time of the correspondence kernel
 * and of the solve kernel (ms; zeros unless params.profile was set), the number of queries that needed
 * the warp-level ring search, and the number that ran a full search at all (the others were certified
 * unchanged by their motion budget).  Returns the number of iterations recorded (may exceed cap). */
int b200icp_last_profile(b200icp_ctx* ctx, int cap, double* nn_ms, double* solve_ms, uint32_t* stage2,
                         uint32_t* searches);

/* ---- LUM link ------------------------------------------------------------------------------------
 * Replaces: lum6DEuler::covarianceEuler (src/slam6d/lum6Deuler.cc:94-260), the per-link work of
 * FillGB3D (:265-304): pairs of (first = Source, second = Target) by Scan::getPtPairs, the 15 running
 * sums, D = MM^-1 MZ, the residual pass, C = MM / s^2 (row-major 6x6) and CD = MZ / s^2.  C = CD = 0
 * when there are fewer than 3 pairs or the clouds are identical, as in the reference.  max_dist_match2
 * is already squared (graphSlam6D::max_dist_match2_LUM).  The sparse solve of the assembled system
 * stays on the host (it is O(scans)). */
int b200icp_lum_link(b200icp_ctx* ctx, const b200icp_scan* first, const b200icp_scan* second,
                     double max_dist_match2, double C[36], double CD[6], uint64_t* npairs);
/* Replaces: lum6DQuat::covarianceQuat (src/slam6d/lum6Dquat.cc:83-240; caller: elch6Dslerp.cc:63) -- the same link
 * in the 7-parameter (translation + quaternion) linearisation: 17 running sums, D = MM^-1 MZ (7x7), residual pass,
 * C = MM / s^2 (row-major 7x7), CD = MZ / s^2.  C = CD = 0 with fewer than 3 pairs.  Same kernel, same neighbour
 * cache as b200icp_lum_link. */
int b200icp_lum_link_quat(b200icp_ctx* ctx, const b200icp_scan* first, const b200icp_scan* second,
                          double max_dist_match2, double C[49], double CD[7], uint64_t* npairs);
/* The context remembers, per (first, second) link, the neighbour found for every point of `second` and uses it to
 * seed the link's next evaluation (the graph relaxation evaluates each link once per iteration while the poses
 * barely move; a seed only bounds the search radius, the pairs do not depend on it).  This call drops all remembered
 * links and sets the memory the context may spend on them (default 4 GiB; 0 disables seeding). */
int b200icp_lum_seed_cache(b200icp_ctx* ctx, size_t limit_bytes);

/* ---- LUM / graph back-end ----------------------------------------------------------------
 * Host-side loop around b200icp_lum_link; scans keep their pose in transMat / dalignxf, points are never moved.
 *   b200icp_graph_from_poses  Graph::Graph(int nodes, double cldist2, int loopsize) (src/slam6d/graph.cc:108-127):
 *                             chain links (i, i+1) plus (j, k) for k - j > loopsize and |rPos_j - rPos_k|^2 <
 *                             cldist2.  links = [2 * cap] ints (from, to); *n_links = links needed (EINVAL when
 *                             it exceeds cap; pass links = NULL to count only).
 *   b200icp_lum_fill_gb       lum6DEuler::FillGB3D (src/slam6d/lum6Deuler.cc:265-304) for the given links: ADDS the
 *                             link blocks into dense row-major G [(6(n-1))^2] and B [6(n-1)] (scan 0 is fixed).
 *                             A link-sharded caller fills its own links and all-reduces [G|B] (parallel.py).
 *   b200icp_lum_solve_update  the rest of one doGraphSlam6D iteration (lum6Deuler.cc:377-470): X = G^-1 B by dense
 *                             Cholesky (graphSlam6D::solveCholesky, graphSlam6D.cc:245-293; the reference's default
 *                             cs_cholsol solves the same system), pose correction Ha^-1 X_i per scan and
 *                             Scan::transformToEuler (scan.cc:1061-1083) on transMat / dalignxf.
 *                             ESTATE when G is not positive definite.
 *   b200icp_lum_graph_slam    lum6DEuler::doGraphSlam6D (lum6Deuler.cc:314-479): iterates the two steps until
 *                             nr_it or sum_position_diff / n_scans <= epsilon_lum; *ret_out = that quotient.
 *   b200icp_matrix4_to_euler  Matrix4ToEuler (include/slam6d/globals.icc:540-578). */
int b200icp_graph_from_poses(const double* rpos, int n_scans, double cldist2, int loopsize, int* links,
                             int cap, int* n_links);
/* Graph::Graph(int nScans, bool loop) (src/slam6d/graph.cc:76-105): the minimally connected chain, optionally closed */
int b200icp_graph_chain(int n_scans, int loop, int* links, int cap, int* n_links);
int b200icp_lum_fill_gb(b200icp_ctx* ctx, b200icp_scan* const* scans, int n_scans, const int* links,
                        int n_links, double max_dist_match2, double* G, double* B, uint64_t* npairs_out);
int b200icp_lum_solve_update(b200icp_scan* const* scans, int n_scans, const double* G, const double* B,
                             double* sum_position_diff, b200icp_frames* frames);
int b200icp_lum_graph_slam(b200icp_ctx* ctx, b200icp_scan* const* scans, int n_scans, const int* links,
                           int n_links, double max_dist_match2, int nr_it, double epsilon_lum,
                           double* ret_out, int* iterations_out, b200icp_frames* frames);
/* Link-sharded relaxation (the north-star's "allreduce only for the global lum6D covariance sum"), for
 * a host that runs one process per GPU: rank r of `world` holds ALL scans on its device, evaluates links r, r + world,
 * ... (lum6Deuler.cc:271-298 hands the links of FillGB3D to OpenMP threads the same way), and the packed fp64 buffer
 * [G | B] ((6(n-1))^2 + 6(n-1) doubles, host memory) is summed over the ranks by the caller's `allreduce` -- one call
 * per LUM iteration: wrap MPI_Allreduce, ncclAllReduce on a staging buffer, or torch.distributed (3dtk_b200/parallel.py
 * does the latter).  Every rank then runs the identical solve + pose update, so the replicated scans stay bit-identical
 * without a broadcast.  The callback returns 0 on success; world == 1 never calls it. */
typedef int (*b200icp_allreduce_fn)(double* sum_inout, size_t count, void* user);
int b200icp_lum_graph_slam_sharded(b200icp_ctx* ctx, b200icp_scan* const* scans, int n_scans, const int* links,
                                   int n_links, double max_dist_match2, int nr_it, double epsilon_lum, int rank,
                                   int world, b200icp_allreduce_fn allreduce, void* user, double* ret_out,
                                   int* iterations_out, b200icp_frames* frames);
void b200icp_matrix4_to_euler(const double m[16], double rPosTheta[3], double rPos[3]);

/* ---- scan files and frames (the wire formats either side of the path; host only) -----------
 *   b200icp_read_uos     ScanIO_uos / readASCII (src/scanio/helper.cc:577-835): "x y z" per line, '#' comments,
 *                        up to 10 unparsable lines tolerated at the top, \n or \r\n.  *xyz_out is malloc'ed
 *                        (n rows of 3 doubles); release it with b200icp_free.  Parsed in parallel chunks.
 *   b200icp_read_pose    scanNNN.pose: position, then Euler angles in DEGREES -> radians (helper.cc:228-232)
 *   b200icp_frames_*     per-scan frame lists (BasicScan::m_frames); _transform applies the frame rule of
 *                        Scan::transform (src/slam6d/scan.cc:941-1000) given the current transMat of every scan;
 *                        _save writes scanNNN.frames exactly like BasicScan::saveFrames (basicScan.cc:902-917):
 *                        16 doubles in the default ostream format, each followed by a blank, then the type. */
enum { B200ICP_FRAME_INVALID = 0, B200ICP_FRAME_ICP = 1, B200ICP_FRAME_ICPINACTIVE = 2, B200ICP_FRAME_LUM = 3,
       B200ICP_FRAME_ELCH = 4 };   /* Scan::AlgoType, include/slam6d/scan.h:126 */
int b200icp_read_uos(const char* path, double** xyz_out, size_t* n_out);
int b200icp_read_pose(const char* path, double rPos[3], double rPosTheta[3]);
/* write_uos (src/scanio/writer.cc:146-178; what bin/scan_red emits after the reduction): x y z per line times
 * `scale`; format 0 = "%lf", 1 = "%.016e" (high precision), 2 = "%.013a" (hex floats, bit-exact). */
int b200icp_write_uos(const char* path, const double* xyz, size_t n, double scale, int format);
void b200icp_free(void* p);
b200icp_frames* b200icp_frames_create(int n_scans);
void b200icp_frames_destroy(b200icp_frames* frames);
int b200icp_frames_add(b200icp_frames* frames, int scan, const double transMat[16], int type);
int b200icp_frames_transform(b200icp_frames* frames, int scan, const double* transmats, int type, int islum);
int b200icp_frames_count(const b200icp_frames* frames, int scan);
int b200icp_frames_get(const b200icp_frames* frames, int scan, int k, double transMat[16], int* type);
int b200icp_frames_save(const b200icp_frames* frames, int scan, const char* path, int append);
/* BasicScan::readFrames (basicScan.cc:872-900): replaces the scan's list by the file's frames */
int b200icp_frames_load(b200icp_frames* frames, int scan, const char* path);
/* Graph::Graph(const std::string& netfile) (src/slam6d/graph.cc:52-74): "<nrScans> <nrLinks>" then the links;
 * *n_scans = the scan count Graph::addLink derives from the links (graph.cc:157-174).  links may be NULL to count. */
int b200icp_graph_read_net(const char* path, int* links, int cap, int* n_links, int* n_scans);

/* ---- normals --------------------------------------------------------------------------------
 * Replaces: calculateNormalsKNN + calculateNormal (src/slam6d/normals.cc:220-295, :518-558):
 * exact k nearest neighbours (the point itself included), PCA, smallest-eigenvalue eigenvector,
 * flipped so that n . (p - rPos) >= 0, unit length.  Host pointers. */
int b200icp_normals_knn(b200icp_ctx* ctx, const double* xyz, size_t n, int k, const double rPos[3],
                        double* normals_out);
/* Scan::calcNormals on a scan that is already resident (src/slam6d/scan.cc:397-426): same kernel, same grid, the
 * normals go straight into the scan's "normal reduced" array -- no host round trip, no second grid build.  rPos in
 * the frame the scan was created in.  Replaces any normals the scan had. */
int b200icp_scan_calc_normals(b200icp_ctx* ctx, b200icp_scan* scan, int k, const double rPos[3]);

/* ---- octree reduction (the step immediately before the path) ----------------------
 * Replaces: Scan::calcReducedPoints + BOctTree + GetOctTreeCenter for `-r voxel_size -O 0`
 * (src/slam6d/scan.cc:560-601; include/slam6d/Boctree.h:224-270, :612-656, :928-949, :1164-1195,
 * :1353-1355): one point per occupied leaf cube -- the cube's CENTRE -- on the reference's lattice (root cube
 * = bbox centre, half-size = max half-extent + 1.0, halved until <= voxel_size), in its depth-first order.
 * xyz_out must hold n rows; *n_out receives the number written.  Host pointers. */
int b200icp_reduce_octree_center(b200icp_ctx* ctx, const double* xyz, size_t n, double voxel_size,
                                 double* xyz_out, size_t* n_out);
/* The other extraction modes of Scan::calcReducedPoints (scan.cc:585-601), with the PointType path that carries
 * normals through the reduction (scan.cc:544-557,652-676): nrpts = 0 centre (as above; no normals -- the reference's
 * GetOctTreeCenter copies POINTDIM values out of a 3-value centre, Boctree.h:938-941), -1 GetOctTreeAvg
 * (Boctree.h:951-983: per voxel the sequential fp64 mean of every attribute, points in input order), 1
 * GetOctTreeRandom (Boctree.h:985-1019: per voxel point number (int)(length * rand() / (RAND_MAX + 1.0)), the k-th
 * voxel in depth-first order consuming the k-th value of std::rand()).  rand_seed / rand_skip: the C library stream
 * to replay (glibc; seed 1 and skip 0 = a process that has not called rand() yet).  normals / nrm_out: both or
 * neither.  Output order = the octree's depth-first leaf order. */
int b200icp_reduce_octree(b200icp_ctx* ctx, const double* xyz, const double* normals, size_t n, double voxel_size,
                          int nrpts, unsigned rand_seed, size_t rand_skip, double* xyz_out, double* nrm_out,
                          size_t* n_out);
/* glibc's rand() stream restated (random_r.c, TYPE_3): out[k] = the (skip + k)-th value rand() returns after srand(seed) */
int b200icp_glibc_rand(unsigned seed, size_t skip, size_t count, int* out);

/* ---- synthetic inputs (host only, no GPU needed) -------------------------
 * scene(geom_seed, sample_seed, n): indoor box room 2000x300x1000 cm + 4 interior walls + 20 boxes whose
 * placement derives from geom_seed; n points sampled area-proportionally on the surfaces from
 * sample_seed (std::mt19937_64), plus N(0, noise_sigma^2) per coordinate.  Two calls with the same
 * geom_seed and different sample_seed are independent resamplings of the same geometry. */
int b200icp_synth_scene(uint64_t geom_seed, uint64_t sample_seed, size_t n, double noise_sigma,
                        double* xyz_out);
/* EulerToMatrix4 (globals.icc:501-531) -- exported so every binding builds poses identically. */
void b200icp_euler_to_matrix4(const double rPos[3], const double rPosTheta[3], double out[16]);
int b200icp_m4inv(const double in[16], double out[16]);
void b200icp_mmult(const double a[16], const double b[16], double out[16]);
void b200icp_transform_points(const double xf[16], double* xyz, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* B200ICP_H_ */
