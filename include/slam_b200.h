/* slam_b200.h -- C ABI of the B200-native GraphSLAM back end.
 *
 * Drop-in boundary for the hot path of cfsd/opendlv-logic-cfsd18-sensation-slam: the private
 * back-half methods of class Slam (src/slam.hpp:66-91) and its `g2o::SparseOptimizer m_optimizer`
 * member (src/slam.hpp:98).  Every entry point below names the reference interface it replaces.
 * Plain C types only; the handle is opaque; callers own every pointer they pass; nothing is
 * allocated across the ABI; nothing throws (a C++ exception raised inside the library, e.g. a failed
 * host allocation, is caught at the entry point and comes back as SLAM_B200_E_NOMEM / SLAM_B200_E_STATE
 * with its message in slam_b200_last_error()).  All floating point is fp64, all indices int32.
 *
 * Return value: >= 0 success (meaning per function), < 0 one of the SLAM_B200_E_* codes;
 * slam_b200_last_error() gives a human-readable reason.  There is no CPU fallback: without a
 * CUDA device every compute entry point returns SLAM_B200_E_CUDA.
 *
 * Threading: a context may be used from any host thread, one call at a time (the reference
 * serialises the same calls with m_mapMutex / m_optimizerMutex, slam.hpp:105-106).  All device
 * work of a context is issued on one CUDA stream (given at creation or owned).
 *
 * "host" pointers are ordinary host memory (pinned memory makes the copies asynchronous);
 * "_dev" entry points take device pointers valid on the context's device.
 */
#ifndef SLAM_B200_H
#define SLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slam_b200_ctx slam_b200_ctx;

#define SLAM_B200_OK 0
#define SLAM_B200_E_CUDA (-100)     /* CUDA runtime error or no device */
#define SLAM_B200_E_ARG (-101)      /* bad argument (null pointer, unknown id, duplicate id ...) */
#define SLAM_B200_E_STATE (-102)    /* call not valid in the current state */
#define SLAM_B200_E_NOMEM (-103)    /* host allocation failed (std::bad_alloc caught at the entry point) */

/* association result codes written to status[] */
#define SLAM_B200_ASSOC_MATCHED 0   /* matched map cone idx[i]            (slam.cpp:584-592) */
#define SLAM_B200_ASSOC_NEW 1       /* created map cone idx[i]            (slam.cpp:608-619) */
#define SLAM_B200_ASSOC_NONE 2      /* unmatched, not added               (slam.cpp:608)     */
#define SLAM_B200_ASSOC_SKIPPED 3   /* loop closing already set           (slam.cpp:575,608) */

/* gate of the match-only / localiser association */
#define SLAM_B200_GATE_MAPPING 0    /* fabs(mapType-obsType)<1e-4 && dist<thr   (slam.cpp:576,584) */
#define SLAM_B200_GATE_LOCALIZER 1  /* dist<thr && (mapType-(int)obsType)<1e-4  (slam.cpp:360)     */

#define SLAM_B200_ALGO_BRUTE 0      /* every observation against every map cone, map tiles in smem */
#define SLAM_B200_ALGO_GRID 1       /* uniform-grid buckets (cell >= thr), identical results        */
#define SLAM_B200_ALGO_GRID_PIPELINED 2 /* ALGO_GRID, launched so that successive frames on the
                                      * context's stream overlap (programmatic dependent launch,
                                      * _dev entry points only): the caller gives every frame in
                                      * flight its own idx buffer and reads results after a stream
                                      * synchronisation, an event, or any ordinary launch/copy     */

#define SLAM_B200_ALGO_GRID_BATCHED 3 /* slam_b200_assoc_bulk_frames_dev only: ALGO_GRID with up to 32
                                      * independent frames per launch (frozen maps), all on the stream of the
                                      * first frame's context; same results, same rules for the idx buffers   */

/* ------------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device`.  `stream` is a cudaStream_t to issue all work on, or
 * NULL to let the context create and own one.  Replaces Slam::setupOptimizer (slam.cpp:53-65). */
int slam_b200_create(int device, void* stream, slam_b200_ctx** out);
int slam_b200_destroy(slam_b200_ctx* ctx);
const char* slam_b200_last_error(const slam_b200_ctx* ctx);
int slam_b200_version(void);
/* blocks until all work issued on the context's stream has finished */
int slam_b200_sync(slam_b200_ctx* ctx);
/* Optional, once after create (the reference pays its set-up in Slam::setupOptimizer, slam.cpp:53-65, before the
 * first frame as well): takes every first-use cost out of the frames -- loads the kernels, sets their attributes,
 * allocates the frame mailbox and the device arrays of a graph of about poses_hint poses / landmarks_hint
 * landmarks by optimising a synthetic ring of that size once, and runs one mapping and one localiser frame.  The
 * context must be empty (no vertices, no map cones) and is empty again afterwards.  Without it the first
 * optimizeGraph() of a drive -- the loop-closing frame -- carries tens of milliseconds of one-off work. */
int slam_b200_warmup(slam_b200_ctx* ctx, int poses_hint, int landmarks_hint);

/* ------------------------------------------------------------------------------------------------
 * polar -> Cartesian conversion on the device
 * Replaces Slam::coneToGlobal (slam.cpp:499-510), Slam::Spherical2Cartesian (637-654) and
 * Slam::transformConeToCoG (513-523) for a whole frame.
 *   cones4xN : host, column-major 4 x n (azimuth deg, zenith deg, range, type) = the frame matrix
 *              performSLAM receives (slam.cpp:298).
 *   global3  : host out, n x 3 (x, y, type) or NULL.   local3 : host out, n x 3 (x, y, z) or NULL.
 * ---------------------------------------------------------------------------------------------- */
int slam_b200_cones_to_global(slam_b200_ctx* ctx, const double* cones4xN, int n,
                              const double pose[3], double* global3, double* local3);

/* ------------------------------------------------------------------------------------------------
 * cone map (replaces std::vector<Cone> m_map, slam.hpp:110; element = Cone{x,y,type}, cone.hpp:51-55;
 * the cone id is its index, slam.cpp:556,610)
 * ---------------------------------------------------------------------------------------------- */
int slam_b200_map_clear(slam_b200_ctx* ctx);
int slam_b200_map_append(slam_b200_ctx* ctx, const double* x, const double* y, const int32_t* type, int n);
int slam_b200_map_size(slam_b200_ctx* ctx);
int slam_b200_map_read(slam_b200_ctx* ctx, int first, int n, double* x, double* y, int32_t* type);
/* Slam::updateMap (slam.cpp:713-732) with caller-provided coordinates */
int slam_b200_map_write_xy(slam_b200_ctx* ctx, int first, int n, const double* x, const double* y);
/* Slam::updateMap (slam.cpp:713-732) on the device: every map cone j whose landmark vertex (id j, slam.cpp:556,610)
 * is in the graph takes the vertex estimate -- a kernel behind the optimise on the context's stream, no host round
 * trip -- and the map is copied to a pinned host mirror asynchronously (an event marks the copy).  Falls back to the
 * host estimates when they are newer than the device's.  Returns the number of cones updated. */
int slam_b200_map_update_from_graph(slam_b200_ctx* ctx);
/* The pinned host mirror of the cone map (what sendCones / drawCones read).  Waits for the copy
 * that refreshes it only, not for other work on the stream; refreshes first if the map has changed since.  The
 * pointers stay valid until the next call that changes the map.  Returns the number of cones. */
int slam_b200_map_mirror(slam_b200_ctx* ctx, const double** x, const double** y, const int32_t** type);

/* ------------------------------------------------------------------------------------------------
 * association
 * ---------------------------------------------------------------------------------------------- */
/* Mapping-phase association of one frame.  Replaces the matching / map-growth / loop-closure
 * detection / current-cone tracking of Slam::addConesToMap (slam.cpp:552-623) incl.
 * Slam::loopClosing (697-706) and Slam::distanceBetweenCones (708-711).  The device map grows by
 * the cones created.  The graph side effects are returned as records:
 *   idx[i], status[i] : see SLAM_B200_ASSOC_*;   z2 : n x 2 vehicle-frame xy each edge carries
 *   (addConeMeasurement, 539-545);   g3 : n x 3 coneToGlobal of every column (572).
 *   *current_cone_index, *loop_closing : in/out = m_currentConeIndex, m_loopClosing.
 *   *first_cone_created : 1 if the map was empty and cone 0 was made from column 0 (554-567; its
 *   edge precedes the per-observation edges).  *loop_closing_obs : column that set m_loopClosing
 *   in this frame, else -1 (the caller then runs optimise once per column >= it, 625-633).
 * Returns the new map size. */
int slam_b200_assoc_map_frame(slam_b200_ctx* ctx, const double* cones4xN, int n, const double pose[3],
                              double same_cone_threshold, double cone_mapping_threshold,
                              uint32_t* current_cone_index, int32_t* loop_closing,
                              int32_t* idx, int32_t* status, double* z2, double* g3,
                              int32_t* first_cone_created, int32_t* loop_closing_obs);

/* Localisation-phase association of one frame.  Replaces the matching loop of Slam::localizer
 * (slam.cpp:350-387).  idx[i] = first map index passing the localiser gate or -1.
 * *current_cone_index in/out (387); *n_reobserved (366); *send_cone_data (385, only meaningful
 * when *n_reobserved > 0).  Returns *n_reobserved. */
int slam_b200_assoc_localize_frame(slam_b200_ctx* ctx, const double* cones4xN, int n, const double pose[3],
                                   double threshold, uint32_t* current_cone_index, int32_t* idx,
                                   double* g3, int32_t* n_reobserved, int32_t* send_cone_data);

/* Whole Monte-Carlo drives, one replica per thread block: the mapping phase of
 * Slam::performSLAM (slam.cpp:298-338 -> addConesToMap 552-623) for n_frames frames of each of n_replicas replicas,
 * frame state carried on the device, every replica with its own map of at most `cap` cones.  A replica stops at the
 * frame that closes its loop (closed_at[r], -1 = never, -2 = map or frame capacity exceeded).
 *   frames4: [R][F][4 * nmax] column-major cone columns, ncols: [R][F] columns of every frame (<= nmax),
 *   poses3: [R][F][3];  records: [R][F][2 * nmax + 8] int32 = the 8 frame scalars (first_cone_created,
 *   loop_closing_obs, map_n, current_cone_index, loop_closing, n_reobserved (-1 = frame not run), -, -) followed by
 *   idx[n] and status[n] exactly as slam_b200_assoc_map_frame returns them;  map_x / map_y / map_type: [R][cap]
 *   (may be NULL), map_n: [R].  Returns n_replicas. */
int slam_b200_drive_replicas(slam_b200_ctx* ctx, int n_replicas, int n_frames, int nmax, const double* frames4,
                             const int32_t* ncols, const double* poses3, double same_cone_threshold,
                             double cone_mapping_threshold, int cap, int32_t* records, double* map_x, double* map_y,
                             int32_t* map_type, int32_t* map_n, int32_t* closed_at, double* kernel_ms);

/* Match-only association of a large observation batch against the frozen device map (phase 1 of
 * addConesToMap / the localizer loop; BASELINE config "large cone field").  gate, algo: see above.
 * Host variant copies cones in and idx out; _dev variant works on device pointers (cones4xN_dev:
 * 4 x n column-major doubles, idx_dev: n int32) and only enqueues work on the context's stream. */
int slam_b200_assoc_bulk(slam_b200_ctx* ctx, const double* cones4xN, int n, const double pose[3],
                         double threshold, int gate, int algo, int32_t* idx);
int slam_b200_assoc_bulk_dev(slam_b200_ctx* ctx, const double* cones4xN_dev, int n, const double pose[3],
                             double threshold, int gate, int algo, int32_t* idx_dev);
/* A train of independent frames (frame f: context ctxs[f] -- the same handle or replicas sharing
 * one stream --, cones_dev[f] = 4 x n[f] doubles, poses[3f..3f+2], idx_dev[f]); one launch each,
 * back to back.  With SLAM_B200_ALGO_GRID_PIPELINED the frames overlap on the device; with
 * SLAM_B200_ALGO_GRID_BATCHED up to 32 frames share one launch (contexts on one device; the work goes to the stream
 * of ctxs[0]).  Returns n_frames. */
int slam_b200_assoc_bulk_frames_dev(int n_frames, slam_b200_ctx* const* ctxs, const double* const* cones4xN_dev,
                                    const int* n, const double* poses3, double threshold, int gate, int algo,
                                    int32_t* const* idx_dev);
/* (Re)builds the uniform-grid index of the device map for SLAM_B200_ALGO_GRID with cells of
 * `cell` metres (>= the threshold used later).  Called implicitly when missing or stale. */
int slam_b200_map_build_grid(slam_b200_ctx* ctx, double cell);

/* ------------------------------------------------------------------------------------------------
 * pose-landmark graph (replaces g2o::SparseOptimizer m_optimizer and the g2o types the reference
 * instantiates: VertexSE2, VertexPointXY, EdgeSE2, EdgeSE2PointXY; slam.hpp:26-35,98)
 * ---------------------------------------------------------------------------------------------- */
int slam_b200_graph_clear(slam_b200_ctx* ctx);
/* Slam::addPoseToGraph vertex part (slam.cpp:434-438) */
int slam_b200_graph_add_pose(slam_b200_ctx* ctx, int id, double x, double y, double theta);
/* Slam::addConeToGraph vertex part (slam.cpp:526-531) */
int slam_b200_graph_add_landmark(slam_b200_ctx* ctx, int id, double x, double y);
/* EdgeSE2 with explicit measurement and row-major 3x3 information (slam.cpp:447-457) */
int slam_b200_graph_add_edge_se2(slam_b200_ctx* ctx, int id_from, int id_to, const double z[3], const double info[9]);
/* Slam::addOdometryMeasurement (slam.cpp:445-459): measurement = estimate(id_prev)^-1 * SE2(pose) */
int slam_b200_graph_add_odometry(slam_b200_ctx* ctx, int id_prev, int id_cur, const double pose[3], const double info[9]);
/* EdgeSE2PointXY, row-major 2x2 information (Slam::addConeMeasurement, slam.cpp:537-547) */
int slam_b200_graph_add_edge_se2_xy(slam_b200_ctx* ctx, int pose_id, int landmark_id, const double z[2], const double info[4]);
/* Vertex::setFixed (slam.cpp:464-474) */
int slam_b200_graph_set_fixed(slam_b200_ctx* ctx, int id, int fixed);
/* Bulk loader (same semantics as the calls above, SoA host arrays).  Vertex ids must be unique. */
int slam_b200_graph_load(slam_b200_ctx* ctx,
                         int n_poses, const int32_t* pose_ids, const double* pose_est3,
                         int n_landmarks, const int32_t* lm_ids, const double* lm_est2,
                         int n_eo, const int32_t* eo_from, const int32_t* eo_to, const double* eo_z3, const double* eo_info9,
                         int n_el, const int32_t* el_pose, const int32_t* el_lm, const double* el_z2, const double* el_info4,
                         int n_fixed, const int32_t* fixed_ids);
/* Overwrites all estimates / measurements (same counts and order as loaded); NULL keeps a field. */
int slam_b200_graph_set_values(slam_b200_ctx* ctx, const double* pose_est3, const double* lm_est2,
                               const double* eo_z3, const double* el_z2);

/* m_optimizer.initializeOptimization(); m_optimizer.optimize(iters) (slam.cpp:480-481).
 * chi2[k] (may be NULL) = active chi2 after iteration k (what g2o prints in verbose mode,
 * slam.cpp:63).  Returns g2o's convention: -1 nothing to optimise, 0 factorisation failed
 * (zero pivot), else the number of iterations done.  Host estimates are refreshed on return. */
int slam_b200_graph_optimize(slam_b200_ctx* ctx, int iters, double* chi2);

/* VertexSE2/VertexPointXY::estimate() (slam.cpp:418-420, 719-720).  Returns the dimension. */
int slam_b200_graph_get_vertex(slam_b200_ctx* ctx, int id, double out[3]);
/* all estimates in insertion order: poses n x 3, landmarks n x 2 (either may be NULL) */
int slam_b200_graph_get_estimates(slam_b200_ctx* ctx, double* pose_est3, double* lm_est2);
int slam_b200_graph_num_poses(slam_b200_ctx* ctx);
int slam_b200_graph_num_landmarks(slam_b200_ctx* ctx);
int slam_b200_graph_num_edges(slam_b200_ctx* ctx);
/* active chi2 at the current estimate (SparseOptimizer::computeActiveErrors + activeRobustChi2) */
int slam_b200_graph_chi2(slam_b200_ctx* ctx, double* chi2);

/* ---- finer-grained, device-resident control (bench / profiling / multi-GPU) ------------------ */
/* Uploads the graph and runs the host symbolic phase (index mapping, block structure, ordering,
 * assembly tree) if the structure changed.  Returns the scalar dimension of the system. */
int slam_b200_graph_prepare(slam_b200_ctx* ctx);
/* Like prepare but without the symbolic phase: only slam_b200_graph_assemble_async / export /
 * chi2 are valid afterwards (used to measure the edge-partitioned assembly of very large graphs). */
int slam_b200_graph_prepare_assembly_only(slam_b200_ctx* ctx);
/* Enqueues `iters` Gauss-Newton iterations on the stream without any host synchronisation or
 * copy.  Status / chi2 stay on the device until slam_b200_graph_finish. */
int slam_b200_graph_iterate_async(slam_b200_ctx* ctx, int iters);
/* Synchronises, returns like slam_b200_graph_optimize for the iterations enqueued since prepare /
 * the last finish, copies chi2 and refreshes host estimates. */
int slam_b200_graph_finish(slam_b200_ctx* ctx, double* chi2, int chi2_cap);
/* Re-uploads the host-side estimates and measurements (device state <- host state). */
int slam_b200_graph_reset_device(slam_b200_ctx* ctx);
/* Enqueues one linearise + assemble pass (no solve).  Pose range [p0,p1) selects the edge shard
 * (edges are partitioned by their pose; landmark blocks receive partial sums) -- whole graph:
 * 0, n_poses. */
int slam_b200_graph_assemble_async(slam_b200_ctx* ctx, int p0, int p1);
/* Device pointers into the assembled system for collectives over peer memory / NCCL:
 * which: 0 = landmark diagonal blocks + landmark rhs (contiguous, the part that needs a
 * reduction across pose shards), 1 = full H value array, 2 = rhs b.  Returns element count. */
long slam_b200_graph_system_dev(slam_b200_ctx* ctx, int which, double** ptr);
/* Peer-memory exchange of that landmark part between the pose-range shards of one graph (one process
 * per GPU of one node; replaces the all-reduce).  Collective set-up, once per graph and shard layout:
 *   shard_landmarks : [l0,l1) = landmarks the edges of pose range [p0,p1) touch;
 *   xchg_create     : allocates this rank's exchange region for `cap` >= max over ranks of (l1-l0)
 *                     landmarks and returns its 64-byte CUDA IPC handle;
 *   xchg_connect    : handles = world x 64 bytes, ranges = world x [l0,l1) (both gathered over ranks).
 * graph_assemble_exchange_async(p0,p1) then enqueues linearise + assemble of the shard with the
 * landmark kernel writing its partial blocks into the rank's exported region and raising a flag in
 * every peer's region, followed by a kernel that waits for all ranks and sums the partials in rank
 * order, reading them straight from the peers' memory over NVLink, into the landmark part of V --
 * every rank ends up with the complete landmark part, bit-identical on all ranks.
 * A rank that waits longer than the time-out for a peer (default 10 s; xchg_set_timeout_ms) gives up:
 * the landmark part of V is zeroed, the graph's fail flag is raised (factorise / update leave the
 * state alone, graph_finish / graph_optimize return 0 iterations) and xchg_error turns 1 for good.
 * xchg_error: synchronises and returns 1 if a rank ever timed out waiting for a peer. */
int slam_b200_graph_shard_landmarks(slam_b200_ctx* ctx, int p0, int p1, int32_t* l0, int32_t* l1);
int slam_b200_xchg_create(slam_b200_ctx* ctx, int world, int rank, int cap, unsigned char handle_out[64]);
int slam_b200_xchg_connect(slam_b200_ctx* ctx, const unsigned char* handles, const int32_t* ranges);
int slam_b200_graph_assemble_exchange_async(slam_b200_ctx* ctx, int p0, int p1);
int slam_b200_xchg_set_timeout_ms(slam_b200_ctx* ctx, double ms);
int slam_b200_xchg_error(slam_b200_ctx* ctx);
/* Enqueues factorise + solve + update for the system currently assembled. */
int slam_b200_graph_solve_async(slam_b200_ctx* ctx);
/* Device-side copy of all estimates (every replica) and its restoration, so repeated runs start
 * from the same state without touching the host.  restore also clears status / iteration count. */
int slam_b200_graph_snapshot(slam_b200_ctx* ctx);
int slam_b200_graph_restore_async(slam_b200_ctx* ctx);
/* Per-phase device timing with CUDA events on the context's stream.  While enabled, iterations are
 * launched kernel by kernel (no CUDA-graph replay).  read: out[0..4] = ms summed over the profiled
 * iterations for assemble, factorise, forward solve, backward solve, update; out[5] = iterations. */
int slam_b200_profile_enable(slam_b200_ctx* ctx, int on);
int slam_b200_profile_read(slam_b200_ctx* ctx, double out[8]);
/* Debug aid: SM-cycle stamps of the phases of block 0 of the most recent CTA-per-front factor launch
 * (start, zeroed, H scattered, children added, factorised, forward done, written; then s, fs, number
 * of children).  Only recorded when SLAM_B200_PHASE_CLOCKS is set in the environment at prepare time. */
int slam_b200_debug_phase_clocks(slam_b200_ctx* ctx, long long out[10]);
int slam_b200_debug_tiny_clocks(slam_b200_ctx* ctx, long long out[8]);
int slam_b200_debug_panel_clocks(slam_b200_ctx* ctx, long long out[8]);
/* Debug aid: front timeline of the last Gauss-Newton iteration (SLAM_B200_TIMELINE set at prepare time): six
 * %globaltimer stamps (ns) per front -- factor kernel entry / after its dependency wait / end, backward kernel entry /
 * after its wait / end.  Returns 6 x fronts (query with out == NULL), -1 when the mode is off. */
long slam_b200_debug_timeline(slam_b200_ctx* ctx, long long* out, long cap);
/* Guard-band mode (environment SLAM_B200_GUARD=1 before the first context; tests and debugging): every
 * device array is allocated with 4 KiB of 0xFF in front of and behind it and is itself filled with 0xFF
 * (fp64 NaN / int32 -1), so out-of-bounds and uninitialised READS poison the results the parity tests
 * compare and out-of-bounds WRITES are counted by this call: returns the number of band bytes that
 * changed over all live arrays of the process (0 = clean), -1 when the mode is off. */
long slam_b200_debug_guard_check(slam_b200_ctx* ctx, long* n_arrays);
/* Measured fp64 FMA throughput of the device in TFLOP/s (roofline denominator for the solve). */
int slam_b200_fp64_peak(slam_b200_ctx* ctx, double* tflops);

/* Debug / test export of the assembled system in g2o's Hessian order (ascending vertex id among
 * non-fixed active vertices): upper-triangular scalar CSC.  Call with Ai == NULL to query
 * (returns nnz, *n = dimension).  Assembles at the current device estimate. */
long slam_b200_graph_export_system(slam_b200_ctx* ctx, int* n, int32_t* Ap, int32_t* Ai, double* Ax, double* b);
/* Statistics of the symbolic phase and last run, see DESIGN.md:
 * out[0..15] = n_scalar, n_blocks, n_fronts, n_levels, nnz_H_upper, nnz_L, factor_flops,
 * max_front, front_storage_doubles, symbolic_seconds, upload_seconds, 5 reserved. */
int slam_b200_graph_stats(slam_b200_ctx* ctx, double out[16]);
/* Export of the symbolic analysis for tests (query the length with out == NULL): what = 0 sizes
 * [nb, n, nf, nlevels, n_upd_rows, n_asm, max_front], 1 pos, 2 boff, 3 level_ptr, 4 piv0, 5 npiv,
 * 6 nupd, 7 parent, 8 rows_ptr, 9 upd_rows, 10 rel, 11 asm_ptr, 12 asm entries (hoff, r, c, meta),
 * 13 g2o scalar offset per block, 14 dims, 15 child_ptr, 16 children.  See csrc/symbolic.h. */
long slam_b200_graph_export_symbolic(slam_b200_ctx* ctx, int what, int32_t* out, long cap);
/* The same symbolic analysis on a bare block pattern, no device needed (host logic under test):
 * nb blocks of dimension dim[], n_pairs distinct off-diagonal pairs a[k] < b[k].  H blocks are
 * assumed stored diag blocks first (d*d each, block order) then the pairs (dim[a] x dim[b]). */
void* slam_b200_symbolic_create(int nb, const int32_t* dim, int n_pairs, const int32_t* a, const int32_t* b, int leaf_size);
long slam_b200_symbolic_export(void* handle, int what, int32_t* out, long cap);
/* what: 0 nnz(L), 1 factor flops, 2 max front, 3 seconds, 4 front storage (doubles) */
double slam_b200_symbolic_stat(void* handle, int what);
/* Host plan of the tiled batched factorisation for the analysed pattern (csrc/tileplan.h): 0 = {ok, nf, max_T,
 * n_items, front storage per replica, nH}, 1 = T, 2 = KT, 3 = fptr, 4 = item_ptr, 5 = item_nv, 6 = items. */
long slam_b200_symbolic_tileplan(void* handle, int what, int32_t* out, long cap);
void slam_b200_symbolic_destroy(void* handle);

/* ---- batched Monte-Carlo replicas of the loaded topology (BASELINE config 3) ----------------- */
/* Each replica r has its own estimates and measurements (host arrays, replica-major:
 * pose_est3 [R][P][3], lm_est2 [R][L][2], eo_z3 [R][Eo][3], el_z2 [R][El][2]); information
 * matrices, topology and gauge are those of the loaded graph.  Runs `iters` GN iterations on all
 * replicas; writes optimised estimates back into pose_est3 / lm_est2, chi2 [R][iters] (may be
 * NULL) and iterations_done [R] (g2o convention).  Returns 0. */
int slam_b200_graph_optimize_batch(slam_b200_ctx* ctx, int n_replicas, double* pose_est3, double* lm_est2,
                                   const double* eo_z3, const double* el_z2, int iters,
                                   double* chi2, int32_t* iterations_done);
/* Device-resident variant: allocate/upload once, iterate without copies. */
int slam_b200_batch_upload(slam_b200_ctx* ctx, int n_replicas, const double* pose_est3, const double* lm_est2,
                           const double* eo_z3, const double* el_z2);
int slam_b200_batch_iterate_async(slam_b200_ctx* ctx, int iters);
int slam_b200_batch_download(slam_b200_ctx* ctx, double* pose_est3, double* lm_est2, double* chi2,
                             int chi2_cap_per_replica, int32_t* iterations_done);

/* number of kernel launches issued by this context since creation (bench bookkeeping) */
long slam_b200_launch_count(slam_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SLAM_B200_H */
