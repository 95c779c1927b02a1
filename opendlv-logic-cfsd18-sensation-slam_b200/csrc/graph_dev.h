// graph_dev.h -- device-resident Gauss-Newton system of one graph topology.
//
// Memory layout (all fp64, replica-major: replica r lives at base + r * stride):
//   est   : [x[P] | y[P] | theta[P] | lx[L] | ly[L]]                      SoA estimates
//   meas  : [el_zx[El] | el_zy[El] | eo_zx[Eo] | eo_zy[Eo] | eo_zt[Eo] | lm_zx[El] | lm_zy[El]]
//           SoA measurements; landmark edges sorted by pose (the order performSLAM inserts them in)
//           plus a second copy in landmark order so both assembly kernels read coalesced
//   V     : [b_lm 2L | H_lm 4L | b_pose 3P | H_pose 9P | H_offdiag ...]   the normal equations
//           H blocks are dense row-major; an off-diagonal block is stored once, oriented like
//           g2o's upper-triangular block matrix (rows = vertex with the lower Hessian index).
//           The first 6L doubles are the landmark part that pose-range shards must reduce.
//   Lv    : concatenated L panels of all fronts (fsize x npiv, column-major, D on the diagonal)
//   Uv    : concatenated update (Schur complement) matrices of all fronts (nupd x nupd)
//   uvec  : update vectors of the forward solve (one per front);   x : solution, solver order
// Information matrices, topology, gauge and the symbolic analysis are shared by all replicas.
#pragma once
#include "ctx.h"
#include "tileplan.h"

enum : int { EF_ACTIVE = 1, EF_OFFDIAG = 2, EF_FIRST = 4, EF_TRANS = 8 };

struct DevSym {  // symbolic analysis on the device (see symbolic.h)
  DevBuf<int> piv0, npiv, nupd, rows_ptr, upd_rows, rel, child_ptr, children, asm_ptr, solver2v;
  DevBuf<long> lptr, uptr, fbig;
  DevBuf<AsmEntry> asm_entries;
  DevBuf<int> launch_list;  // front ids grouped per (level, small|big)
  // forward-solve gather: for front g and row i, gather_src[gather_ptr[frow_ptr[g]+i] ..) are the
  // positions in the update-vector array that the children contribute to that row (child order)
  DevBuf<int> frow_ptr, gather_ptr, gather_src;
};

// Leading dimension of a front held in shared memory by factor2_kernel: the front's rows + the right-hand-side row,
// rounded up to 4 (mod 8) doubles.  With that stride the operand and accumulator fragments of mma.sync.m8n8k4.f64 (8 rows
// x 4 columns, resp. 8 rows x every second of 8 columns per load) fall on every 8-byte bank exactly twice: two
// wavefronts per 32-lane load, the minimum (an odd stride gives three to four).
#ifdef __CUDACC__
__host__ __device__
#endif
inline int front_ld(int fs) { return ((fs + 1 + 3) & ~7) + 4; }

struct LevelLaunch {
  // launch_list[list_off ..): n_tiny fronts (fs <= 64, 128-thread CTAs), then n_small (front fits in
  // shared memory, 256-thread CTAs), then n_big (front in a global scratch slab)
  int list_off = 0, n_tiny = 0, n_small = 0, n_big = 0;
  size_t smem_tiny = 0, smem_factor = 0, smem_solve = 0;
  int max_fs = 0, max_fs_tiny = 0;
  // the tiny fronts are listed by ascending size and launched in up to four size classes
  // (<= 40, <= 48, <= 56, <= 64 rows): the warp-per-front kernels are bound by how many fronts fit
  // in an SM's shared memory, so a class must not pay for the largest front of the whole level
  int tiny_cls_n[4] = {0, 0, 0, 0}, tiny_cls_fs[4] = {0, 0, 0, 0};
};

// Tiled batched factorisation (tileplan.h): one launch per (level, class of tile rows)
struct TileLaunch {
  int level = 0, list_off = 0, count = 0, T = 0;  // T = largest number of tile rows in the class
};

// Peer exchange of the landmark part between pose-range shards (graph.cu: PeerArgs); one per context
struct PeerExchange {
  int world = 0, rank = 0, cap = 0;  // cap = landmarks per slot (max shard range over the ranks)
  bool connected = false;
  unsigned long long epoch = 0;
  // how long a rank waits for its peers before it gives up (slam_b200_xchg_set_timeout_ms; default 10 s:
  // structure rebuilds, host jitter or a first launch on one rank must not look like a dead peer)
  unsigned long long timeout_ns = 10000000000ull;
  char* local = nullptr;             // this rank's region (cudaMalloc, exported through CUDA IPC)
  size_t bytes = 0;
  std::vector<char*> peers;          // region base per rank; peers[rank] == local, others opened via IPC
  std::vector<int> ranges_host;      // world x [l0, l1)
  DevBuf<char*> peer_tab;
  DevBuf<int> ranges, err;
  DevBuf<unsigned> done;
};

struct DeviceSystem {
  uint64_t structure_version = 0;  // HostGraph version this was built from
  uint64_t values_version = 0;     // HostGraph values currently on the device (replica 0)
  int P = 0, L = 0, Eo = 0, El = 0;
  int nb = 0, n = 0;               // free blocks / scalar dimension
  long nV = 0, nL = 0, nU = 0, nUvec = 0, nFbig = 0;
  int R = 0;                       // replicas currently allocated
  long estStride = 0, measStride = 0;
  // host structure
  std::vector<int> pose_b, lm_b;   // g2o block index or -1
  std::vector<int> blk_hidx;       // g2o scalar offset of every block
  std::vector<int> blk_kind_local; // (local << 1) | is_landmark
  std::vector<int> off_a, off_b, hoff_diag, hoff_off;
  std::vector<int> el_perm;        // sorted landmark-edge position -> insertion index
  Symbolic sym;
  std::vector<LevelLaunch> levels;
  std::vector<int> launch_list_host;
  double upload_seconds = 0, t_structure = 0, t_lists = 0, overlapped_seconds = 0;
  // device structure
  DevBuf<unsigned char> pose_free, lm_free;
  DevBuf<int> pose_boff, lm_boff;  // solver scalar offset or -1
  DevBuf<int> el_start;            // landmark edges sorted by pose: edge range of every pose
  DevBuf<int4> el_rec;             // ONE 16-byte record per pose-sorted edge {pose, lm, block slot, flags}
  DevBuf<double> el_info;          // [3][El] SoA
  DevBuf<int> lm_start, lm_edges;  // CSR landmark -> sorted edge positions
  DevBuf<int> lmo_pose;            // landmark order: pose of the edge, -1 if inactive
  DevBuf<double> lmo_info;         // landmark order: [3][El] information
  std::vector<int> lm_order;       // host copy of lm_edges (to lay measurements out in landmark order)
  std::vector<int> s_lm_host, el_start_host;  // landmark of every pose-sorted edge; edge range per pose
  int shard_p0 = -1, shard_p1 = -1, shard_l0 = 0, shard_l1 = 0;  // cached landmark range of the last pose shard
  DevBuf<int> eo_i, eo_j, eo_slot, eo_flags, po_start, po_list;
  DevBuf<double> eo_info;          // [6][Eo] SoA
  DevSym ds;
  // tiled path for replica batches (solver.cu: factor_tile_kernel / backward_tile_kernel): chosen at structure
  // time when the graph is analysed for a batch and every front has <= TILE_MAX_ROWS local rows.  Lv then holds
  // the fronts' tile storage (L and Schur-complement tiles together), Uv / uvec are unused.
  bool tile_path = false;
  TilePlan tile;
  std::vector<TileLaunch> tile_launches;   // factor order (levels ascending); the backward sweep walks it in reverse
  DevBuf<int> tile_list, tile_item_ptr, tile_item_nv;
  DevBuf<long> tile_fptr;
  DevBuf<int2> tile_items;
  // device values
  DevBuf<double> est, meas, V, Lv, Uv, uvec, x, Fbig, chi2, chi2_part, est0, trig;
  DevBuf<int> status;              // per replica: [0] fail flag, [1] iterations done (= chi2 slot)
  DevBuf<long long> dbg_clocks;    // optional (SLAM_B200_PHASE_CLOCKS): phase clocks of one factor CTA
  DevBuf<long long> timeline;      // optional (SLAM_B200_TIMELINE): %globaltimer stamps per front, 6 per front
  PeerExchange xchg;
  int chi2_cap = 0, chi2_blocks = 0;
  int iters_enqueued = 0;
  bool assembled = false;
  bool have_snapshot = false;
  bool assembly_only = false;
  bool batch_ordering = false;
  // one Gauss-Newton iteration (assemble + factor + solves + update) captured as a CUDA graph:
  // ~40 dependent small launches replayed with one host call per iteration
  cudaGraphExec_t iter_graph = nullptr;
  int launches_per_iter = 0;
  long launches_captured = 0;  // kernel launches counted while the iteration was captured
  void drop_graph() {
    if (iter_graph) cudaGraphExecDestroy(iter_graph);
    iter_graph = nullptr;
  }
  // profiling (bench): CUDA events between the phases of every iteration
  bool profile = false;
  std::vector<cudaEvent_t> prof_events;  // 6 per iteration: start, assembled, factored, forward, backward, updated
};

int graph_build_structure(slam_b200_ctx* c);            // host: index mapping, blocks, symbolic
int graph_alloc_values(slam_b200_ctx* c, int R);        // device value arrays for R replicas
int graph_upload_host_values(slam_b200_ctx* c);         // replica 0 <- HostGraph numbers
int graph_enqueue_assemble(slam_b200_ctx* c, int p0, int p1, bool chi2_only, bool peer = false);
void graph_shard_landmarks(DeviceSystem& D, int p0, int p1, int* l0, int* l1);  // landmark range a pose shard touches
void xchg_release(DeviceSystem& D);
int graph_enqueue_solve(slam_b200_ctx* c);              // factor + forward + backward + update
int graph_enqueue_iteration(slam_b200_ctx* c);          // one GN iteration (CUDA graph replay when possible)
