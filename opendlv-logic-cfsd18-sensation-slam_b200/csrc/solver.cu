// solver.cu -- numeric phase of the block-sparse multifrontal LDL^T and the triangular solves
// (sm_100a, fp64).
//
// Replaces Eigen::SimplicialLDLT::factorize / _solve_impl (thirdparty/Eigen/src/SparseCholesky/
// SimplicialCholesky_impl.h:101-195, SimplicialCholesky.h:156-180) that the reference reaches via
// g2o's LinearSolverEigen::solve from Slam::optimizeGraph (slam.cpp:481).  Same mathematics
// (H = L D L^T without pivoting, failure iff a pivot is exactly zero), different schedule: the
// assembly tree of symbolic.cpp is walked level by level; every front is one CTA that assembles
// its dense frontal matrix in shared memory (original H blocks + the children's Schur complements),
// eliminates its pivot columns and hands its own Schur complement to the parent.  Landmark and pose
// blocks are ordered inside this factorisation; each update matrix IS the Schur complement of the
// eliminated variables, it is just never formed globally.
#include "graph_dev.h"

namespace {

constexpr int FACTOR_THREADS = 256;
constexpr int SOLVE_THREADS = 128;

struct SymArgs {
  const int *piv0, *npiv, *nupd, *rows_ptr, *upd_rows, *rel, *child_ptr, *children, *asm_ptr, *solver2v;
  const long *lptr, *uptr, *fbig;
  const AsmEntry* asm_entries;
  const int* launch_list;
};

// ---- factorisation -----------------------------------------------------------------------------
// F is the dense frontal matrix, column-major with leading dimension fs; only the lower triangle
// is meaningful.  SMEM: F in dynamic shared memory; otherwise in a per-front global scratch slab.
template <bool SMEM>
__global__ void __launch_bounds__(FACTOR_THREADS)
factor_kernel(SymArgs S, int list_off, const double* __restrict__ V_all, long nV, double* Lv_all, long nL,
              double* Uv_all, long nU, double* Fbig_all, long nFbig, int* status) {
  extern __shared__ double smem[];
  const int g = S.launch_list[list_off + blockIdx.x];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u;
  const double* V = V_all + (size_t)r * nV;
  double* Uv = Uv_all + (size_t)r * nU;
  double* F = SMEM ? smem : (Fbig_all + (size_t)r * nFbig + S.fbig[g]);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < fs * fs; t += nt) F[t] = 0.0;
  __syncthreads();
  // original entries: every H block whose earlier-eliminated vertex is a pivot of this front
  for (int q = S.asm_ptr[g] + tid; q < S.asm_ptr[g + 1]; q += nt) {
    const AsmEntry en = S.asm_entries[q];
    const double* hv = V + en.hoff;
    const int dr = en.meta & 0xff, dc = (en.meta >> 8) & 0xff;
    const bool trans = (en.meta >> 16) & 1, diag = (en.meta >> 17) & 1;
    if (diag) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j <= i; j++) F[(size_t)(en.c + j) * fs + en.r + i] = hv[i * dc + j];
    } else if (!trans) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j < dc; j++) F[(size_t)(en.c + j) * fs + en.r + i] = hv[i * dc + j];
    } else {
      for (int i = 0; i < dc; i++)
        for (int j = 0; j < dr; j++) F[(size_t)(en.c + j) * fs + en.r + i] = hv[j * dc + i];
    }
  }
  __syncthreads();
  // extend-add the children's Schur complements (one child at a time: positions of different
  // children overlap, positions inside one child do not)
  for (int ci = S.child_ptr[g]; ci < S.child_ptr[g + 1]; ci++) {
    const int ch = S.children[ci];
    const int uc = S.nupd[ch];
    const double* Uc = Uv + S.uptr[ch];
    const int* rel = S.rel + S.rows_ptr[ch];
    for (int t = tid; t < uc * uc; t += nt) {
      const int i = t % uc, j = t / uc;
      if (i >= j) F[(size_t)rel[j] * fs + rel[i]] += Uc[t];
    }
    __syncthreads();
  }
  // right-looking LDL^T on the pivot columns; columns stay unscaled (F[i,k] = l_ik d_k) until the
  // write-out so one barrier per pivot suffices
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int k = 0; k < s; k++) {
    const double d = F[(size_t)k * fs + k];
    if (d == 0.0 || !isfinite(d)) {  // SimplicialCholesky_impl.h:175-179: zero pivot = failure
      if (tid == 0) status[2 * r] = 1;
    }
    const double inv = 1.0 / d;
    const double* colk = F + (size_t)k * fs;
    for (int j = k + 1 + warp; j < fs; j += nw) {
      const double cj = colk[j] * inv;
      double* colj = F + (size_t)j * fs;
      for (int i = j + lane; i < fs; i += 32) colj[i] -= colk[i] * cj;
    }
    __syncthreads();
  }
  // L panel (fs x s, unit lower with D on the diagonal) and the Schur complement for the parent
  double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  for (int t = tid; t < fs * s; t += nt) {
    const int i = t % fs, j = t / fs;
    double v = 0.0;
    if (i == j) v = F[t];
    else if (i > j) v = F[t] / F[(size_t)j * fs + j];
    Lg[t] = v;
  }
  double* Ug = Uv + S.uptr[g];
  for (int t = tid; t < u * u; t += nt) {
    const int i = t % u, j = t / u;
    if (i >= j) Ug[t] = F[(size_t)(s + j) * fs + s + i];
  }
}

// ---- forward solve: L y = b, then z = D^-1 y, one front per CTA, leaves -> root ------------------
template <bool SMEM>
__global__ void __launch_bounds__(SOLVE_THREADS)
forward_kernel(SymArgs S, int list_off, const double* __restrict__ V_all, long nV,
               const double* __restrict__ Lv_all, long nL, double* uvec_all, long nUvec, double* x_all, int n) {
  extern __shared__ double smem[];
  const int g = S.launch_list[list_off + blockIdx.x];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u, p0 = S.piv0[g];
  const double* V = V_all + (size_t)r * nV;
  const double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  double* uvec = uvec_all + (size_t)r * nUvec;
  double* x = x_all + (size_t)r * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* w = smem;             // fs
  double* Ls = smem + fs;       // fs * s when SMEM
  for (int i = tid; i < fs; i += nt) w[i] = i < s ? V[S.solver2v[p0 + i]] : 0.0;
  if (SMEM)
    for (int t = tid; t < fs * s; t += nt) Ls[t] = Lg[t];
  __syncthreads();
  for (int ci = S.child_ptr[g]; ci < S.child_ptr[g + 1]; ci++) {
    const int ch = S.children[ci];
    const int uc = S.nupd[ch];
    const double* uv = uvec + S.rows_ptr[ch];
    const int* rel = S.rel + S.rows_ptr[ch];
    for (int i = tid; i < uc; i += nt) w[rel[i]] += uv[i];
    __syncthreads();
  }
  const double* Lp = SMEM ? Ls : Lg;
  for (int k = 0; k < s; k++) {
    const double wk = w[k];
    const double* col = Lp + (size_t)k * fs;
    for (int i = k + 1 + tid; i < fs; i += nt) w[i] -= col[i] * wk;
    __syncthreads();
  }
  for (int i = tid; i < fs; i += nt) {
    if (i < s) x[p0 + i] = w[i] / Lp[(size_t)i * fs + i];
    else uvec[S.rows_ptr[g] + i - s] = w[i];
  }
}

// ---- backward solve: L^T x = z, root -> leaves ----------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(SOLVE_THREADS)
backward_kernel(SymArgs S, int list_off, const double* __restrict__ Lv_all, long nL, double* x_all, int n) {
  extern __shared__ double smem[];
  const int g = S.launch_list[list_off + blockIdx.x];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u, p0 = S.piv0[g];
  const double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  double* x = x_all + (size_t)r * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  double* xs = smem;
  double* Ls = smem + fs;
  const int* rows = S.upd_rows + S.rows_ptr[g];
  for (int i = tid; i < fs; i += nt) xs[i] = i < s ? x[p0 + i] : x[rows[i - s]];
  if (SMEM)
    for (int t = tid; t < fs * s; t += nt) Ls[t] = Lg[t];
  __syncthreads();
  const double* Lp = SMEM ? Ls : Lg;
  // contribution of the already-solved ancestor rows: xs[k] -= sum_{i>=s} L[i,k] xs[i]
  for (int k = warp; k < s; k += nw) {
    const double* col = Lp + (size_t)k * fs;
    double t = 0.0;
    for (int i = s + lane; i < fs; i += 32) t += col[i] * xs[i];
    for (int o = 16; o; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) xs[k] -= t;
  }
  __syncthreads();
  // unit upper-triangular solve with L11^T, column sweep from the last pivot
  for (int i = s - 1; i > 0; i--) {
    const double xi = xs[i];
    for (int k = tid; k < i; k += nt) xs[k] -= Lp[(size_t)k * fs + i] * xi;
    __syncthreads();
  }
  for (int i = tid; i < s; i += nt) x[p0 + i] = xs[i];
}

SymArgs sym_args(const DeviceSystem& D) {
  SymArgs a;
  a.piv0 = D.ds.piv0.p; a.npiv = D.ds.npiv.p; a.nupd = D.ds.nupd.p; a.rows_ptr = D.ds.rows_ptr.p;
  a.upd_rows = D.ds.upd_rows.p; a.rel = D.ds.rel.p; a.child_ptr = D.ds.child_ptr.p;
  a.children = D.ds.children.p; a.asm_ptr = D.ds.asm_ptr.p; a.solver2v = D.ds.solver2v.p;
  a.lptr = D.ds.lptr.p; a.uptr = D.ds.uptr.p; a.fbig = D.ds.fbig.p;
  a.asm_entries = D.ds.asm_entries.p; a.launch_list = D.ds.launch_list.p;
  return a;
}

bool g_attr_set = false;

}  // namespace

int graph_enqueue_update(slam_b200_ctx* c);  // graph.cu

static int solver_init_attrs(slam_b200_ctx* c) {
  if (!g_attr_set) {
    int lim = c->max_smem_optin;
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    g_attr_set = true;
  }
  return 0;
}

int graph_enqueue_solve(slam_b200_ctx* c) {
  DeviceSystem& D = *c->sys;
  if (int rc = solver_init_attrs(c)) return rc;
  SymArgs S = sym_args(D);
  const size_t smem_limit = (size_t)std::max(0, c->max_smem_optin - 1024);
  const int nlv = (int)D.levels.size();
  auto mark = [&]() {
    if (!D.profile) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    D.prof_events.push_back(e);
  };
  for (int lv = 0; lv < nlv; lv++) {
    const LevelLaunch& LL = D.levels[lv];
    if (LL.n_small) {
      dim3 grid(LL.n_small, D.R);
      factor_kernel<true><<<grid, FACTOR_THREADS, LL.smem_factor, c->stream>>>(
          S, LL.list_off, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig, D.status.p);
      c->launches++;
    }
    if (LL.n_big) {
      dim3 grid(LL.n_big, D.R);
      factor_kernel<false><<<grid, FACTOR_THREADS, 0, c->stream>>>(
          S, LL.list_off + LL.n_small, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig, D.status.p);
      c->launches++;
    }
  }
  mark();  // factored
  // solves: a level's fronts either all stage their L panel in shared memory or none does
  for (int lv = 0; lv < nlv; lv++) {
    const LevelLaunch& LL = D.levels[lv];
    int nfr = LL.n_small + LL.n_big;
    if (!nfr) continue;
    dim3 grid(nfr, D.R);
    // smem_solve was clamped to the limit; recompute whether every front of the level fits
    bool fits = true;
    size_t need = 0;
    for (int q = 0; q < nfr; q++) {
      int f = D.sym.level_ptr[lv] + q;  // same set as the launch list of this level
      size_t fs = (size_t)D.sym.npiv[f] + D.sym.nupd[f];
      size_t nd = (fs * D.sym.npiv[f] + fs) * sizeof(double);
      need = std::max(need, nd);
      if (nd > smem_limit) fits = false;
    }
    if (fits)
      forward_kernel<true><<<grid, SOLVE_THREADS, need, c->stream>>>(S, LL.list_off, D.V.p, D.nV, D.Lv.p, D.nL,
                                                                   D.uvec.p, D.nUvec, D.x.p, D.n);
    else {
      size_t wneed = 0;
      for (int q = 0; q < nfr; q++) {
        int f = D.sym.level_ptr[lv] + q;
        wneed = std::max(wneed, ((size_t)D.sym.npiv[f] + D.sym.nupd[f]) * sizeof(double));
      }
      forward_kernel<false><<<grid, SOLVE_THREADS, wneed, c->stream>>>(S, LL.list_off, D.V.p, D.nV, D.Lv.p, D.nL,
                                                                     D.uvec.p, D.nUvec, D.x.p, D.n);
    }
    c->launches++;
  }
  mark();  // forward done
  for (int lv = nlv - 1; lv >= 0; lv--) {
    const LevelLaunch& LL = D.levels[lv];
    int nfr = LL.n_small + LL.n_big;
    if (!nfr) continue;
    dim3 grid(nfr, D.R);
    bool fits = true;
    size_t need = 0, wneed = 0;
    for (int q = 0; q < nfr; q++) {
      int f = D.sym.level_ptr[lv] + q;
      size_t fs = (size_t)D.sym.npiv[f] + D.sym.nupd[f];
      size_t nd = (fs * D.sym.npiv[f] + fs) * sizeof(double);
      need = std::max(need, nd);
      wneed = std::max(wneed, fs * sizeof(double));
      if (nd > smem_limit) fits = false;
    }
    if (fits)
      backward_kernel<true><<<grid, SOLVE_THREADS, need, c->stream>>>(S, LL.list_off, D.Lv.p, D.nL, D.x.p, D.n);
    else
      backward_kernel<false><<<grid, SOLVE_THREADS, wneed, c->stream>>>(S, LL.list_off, D.Lv.p, D.nL, D.x.p, D.n);
    c->launches++;
  }
  SLAM_CUDA_TRY(c, cudaGetLastError());
  mark();  // backward done
  int rc = graph_enqueue_update(c);
  mark();  // updated
  return rc;
}

// One Gauss-Newton iteration.  Kernel arguments do not change between iterations (the chi2 slot
// is a device-side counter), so the ~40 launches are captured once into a CUDA graph and replayed.
int graph_enqueue_iteration(slam_b200_ctx* c) {
  DeviceSystem& D = *c->sys;
  static const bool env_no_graph = getenv("SLAM_B200_NO_CUDA_GRAPH") != nullptr;
  // the legacy / per-thread default streams cannot be captured
  const bool no_graph = env_no_graph || c->stream == nullptr || c->stream == cudaStreamLegacy ||
                        c->stream == cudaStreamPerThread;
  if (int rc = solver_init_attrs(c)) return rc;
  if (D.profile || no_graph) {
    if (D.profile) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, c->stream);
      D.prof_events.push_back(e);
    }
    int rc = graph_enqueue_assemble(c, 0, D.P, false);
    if (rc) return rc;
    if (D.profile) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, c->stream);
      D.prof_events.push_back(e);
    }
    return graph_enqueue_solve(c);
  }
  if (!D.iter_graph) {
    long before = c->launches;
    cudaGraph_t graph = nullptr;
    SLAM_CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = graph_enqueue_assemble(c, 0, D.P, false);
    if (!rc) rc = graph_enqueue_solve(c);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    c->launches = before;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    SLAM_CUDA_TRY(c, e);
    SLAM_CUDA_TRY(c, cudaGraphInstantiate(&D.iter_graph, graph, 0));
    cudaGraphDestroy(graph);
    D.launches_per_iter = 0;
    {
      // count the kernel nodes once: bookkeeping for slam_b200_launch_count
      // (assemble 2 + per level factor/forward/backward + update)
      int n = 2 + 1;
      for (const LevelLaunch& LL : D.levels) {
        n += (LL.n_small ? 1 : 0) + (LL.n_big ? 1 : 0);
        n += (LL.n_small + LL.n_big) ? 2 : 0;
      }
      D.launches_per_iter = n;
    }
  }
  SLAM_CUDA_TRY(c, cudaGraphLaunch(D.iter_graph, c->stream));
  c->launches += D.launches_per_iter;
  D.assembled = true;
  return 0;
}
