// assoc_lab.cu -- experiment bench for the bulk grid association kernel (config 4).  Not product
// code: it includes the product translation unit to reach its device functions and times kernel
// VARIANTS (stages of the dependent chain, alternative decompositions) with CUDA events.
//
// build (from the repo root, after the package has been built so build/*.o exist):
//   P=opendlv-logic-cfsd18-sensation-slam_b200
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
//        -o gpurun_out/assoc_lab profiles/tools/assoc_lab.cu \
//        $P/build/capi.cu.o $P/build/graph.cu.o $P/build/solver.cu.o $P/build/symbolic.cpp.o -lcudart
#include "../../opendlv-logic-cfsd18-sensation-slam_b200/csrc/assoc.cu"

#include <algorithm>
#include <cstdlib>
#include <random>

namespace {

__global__ void lab_empty_kernel(int* out) {
  if (out == nullptr && threadIdx.x == 9999) printf("x");
}

// stage 1: observation load + conversion only
__global__ void __launch_bounds__(256) lab_convert_kernel(const double* __restrict__ cones, int n, PoseTrig pt,
                                                          int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double gx, gy, ot;
  load_obs(cones, i, pt, gx, gy, ot);
  idx[i] = (int)(gx + gy + ot);
}

// stage 0: observation load only
__global__ void __launch_bounds__(256) lab_load_kernel(const double* __restrict__ cones, int n, int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2* c2 = reinterpret_cast<const double2*>(cones + 4 * (size_t)i);
  double2 a = __ldg(c2), b = __ldg(c2 + 1);
  idx[i] = (int)(a.x + a.y + b.x + b.y);
}

// stage 2: + cell table
__global__ void __launch_bounds__(256) lab_cells_kernel(const double* __restrict__ cones, int n, PoseTrig pt,
                                                        GridParams gp, const int* __restrict__ cell_start,
                                                        int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double gx, gy, ot;
  load_obs(cones, i, pt, gx, gy, ot);
  int r = -1;
  if (isfinite(gx) && isfinite(gy)) {
    const double fx0 = floor((gx - gp.h - gp.x0) * gp.inv), fx1 = floor((gx + gp.h - gp.x0) * gp.inv);
    const double fy0 = floor((gy - gp.h - gp.y0) * gp.inv), fy1 = floor((gy + gp.h - gp.y0) * gp.inv);
    if (fx1 >= 0.0 && fx0 <= (double)(gp.nx - 1) && fy1 >= 0.0 && fy0 <= (double)(gp.ny - 1)) {
      const int cx0 = (int)fmax(fx0, 0.0), cx1 = (int)fmin(fx1, (double)(gp.nx - 1));
      const int cy0 = (int)fmax(fy0, 0.0), cy1 = (int)fmin(fy1, (double)(gp.ny - 1));
      const size_t b0 = (size_t)cy0 * gp.nx, b1 = (size_t)cy1 * gp.nx;
      r = __ldg(cell_start + b0 + cx0) + __ldg(cell_start + b0 + cx1 + 1) + __ldg(cell_start + b1 + cx0) +
          __ldg(cell_start + b1 + cx1 + 1);
    }
  }
  idx[i] = r;
}


// ---- duplicated-cell index: every cone is stored in each of the <= 2 x 2 cells its disc of radius h
// overlaps, so a query reads ONE cell.  The grid is padded by one cell on every side.
__device__ __forceinline__ void dup_range(double v, double v0, double inv, double h, int nmax, int& lo, int& hi) {
  lo = min(max((int)floor((v - h - v0) * inv), 0), nmax - 1);
  hi = min(max((int)floor((v + h - v0) * inv), 0), nmax - 1);
}
__global__ void dup_count_kernel(const double* __restrict__ x, const double* __restrict__ y, int M, GridParams gp,
                                 int* __restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double a = x[i], b = y[i];
  if (!(isfinite(a) && isfinite(b))) return;
  int x0, x1, y0, y1;
  dup_range(a, gp.x0, gp.inv, gp.h, gp.nx, x0, x1);
  dup_range(b, gp.y0, gp.inv, gp.h, gp.ny, y0, y1);
  for (int cy = y0; cy <= y1; cy++)
    for (int cx = x0; cx <= x1; cx++) atomicAdd(count + (size_t)cy * gp.nx + cx, 1);
}
__global__ void dup_fill_kernel(const double* __restrict__ x, const double* __restrict__ y, const int* __restrict__ type,
                                int M, GridParams gp, int* __restrict__ cursor, GridRec* __restrict__ rec) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double a = x[i], b = y[i];
  if (!(isfinite(a) && isfinite(b))) return;
  int x0, x1, y0, y1;
  dup_range(a, gp.x0, gp.inv, gp.h, gp.nx, x0, x1);
  dup_range(b, gp.y0, gp.inv, gp.h, gp.ny, y0, y1);
  GridRec r;
  r.x = a; r.y = b; r.type = type[i]; r.idx = i; r.pad0 = 0; r.pad1 = 0;
  for (int cy = y0; cy <= y1; cy++)
    for (int cx = x0; cx <= x1; cx++) rec[atomicAdd(cursor + (size_t)cy * gp.nx + cx, 1)] = r;
}
__global__ void dup_table_kernel(const int* __restrict__ start, size_t ncell, int2* __restrict__ table) {
  size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (c < ncell) table[c] = make_int2(start[c], start[c + 1] - start[c]);
}
constexpr int SLAB = 8;
__global__ void dup_slab_kernel(const int* __restrict__ start, size_t ncell, const GridRec* __restrict__ rec,
                                GridRec* __restrict__ slab) {
  size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int s = start[c], cnt = start[c + 1] - s;
  for (int q = 0; q < SLAB; q++) {
    GridRec r;
    if (q < cnt) r = rec[s + q];
    else { r.x = r.y = __longlong_as_double(0x7ff8000000000000LL); r.type = -1000; r.idx = 0x7fffffff; r.pad1 = 0; }
    r.pad0 = cnt;
    slab[c * SLAB + q] = r;
  }
}

// query, table variant: observation -> {start,count} (one 8-byte read) -> one contiguous run
template <int GATE, bool EARLY = false>
__global__ void __launch_bounds__(BULK_THREADS)
lab_dup_table_kernel(const double* __restrict__ cones, int n, PoseTrig pt, double thr2x, GridParams gp,
                     const int2* __restrict__ table, const GridRec* __restrict__ rec, int* __restrict__ idx) {
  if (EARLY) asm volatile("griddepcontrol.launch_dependents;");  // frames are independent
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double gx, gy, ot;
  load_obs(cones, i, pt, gx, gy, ot);
  int oti = (int)ot;
  int best = 0x7fffffff;
  const double fx = floor((gx - gp.x0) * gp.inv), fy = floor((gy - gp.y0) * gp.inv);
  if (fx >= 0.0 && fx <= (double)(gp.nx - 1) && fy >= 0.0 && fy <= (double)(gp.ny - 1)) {
    const int2 sc = __ldg(table + (size_t)(int)fy * gp.nx + (int)fx);
    for (int k = 0; k < sc.y; k += GRID_BATCH) {
      double2 xy[GRID_BATCH];
      int2 ti[GRID_BATCH];
#pragma unroll
      for (int q = 0; q < GRID_BATCH; q++)
        if (k + q < sc.y) {
          const GridRec* r = rec + sc.x + k + q;
          xy[q] = __ldg(reinterpret_cast<const double2*>(r));
          ti[q] = __ldg(reinterpret_cast<const int2*>(r) + 2);
        }
#pragma unroll
      for (int q = 0; q < GRID_BATCH; q++)
        if (k + q < sc.y && type_gate<GATE>(ti[q].x, ot, oti) && ti[q].y < best &&
            cone_distance2(xy[q].x, xy[q].y, gx, gy) < thr2x)
          best = ti[q].y;
    }
  }
  idx[i] = best == 0x7fffffff ? -1 : best;
}

// query, slab variant: observation -> one 256-byte slab of <= SLAB records (count in pad0; cells
// with more cones continue in the sorted record array through the start table)
template <int GATE>
__global__ void __launch_bounds__(BULK_THREADS)
lab_dup_slab_kernel(const double* __restrict__ cones, int n, PoseTrig pt, double thr2x, GridParams gp,
                    const GridRec* __restrict__ slab, const int* __restrict__ start, const GridRec* __restrict__ rec,
                    int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double gx, gy, ot;
  load_obs(cones, i, pt, gx, gy, ot);
  int oti = (int)ot;
  int best = 0x7fffffff;
  const double fx = floor((gx - gp.x0) * gp.inv), fy = floor((gy - gp.y0) * gp.inv);
  if (fx >= 0.0 && fx <= (double)(gp.nx - 1) && fy >= 0.0 && fy <= (double)(gp.ny - 1)) {
    const size_t cell = (size_t)(int)fy * gp.nx + (int)fx;
    const GridRec* sl = slab + cell * SLAB;
    double2 xy[SLAB];
    int4 ti[SLAB];
#pragma unroll
    for (int q = 0; q < SLAB; q++) {
      xy[q] = __ldg(reinterpret_cast<const double2*>(sl + q));
      ti[q] = __ldg(reinterpret_cast<const int4*>(sl + q) + 1);
    }
#pragma unroll
    for (int q = 0; q < SLAB; q++)
      if (type_gate<GATE>(ti[q].x, ot, oti) && ti[q].y < best && cone_distance2(xy[q].x, xy[q].y, gx, gy) < thr2x)
        best = ti[q].y;
    const int cnt = ti[0].z;
    if (cnt > SLAB) {
      const int s = __ldg(start + cell);
      for (int k = SLAB; k < cnt; k++) {
        const GridRec* r = rec + s + k;
        double2 p = __ldg(reinterpret_cast<const double2*>(r));
        int2 t = __ldg(reinterpret_cast<const int2*>(r) + 2);
        if (type_gate<GATE>(t.x, ot, oti) && t.y < best && cone_distance2(p.x, p.y, gx, gy) < thr2x) best = t.y;
      }
    }
  }
  idx[i] = best == 0x7fffffff ? -1 : best;
}

struct Timer {
  cudaStream_t st;
  char* flush;
  size_t flush_bytes;
  template <class F>
  void run(const char* name, int n, bool do_flush, F launch, int reps = 15) {
    std::vector<float> v;
    for (int r = 0; r < reps; r++) {
      if (do_flush) cudaMemsetAsync(flush, r, flush_bytes, st);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0, st);
      launch();
      cudaEventRecord(e1, st);
      cudaStreamSynchronize(st);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      v.push_back(ms * 1e3f);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaError_t e = cudaGetLastError();
    std::sort(v.begin() + 0, v.end());
    printf("%-34s n=%7d %s  median %7.2f us  min %7.2f us %s\n", name, n, do_flush ? "cold" : "hot ", v[v.size() / 2],
           v[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
};

}  // namespace

int main(int argc, char** argv) {
  const int M = 1000000, NMAX = 800000;
  const double side = std::sqrt(M / 0.1), thr = 1.2;
  std::mt19937_64 rng(4);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::normal_distribution<double> G(0.0, 0.2);
  std::vector<double> mx(M), my(M);
  std::vector<int> mt(M);
  for (int i = 0; i < M; i++) { mx[i] = (U(rng) - 0.5) * side; my[i] = (U(rng) - 0.5) * side; mt[i] = 1 + (int)(U(rng) * 4) % 4; }
  const double pose[3] = {0.0, 0.0, 0.3};
  const double cp = std::cos(pose[2]), sp = std::sin(pose[2]);
  std::vector<double> fr(4 * (size_t)NMAX);
  for (int i = 0; i < NMAX; i++) {
    double gx, gy, ty;
    for (;;) {
      if (U(rng) < 0.9) {
        int k = (int)(U(rng) * M) % M;
        gx = mx[k] + G(rng); gy = my[k] + G(rng); ty = mt[k];
      } else {
        gx = (U(rng) - 0.5) * side; gy = (U(rng) - 0.5) * side; ty = 1 + (int)(U(rng) * 4) % 4;
      }
      if (cp * gx + sp * gy > 3.5) break;
    }
    double vx = cp * gx + sp * gy - 1.5, vy = -sp * gx + cp * gy;
    float az = (float)(std::atan2(vy, vx) * 180.0 / M_PI);
    if (std::fabs(az) < 1e-3f) az = 1e-3f;
    fr[4 * (size_t)i] = az; fr[4 * (size_t)i + 1] = 0; fr[4 * (size_t)i + 2] = (float)std::hypot(vx, vy); fr[4 * (size_t)i + 3] = ty;
  }
  slam_b200_ctx* c = nullptr;
  if (slam_b200_create(0, nullptr, &c)) { printf("create failed\n"); return 1; }
  slam_b200_map_append(c, mx.data(), my.data(), mt.data(), M);
  slam_b200_map_build_grid(c, thr);
  double* d_in; int *d_out, *d_ref;
  cudaMalloc(&d_in, sizeof(double) * 4 * NMAX);
  cudaMalloc(&d_out, sizeof(int) * NMAX);
  cudaMalloc(&d_ref, sizeof(int) * NMAX);
  cudaMemcpy(d_in, fr.data(), sizeof(double) * 4 * NMAX, cudaMemcpyHostToDevice);
  Timer T;
  T.st = c->stream;
  T.flush_bytes = (size_t)512 << 20;
  cudaMalloc(&T.flush, T.flush_bytes);
  PoseTrig pt{pose[0], pose[1], cp, sp};
  const double thr2x = sqrt_gate_threshold(thr);
  GridParams gp{c->grid_x0, c->grid_y0, c->grid_inv, c->grid_h, c->grid_nx, c->grid_ny};
  const int* cs = c->grid_cell_start.p;
  const GridRec* rec = c->grid_rec.p;
  cudaStream_t st = c->stream;
  auto check = [&](const char* name, int n) {
    std::vector<int> a(n), b(n);
    cudaMemcpy(a.data(), d_out, sizeof(int) * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), d_ref, sizeof(int) * n, cudaMemcpyDeviceToHost);
    int bad = 0, matched = 0;
    for (int i = 0; i < n; i++) { bad += a[i] != b[i]; matched += b[i] >= 0; }
    printf("   check %-28s mismatches %d of %d (matched %d)\n", name, bad, n, matched);
  };
  for (int pass = 0; pass < 2; pass++) {
    const bool cold = pass == 0;
    const int n = 100000;
    for (int th : {128}) {
      int blocks = (n + th - 1) / th;
      printf("-- %d threads per CTA\n", th);
      T.run("empty", n, cold, [&] { lab_empty_kernel<<<blocks, th, 0, st>>>(d_out); });
      T.run("load only", n, cold, [&] { lab_load_kernel<<<blocks, th, 0, st>>>(d_in, n, d_out); });
      T.run("load + convert", n, cold, [&] { lab_convert_kernel<<<blocks, th, 0, st>>>(d_in, n, pt, d_out); });
      T.run("load + convert + cell table", n, cold, [&] { lab_cells_kernel<<<blocks, th, 0, st>>>(d_in, n, pt, gp, cs, d_out); });
      T.run("full (product kernel)", n, cold, [&] { assoc_bulk_grid_kernel<0, false><<<blocks, th, 0, st>>>(d_in, n, pt, thr2x, gp, cs, rec, d_ref); });
    }
  }
  printf("-- scaling with the number of observations (product kernel), cold\n");
  for (int n : {32, 1000, 10000, 50000, 100000, 200000, 400000, 800000}) {
    int th = 128;
    T.run("product", n, true, [&] { assoc_bulk_grid_kernel<0, false><<<(n + th - 1) / th, th, 0, st>>>(d_in, n, pt, thr2x, gp, cs, rec, d_ref); });
  }
  // ---- duplicated-cell index ----
  GridParams gd = gp;
  {
    const double width = 1.0 / gp.inv;
    gd.x0 = gp.x0 - width; gd.y0 = gp.y0 - width; gd.nx = gp.nx + 2; gd.ny = gp.ny + 2;
  }
  const size_t ncd = (size_t)gd.nx * gd.ny;
  int *d_cnt, *d_start; int2* d_table; GridRec *d_drec, *d_slab; void* d_tmp; size_t tmpb = 0;
  cudaMalloc(&d_cnt, sizeof(int) * (ncd + 1)); cudaMalloc(&d_start, sizeof(int) * (ncd + 1));
  cudaMalloc(&d_table, sizeof(int2) * ncd); cudaMalloc(&d_slab, sizeof(GridRec) * SLAB * ncd);
  cudaMemsetAsync(d_cnt, 0, sizeof(int) * (ncd + 1), st);
  dup_count_kernel<<<(M + 255) / 256, 256, 0, st>>>(c->map_x.p, c->map_y.p, M, gd, d_cnt);
  cub::DeviceScan::ExclusiveSum(nullptr, tmpb, d_cnt, d_start, (int)(ncd + 1), st);
  cudaMalloc(&d_tmp, tmpb);
  cub::DeviceScan::ExclusiveSum(d_tmp, tmpb, d_cnt, d_start, (int)(ncd + 1), st);
  int total = 0;
  cudaMemcpyAsync(&total, d_start + ncd, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  cudaMalloc(&d_drec, sizeof(GridRec) * (size_t)total);
  cudaMemcpyAsync(d_cnt, d_start, sizeof(int) * ncd, cudaMemcpyDeviceToDevice, st);
  dup_fill_kernel<<<(M + 255) / 256, 256, 0, st>>>(c->map_x.p, c->map_y.p, c->map_type.p, M, gd, d_cnt, d_drec);
  dup_table_kernel<<<(unsigned)((ncd + 255) / 256), 256, 0, st>>>(d_start, ncd, d_table);
  dup_slab_kernel<<<(unsigned)((ncd + 255) / 256), 256, 0, st>>>(d_start, ncd, d_drec, d_slab);
  cudaStreamSynchronize(st);
  printf("-- duplicated-cell index: %zu cells, %d records (x%.2f), table %.1f MB, records %.1f MB, slabs %.1f MB  %s\n", ncd, total,
         total / (double)M, sizeof(int2) * ncd / 1e6, sizeof(GridRec) * (double)total / 1e6, sizeof(GridRec) * SLAB * (double)ncd / 1e6,
         cudaGetErrorString(cudaGetLastError()));
  {
    const int n = 100000, th = 128, blocks = (n + th - 1) / th;
    for (int pass = 0; pass < 2; pass++) {
      const bool cold = pass == 0;
      T.run("product (2x2 cells)", n, cold, [&] { assoc_bulk_grid_kernel<0, false><<<blocks, th, 0, st>>>(d_in, n, pt, thr2x, gp, cs, rec, d_ref); });
      T.run("dup cells, table", n, cold, [&] { lab_dup_table_kernel<0><<<blocks, th, 0, st>>>(d_in, n, pt, thr2x, gd, d_table, d_drec, d_out); });
      if (cold) check("dup table", n);
      T.run("dup cells, slab", n, cold, [&] { lab_dup_slab_kernel<0><<<blocks, th, 0, st>>>(d_in, n, pt, thr2x, gd, d_slab, d_start, d_drec, d_out); });
      if (cold) check("dup slab", n);
    }
  }
  // back-to-back launches over COPIES of the index and the frame whose total size exceeds L2, one
  // event pair around the whole train: the per-launch time without the event/launch floor
  {
    const int COPIES = 8, n = 100000, th = 128, K = 64, blocks = (n + th - 1) / th;
    size_t ncell1 = (size_t)gp.nx * gp.ny + 1;
    int* csc[COPIES]; GridRec* recc[COPIES]; double* inc[COPIES]; int* outc[COPIES];
    int2* tabc[COPIES]; GridRec* drecc[COPIES];
    for (int q = 0; q < COPIES; q++) {
      cudaMalloc(&csc[q], sizeof(int) * ncell1); cudaMemcpy(csc[q], cs, sizeof(int) * ncell1, cudaMemcpyDeviceToDevice);
      cudaMalloc(&recc[q], sizeof(GridRec) * M); cudaMemcpy(recc[q], rec, sizeof(GridRec) * M, cudaMemcpyDeviceToDevice);
      cudaMalloc(&inc[q], sizeof(double) * 4 * n); cudaMemcpy(inc[q], d_in, sizeof(double) * 4 * n, cudaMemcpyDeviceToDevice);
      cudaMalloc(&outc[q], sizeof(int) * n);
      cudaMalloc(&tabc[q], sizeof(int2) * ncd); cudaMemcpy(tabc[q], d_table, sizeof(int2) * ncd, cudaMemcpyDeviceToDevice);
      cudaMalloc(&drecc[q], sizeof(GridRec) * (size_t)total); cudaMemcpy(drecc[q], d_drec, sizeof(GridRec) * (size_t)total, cudaMemcpyDeviceToDevice);
    }
    cudaStream_t st2;
    cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking);
    cudaEvent_t fork, join;
    cudaEventCreateWithFlags(&fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
    auto train = [&](const char* name, int variant, bool two_streams, bool pdl) {
      float best_ms = 1e9f;
      for (int rep = 0; rep < 4; rep++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaMemsetAsync(T.flush, rep, T.flush_bytes, st);
        cudaEventRecord(e0, st);
        if (two_streams) { cudaEventRecord(fork, st); cudaStreamWaitEvent(st2, fork, 0); }
        for (int k = 0; k < K; k++) {
          int q = k % COPIES;
          cudaStream_t s_ = (two_streams && (k & 1)) ? st2 : st;
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(th); cfg.stream = s_;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
          if (variant == 0)
            cudaLaunchKernelEx(&cfg, assoc_bulk_grid_kernel<0, false>, (const double*)inc[q], n, pt, thr2x, gp, (const int*)csc[q], (const GridRec*)recc[q], outc[q]);
          else if (variant == 2)
            cudaLaunchKernelEx(&cfg, lab_dup_table_kernel<0, true>, (const double*)inc[q], n, pt, thr2x, gd, (const int2*)tabc[q], (const GridRec*)drecc[q], outc[q]);
          else
            cudaLaunchKernelEx(&cfg, lab_dup_table_kernel<0>, (const double*)inc[q], n, pt, thr2x, gd, (const int2*)tabc[q], (const GridRec*)drecc[q], outc[q]);
        }
        if (two_streams) { cudaEventRecord(join, st2); cudaStreamWaitEvent(st, join, 0); }
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best_ms = std::min(best_ms, ms);
      }
      printf("   %-44s %.2f us per frame  (%.1f%% of the HBM roofline 3.60 us)  %s\n", name, best_ms * 1e3 / K,
             100.0 * 3.6027 / (best_ms * 1e3 / K), cudaGetErrorString(cudaGetLastError()));
    };
    printf("-- trains of %d launches over %d copies (> L2), best of 4\n", K, COPIES);
    train("product, one stream", 0, false, false);
    train("product, one stream, PDL attribute", 0, false, true);
    train("product, two streams alternating", 0, true, false);
    train("dup table, one stream", 1, false, false);
    train("dup table, one stream, PDL attribute", 1, false, true);
    train("dup table, two streams alternating", 1, true, false);
    train("dup table, PDL + early launch_dependents", 2, false, true);
    cudaMemcpy(d_out, outc[3], sizeof(int) * n, cudaMemcpyDeviceToDevice);
    check("copy 3 vs product", n);
  }
  slam_b200_destroy(c);
  return 0;
}
