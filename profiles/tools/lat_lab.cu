// Latency micro-benchmarks behind the front-factorisation design (sm_100a): dependent fp64 chains,
// reciprocal, shared-memory round trip, shuffle, block barrier.  One warp (or one 512-thread CTA for
// the barrier) on one SM; cycles per dependent operation from clock64().
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lat_lab profiles/tools/lat_lab.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 512;

__global__ void lat_kernel(double* out, long long* cyc, double x0, int threads_active) {
  __shared__ double sh[1024];
  __shared__ int chase[1024];
  const int tid = threadIdx.x;
  for (int i = tid; i < 1024; i += blockDim.x) { sh[i] = 1.0 + i * 1e-9; chase[i] = (i * 33 + 7) & 1023; }
  __syncthreads();
  double x = x0 + tid * 1e-12, y = 1.0000001;
  long long t0, t1;
  // 0: dependent DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  // 1: dependent DMUL
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x * y;
  t1 = clock64();
  if (tid == 0) cyc[1] = t1 - t0;
  // 2: dependent DADD
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x + y;
  t1 = clock64();
  if (tid == 0) cyc[2] = t1 - t0;
  // 3: dependent __drcp_rn
  x = 1.5 + tid * 1e-9;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) x = __drcp_rn(x) + 0.25;
  t1 = clock64();
  if (tid == 0) cyc[3] = t1 - t0;  // includes one DADD
  // 4: dependent 1.0 / x
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) x = 1.0 / x + 0.25;
  t1 = clock64();
  if (tid == 0) cyc[4] = t1 - t0;
  // 5: shared-memory pointer chase (int)
  int p = tid & 1023;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) p = chase[p];
  t1 = clock64();
  if (tid == 0) cyc[5] = t1 - t0;
  // 6: shared load (double) + DFMA dependent: x = fma(sh[idx(x)], ...) -> address independent, value dependent
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(sh[(p + i) & 1023], x, 1e-9);
  t1 = clock64();
  if (tid == 0) cyc[6] = t1 - t0;
  // 7: shuffle (double) dependent
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (tid + 1) & 31);
  t1 = clock64();
  if (tid == 0) cyc[7] = t1 - t0;
  // 8: __syncthreads
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) __syncthreads();
  t1 = clock64();
  if (tid == 0) cyc[8] = t1 - t0;
  // 9: store to shared by one lane then barrier then load by all (publish pattern)
  t0 = clock64();
  for (int i = 0; i < N; i++) {
    if (tid == 0) sh[i & 1023] = x;
    __syncthreads();
    x += sh[i & 1023];
  }
  t1 = clock64();
  if (tid == 0) cyc[9] = t1 - t0;
  // 10: independent DFMA throughput per warp: 8 chains
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) {
    a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
    a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
  }
  t1 = clock64();
  if (tid == 0) cyc[10] = t1 - t0;
  out[blockIdx.x * blockDim.x + tid] = x + p + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMalloc(&cyc, 16 * sizeof(long long));
  const char* names[11] = {"DFMA dependent", "DMUL dependent", "DADD dependent", "__drcp_rn + DADD", "1.0/x + DADD",
                           "LDS pointer chase", "LDS value -> DFMA", "SHFL.f64 dependent", "__syncthreads",
                           "publish: STS lane0 + bar + LDS + DADD", "8 independent DFMA chains (per 8 DFMA)"};
  for (int threads : {32, 128, 512}) {
    lat_kernel<<<1, threads>>>(out, cyc, 1.0, threads);
    cudaDeviceSynchronize();
    lat_kernel<<<1, threads>>>(out, cyc, 1.0, threads);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("--- %d threads (%s)\n", threads, cudaGetErrorString(e));
    for (int k = 0; k < 11; k++) printf("%-42s %8.1f cycles/op\n", names[k], (double)h[k] / N);
  }
  return 0;
}
