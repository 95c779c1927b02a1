// dmma_lab.cu -- fp64 tensor-core MMA (mma.sync.m8n8k4.f64) on sm_100a: fragment layout check against the
// host and issue-rate measurement, beside the plain DFMA rate.  Decides whether the batched front kernel may
// do its 8x8 tile updates with DMMA.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_lab dmma_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// D(8x8) = A(8x4) B(4x8) + C; assumed layout: g = lane/4, t = lane%4: a = A[g][t], b = B[t][g], c = C[g][2t], C[g][2t+1]
__global__ void layout_kernel(const double* A, const double* B, const double* C, double* D) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double d0, d1;
  dmma(d0, d1, A[g * 4 + t], B[t * 8 + g], C[g * 8 + 2 * t], C[g * 8 + 2 * t + 1]);
  D[g * 8 + 2 * t] = d0;
  D[g * 8 + 2 * t + 1] = d1;
}

template <int ILP>
__global__ void rate_dmma(double* out, int iters) {
  double c0[ILP], c1[ILP];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
  for (int k = 0; k < ILP; k++) { c0[k] = k; c1[k] = -k; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) dmma(c0[k], c1[k], a, b, c0[k], c1[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void rate_dfma(double* out, int iters) {
  double c[ILP];
  const double a = 0.999999 + threadIdx.x * 1e-12, b = 1e-9;
#pragma unroll
  for (int k = 0; k < ILP; k++) c[k] = k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) c[k] = fma(c[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double hA[32], hB[32], hC[64], hD[64], ref[64];
  for (int i = 0; i < 32; i++) { hA[i] = 0.37 * i - 3.1; hB[i] = 1.0 / (i + 2.5); }
  for (int i = 0; i < 64; i++) hC[i] = 0.01 * i * i - 7;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) {
    double s = hC[i * 8 + j];
    for (int k = 0; k < 4; k++) s = fma(hA[i * 4 + k], hB[k * 8 + j], s);
    ref[i * 8 + j] = s;
  }
  double *A, *B, *C, *D;
  cudaMalloc(&A, 256); cudaMalloc(&B, 256); cudaMalloc(&C, 512); cudaMalloc(&D, 512);
  cudaMemcpy(A, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, 256, cudaMemcpyHostToDevice);
  cudaMemcpy(C, hC, 512, cudaMemcpyHostToDevice);
  layout_kernel<<<1, 32>>>(A, B, C, D);
  cudaMemcpy(hD, D, 512, cudaMemcpyDeviceToHost);
  double md = 0;
  for (int i = 0; i < 64; i++) md = fmax(md, fabs(hD[i] - ref[i]));
  printf("layout check: max |D - ref| = %.3e (%s); cuda: %s\n", md, md < 1e-12 ? "layout as assumed" : "LAYOUT DIFFERS",
         cudaGetErrorString(cudaGetLastError()));
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 64 * 1024);
  const int iters = 20000;
  for (int wps : {1, 2, 4, 8}) {   // warps per SM sub-partition (4 per SM)
    const int threads = 32 * 4 * wps > 1024 ? 1024 : 32 * 4 * wps, blocks = sms * (32 * 4 * wps / threads);
    float m1 = time_ms([&] { rate_dmma<1><<<blocks, threads>>>(out, iters); });
    float m4 = time_ms([&] { rate_dmma<4><<<blocks, threads>>>(out, iters / 4); });
    float m8 = time_ms([&] { rate_dmma<8><<<blocks, threads>>>(out, iters / 8); });
    float f8 = time_ms([&] { rate_dfma<8><<<blocks, threads>>>(out, iters / 8 * 16); });
    const double nw = (double)blocks * threads / 32;
    printf("warps/SMSP %d: DMMA ILP1 %.2f TFLOP/s (%.1f clk/mma/warp), ILP4 %.2f, ILP8 %.2f TFLOP/s | DFMA ILP8 %.2f TFLOP/s\n", wps,
           nw * iters * 512 / (m1 * 1e-3) / 1e12, m1 * 1e-3 * 1.965e9 / iters, nw * iters * 512 / (m4 * 1e-3) / 1e12,
           nw * iters * 512 / (m8 * 1e-3) / 1e12, nw * (iters * 16.0) * 64 / (f8 * 1e-3) / 1e12);
  }
  return 0;
}
