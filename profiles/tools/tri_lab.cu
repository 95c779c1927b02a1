// Micro-benchmark of step (1a) of the front factorisation (solver.cu): one warp factorises an 8 x 8
// diagonal triangle held redundantly in registers and publishes it.  Variants timed with clock64():
//  0: as in factor2_kernel (8 dependent reciprocals, one-store-per-lane publish)
//  1: only the load of the triangle
//  2: load + factorisation chain, no publish
//  3: pivots taken in pairs (the two reciprocals of a pair are independent: 1/a and 1/(a c - b^2))
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tri_lab profiles/tools/tri_lab.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 8;

template <int VAR>
__global__ void tri_kernel(double* out, long long* cyc, int fs, int reps) {
  extern __shared__ double smem[];
  const int ld = fs + 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* F = smem;
  double* Tsm = F + fs * ld;
  double* dinv = Tsm + 80;
  for (int t = tid; t < fs * ld; t += blockDim.x) { int i = t % ld, j = t / ld; F[t] = (i == j) ? 10.0 + i : 0.01 * ((i * 7 + j * 3) % 11); }
  __syncthreads();
  long long tot = 0;
  double sink = 0;
  for (int rep = 0; rep < reps; rep++) {
    const int k0 = (rep * NB) % (fs - NB);
    const int nb = NB;
    double* Pk = F + (size_t)k0 * ld;
    long long t0 = clock64();
    if (warp == 0) {
      double T[NB][NB], iv[NB];
#pragma unroll
      for (int p = 0; p < NB; p++)
#pragma unroll
        for (int q = p; q < NB; q++) T[q][p] = (q < nb) ? Pk[p * ld + k0 + q] : (q == p ? 1.0 : 0.0);
      bool bad = false;
      if (VAR == 0 || VAR == 2) {
#pragma unroll
        for (int p = 0; p < NB; p++) {
          const double d = T[p][p];
          if (p < nb && (d == 0.0 || !isfinite(d))) bad = true;
          iv[p] = __drcp_rn(d);
#pragma unroll
          for (int q = p + 1; q < NB; q++) {
            const double lqp = T[q][p] * iv[p];
#pragma unroll
            for (int q2 = q; q2 < NB; q2++) T[q2][q] -= T[q2][p] * lqp;
          }
        }
      } else if (VAR == 3) {
#pragma unroll
        for (int p = 0; p < NB; p += 2) {
          const double a = T[p][p], b = T[p + 1][p], c = T[p + 1][p + 1];
          const double det = a * c - b * b;
          const double r1 = __drcp_rn(a), r2 = __drcp_rn(det);
          if (a == 0.0 || !isfinite(a) || det == 0.0 || !isfinite(det)) bad = true;
          iv[p] = r1;
          iv[p + 1] = a * r2;
          // eliminate column p from everything right of it, then column p+1
#pragma unroll
          for (int q = p + 1; q < NB; q++) {
            const double lqp = T[q][p] * r1;
#pragma unroll
            for (int q2 = q; q2 < NB; q2++) T[q2][q] -= T[q2][p] * lqp;
          }
#pragma unroll
          for (int q = p + 2; q < NB; q++) {
            const double lqp = T[q][p + 1] * iv[p + 1];
#pragma unroll
            for (int q2 = q; q2 < NB; q2++) T[q2][q] -= T[q2][p + 1] * lqp;
          }
        }
      } else {
#pragma unroll
        for (int p = 0; p < NB; p++) iv[p] = T[p][p];
      }
      if (bad && lane == 0) out[0] = -1;
      if (VAR == 0 || VAR == 3) {
        double mine = 0.0, mine2 = 0.0, myiv = 0.0;
        int mq = 0, mp = 0, mq2 = 0, mp2 = 0;
        int e = 0;
#pragma unroll
        for (int p = 0; p < NB; p++)
#pragma unroll
          for (int q = p; q < NB; q++) {
            if (e < 32) { if (lane == e) { mine = T[q][p]; mq = q; mp = p; } }
            else if (lane == e - 32) { mine2 = T[q][p]; mq2 = q; mp2 = p; }
            e++;
          }
#pragma unroll
        for (int p = 0; p < NB; p++)
          if (lane == p) myiv = iv[p];
        Tsm[mq * NB + mp] = mine;
        if (lane < 4) Tsm[mq2 * NB + mp2] = mine2;
        if (lane < NB) { Tsm[NB * NB + lane] = myiv; dinv[k0 + lane] = myiv; }
      } else {
        double acc = 0;
#pragma unroll
        for (int p = 0; p < NB; p++) {
          acc += iv[p];
#pragma unroll
          for (int q = p; q < NB; q++) acc += T[q][p];
        }
        sink += acc;
      }
    }
    long long t1 = clock64();
    tot += t1 - t0;
    __syncthreads();
  }
  if (tid == 0) cyc[VAR] = tot / reps;
  out[1 + tid] = sink + Tsm[tid & 63];
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 2048 * sizeof(double));
  cudaMalloc(&cyc, 16 * sizeof(long long));
  const int fs = 138, reps = 200;
  size_t sm = (size_t)(fs * (fs + 1) + 80 + fs + 8) * sizeof(double);
  cudaFuncSetAttribute(tri_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaFuncSetAttribute(tri_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaFuncSetAttribute(tri_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaFuncSetAttribute(tri_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  for (int threads : {32, 512}) {
    for (int it = 0; it < 2; it++) {
      tri_kernel<0><<<1, threads, sm>>>(out, cyc, fs, reps);
      tri_kernel<1><<<1, threads, sm>>>(out, cyc, fs, reps);
      tri_kernel<2><<<1, threads, sm>>>(out, cyc, fs, reps);
      tri_kernel<3><<<1, threads, sm>>>(out, cyc, fs, reps);
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%d threads (%s): full %lld  load-only %lld  load+chain %lld  paired pivots + publish %lld cycles per 8-pivot triangle\n",
           threads, cudaGetErrorString(e), h[0], h[1], h[2], h[3]);
  }
  return 0;
}
