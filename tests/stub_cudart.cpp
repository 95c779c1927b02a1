// stub_cudart.cpp -- TEST INFRASTRUCTURE ONLY (tests/test_abi.py::test_host_code_over_a_stub_runtime and
// profiles/tools/host_sanitize.sh).  A stand-in for the ~45 CUDA runtime entry points libslam_b200.so imports,
// LD_PRELOADed into a test subprocess on a box without a GPU so that the HOST code behind the C ABI -- argument
// checks, graph_load, the structure pass, the symbolic phase, launch lists, the packing of the structure upload
// -- can run (and be sanitised) there.  "Device" memory is host heap, so every copy is a memcpy ASan can see;
// kernels and CUDA graphs are no-ops: NOTHING IS COMPUTED, this is not a fallback and the product never sees it.
// Host-to-device copies are hashed in order (stub_h2d_hash) so a test can tell whether two runs uploaded the
// same structure bit for bit.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
struct dim3 { unsigned x, y, z; };
static thread_local unsigned long g_h2d_hash = 1469598103934665603ull;
static thread_local size_t g_h2d_bytes = 0;  // per host thread: the copies of one ABI call happen on the calling thread
extern "C" {
unsigned long stub_h2d_hash() { return g_h2d_hash; }
size_t stub_h2d_bytes() { return g_h2d_bytes; }
void stub_reset() { g_h2d_hash = 1469598103934665603ull; g_h2d_bytes = 0; }
void** __cudaRegisterFatBinary(void*) { static void* h; return &h; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char*, char*, const char*, int, void*, void*, void*, void*, int*) {}
void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}
unsigned __cudaPushCallConfiguration(dim3, dim3, size_t, void*) { return 0; }
cudaError_t __cudaPopCallConfiguration(dim3*, dim3*, size_t*, void*) { return 0; }
cudaError_t cudaLaunchKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return 0; }
cudaError_t cudaLaunchKernelExC(const void*, const void*, void**) { return 0; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
cudaError_t cudaSetDevice(int) { return 0; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) { *v = (attr == 16) ? 148 : 232448; return 0; }  // 16 = multiprocessor count
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x10; return 0; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
cudaError_t cudaFree(void* p) { free(p); return 0; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return 0; }
cudaError_t cudaStreamQuery(cudaStream_t) { return 0; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static void note(const void* src, size_t n, int kind) {
  if (kind != 1) return;  // host -> device: hash what is uploaded, in order
  const unsigned char* b = (const unsigned char*)src;
  for (size_t i = 0; i < n; i++) { g_h2d_hash ^= b[i]; g_h2d_hash *= 1099511628211ull; }
  g_h2d_bytes += n;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int kind, cudaStream_t) { note(s, n, kind); memcpy(d, s, n); return 0; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int kind) { note(s, n, kind); memcpy(d, s, n); return 0; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
cudaError_t cudaGetLastError() { return 0; }
cudaError_t cudaPeekAtLastError() { return 0; }
const char* cudaGetErrorString(cudaError_t) { return "stub"; }
cudaError_t cudaFuncSetAttribute(const void*, int, int) { return 0; }
cudaError_t cudaFuncGetAttributes(void* a, const void*) { memset(a, 0, 64); return 0; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) { *n = 1; return 0; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)0x20; return 0; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)0x20; return 0; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return 0; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = (void*)0x30; return 0; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = (void*)0x40; return 0; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return 0; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return 0; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return 0; }
cudaError_t cudaIpcGetMemHandle(void* h, void*) { memset(h, 0, 64); return 0; }
cudaError_t cudaIpcOpenMemHandle(void** p, ...) { *p = nullptr; return 1; }
cudaError_t cudaIpcCloseMemHandle(void*) { return 0; }
}
