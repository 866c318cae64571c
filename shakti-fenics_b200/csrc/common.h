// Common infrastructure: error handling, device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/shakti_b200.h"

namespace shakti {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

#define SHAKTI_CUDA(call)                                                                  \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      throw ::shakti::Error(SHAKTI_ERR_CUDA, std::string(#call) + ": " +                   \
                                                 cudaGetErrorString(e_) + " (" + __FILE__ + \
                                                 ":" + std::to_string(__LINE__) + ")");   \
  } while (0)

#define SHAKTI_REQUIRE(cond, msg)                                          \
  do {                                                                     \
    if (!(cond)) throw ::shakti::Error(SHAKTI_ERR_INVALID, std::string(msg)); \
  } while (0)

// Number of kernels this library has launched (reported as gpu_launches by bench.py).
extern int64_t g_kernel_launches;

#define SHAKTI_LAUNCH(kernel, grid, block, smem, stream, ...)              \
  do {                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
    ++::shakti::g_kernel_launches;                                         \
    cudaError_t e_ = cudaPeekAtLastError();                                \
    if (e_ != cudaSuccess)                                                 \
      throw ::shakti::Error(SHAKTI_ERR_CUDA, std::string(#kernel) + " launch: " + \
                                                 cudaGetErrorString(e_)); \
  } while (0)

// Programmatic dependent launch (default on, SHAKTI_PDL=0 disables): the kernels of the V-cycle are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the NEXT kernel of the chain is scheduled while the
// current one drains and blocks in `griddepcontrol.wait` until its predecessor has completed and flushed.
// Every kernel launched this way calls pdl_sync() before it touches memory; launched normally that is a no-op.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <class K, class... Args>
inline cudaError_t launch_maybe_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#define SHAKTI_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                \
  do {                                                                                           \
    cudaError_t e_ = ::shakti::launch_maybe_pdl(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__); \
    ++::shakti::g_kernel_launches;                                                               \
    if (e_ != cudaSuccess)                                                                       \
      throw ::shakti::Error(SHAKTI_ERR_CUDA, std::string(#kernel) + " launch: " + cudaGetErrorString(e_)); \
  } while (0)
#endif

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) SHAKTI_CUDA(cudaMalloc(&p, count * sizeof(T)));
  }
  void alloc_zero(size_t count, cudaStream_t s = 0) {
    alloc(count);
    if (count) SHAKTI_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), s));
  }
  void upload(const std::vector<T>& h) { upload(h.data(), h.size()); }
  void upload(const T* h, size_t count) {
    alloc(count);
    if (count) SHAKTI_CUDA(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice));
  }
  std::vector<T> download(cudaStream_t s = 0) const {
    std::vector<T> h(n);
    if (n) {
      SHAKTI_CUDA(cudaStreamSynchronize(s));
      SHAKTI_CUDA(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
    }
    return h;
  }
};

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Diagnostic phase timer (SHAKTI_TRACE_PHASES=1): synchronises the stream at both ends of a scope and
// accumulates wall time per name; shakti_step prints and clears the table on rank 0.  Off by default
// (no synchronisation, no cost).
bool phase_trace_enabled();
void phase_add(const char* name, double ms);
void phase_report(const char* title);
struct PhaseScope {
  const char* name;
  cudaStream_t s;
  double t0 = 0.0;
  bool on;
  PhaseScope(const char* n, cudaStream_t st);
  ~PhaseScope();
};
#define SHAKTI_PHASE(name, stream) ::shakti::PhaseScope phase_scope_##__LINE__(name, stream)

}  // namespace shakti
