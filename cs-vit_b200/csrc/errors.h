// Error plumbing for the C ABI: nothing throws across the boundary.  Every entry point returns 0 on success
// or a non-zero code after recording a thread-local message retrievable with csvit_last_error().
#pragma once
#include <cuda_runtime.h>

namespace csvit {

// printf-style; returns a non-zero error code so callers can `return set_error(...)`.
int set_error(const char* fmt, ...);
const char* last_error();

// Function attributes (cudaFuncSetAttribute) belong to a device / context, not to the process: a launcher keeps one of these as
// a function-local static and configures its kernel the first time it runs on EACH device.
struct DeviceOnce {
  unsigned long long mask = 0;
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
  }
};

// Programmatic dependent launch (PDL): kernels launched through launch_pdl() may be scheduled while the previous kernel of the
// stream is still draining - their prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps its tail - and call
// griddep_wait() (common.cuh) before they touch global memory; every such kernel also calls griddep_launch() at its top so that
// ITS successor may be scheduled early.  Only kernels that contain the wait are launched this way.  Opt-in (CSVIT_PDL=1): with the CUDA-graph replay of the step it measured
// 2-3 % SLOWER than plain stream order on this pool (profiles/r2_pdl_experiment.txt), so the default launch carries no attribute.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace csvit

#define CSVIT_CUDA(expr)                                                                              \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return ::csvit::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define CSVIT_REQUIRE(cond, ...)                          \
  do {                                                    \
    if (!(cond)) return ::csvit::set_error(__VA_ARGS__);  \
  } while (0)
