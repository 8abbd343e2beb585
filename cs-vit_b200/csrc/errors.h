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
