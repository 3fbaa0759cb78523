// worldforge_b200 - host-side helpers shared by the C-ABI entry points.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "../../include/wf_b200.h"

namespace wf {

// thread-local last error message, returned by wf_last_error()
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define WF_CUDA_OK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::wf::fail(WF_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
  } while (0)

#define WF_REQUIRE(cond, msg)                                          \
  do {                                                                 \
    if (!(cond)) return ::wf::fail(WF_EINVAL, std::string(msg));       \
  } while (0)

#define WF_LAUNCH_OK()                                                                           \
  do {                                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess) return ::wf::fail(WF_ECUDA, std::string("launch: ") + cudaGetErrorString(_e)); \
  } while (0)

// Tiled TMA descriptor over a row-major tensor of up to 5 dims (dims[0] innermost).
// strides_bytes[i] is the byte stride of dims[i+1].  Returns 0 or a WF_E* code.
int make_tmap(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

int sm_count();     // of the current device

// Runs `f` (returns a WF_* code) the first time it is reached on each CUDA device, under a lock: function attributes such as
// cudaFuncAttributeMaxDynamicSharedMemorySize are per device, so a process that launches on a second GPU must set them
// again there (one process per GPU is the deployed form, but the C ABI does not forbid more).
struct PerDeviceOnce {
  std::mutex m;
  unsigned long long seen = 0;
  template <class F> int run(F f) {
    int d = 0;
    cudaGetDevice(&d);
    std::lock_guard<std::mutex> g(m);
    const unsigned long long bit = 1ull << (d & 63);
    if (seen & bit) return 0;
    const int rc = f();
    if (rc == 0) seen |= bit;
    return rc;
  }
};

}  // namespace wf
