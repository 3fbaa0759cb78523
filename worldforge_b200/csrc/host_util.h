// worldforge_b200 - host-side helpers shared by the C-ABI entry points.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

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

int sm_count();

}  // namespace wf
