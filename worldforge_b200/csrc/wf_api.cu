// worldforge_b200 - C-ABI plumbing: error reporting, device queries, TMA descriptor encoding.
#include <cudaTypedefs.h>

#include <string.h>

#include <mutex>
#include <string>

#include "host_util.h"

namespace wf {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& n = cache[dev & 63];
  if (n == 0) {
    int v = 0;
    n = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
  }
  return n;
}

// cuTensorMapEncodeTiled is a driver-API symbol; resolve it through the runtime so the
// library links against libcudart only (libcuda.so.1 exists only where a driver is installed).
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  auto fn = encode_fn();
  if (!fn) return fail(WF_ECUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(WF_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
  return WF_OK;
}

}  // namespace wf

extern "C" const char* wf_last_error(void) { return wf::g_last_error.c_str(); }
extern "C" int wf_abi_version(void) { return WF_ABI_VERSION; }
extern "C" int wf_sm_count(void) { return wf::sm_count(); }

// ---- peer memory (one process per GPU on one box): cudaMalloc + CUDA IPC ---------------------------------------------
extern "C" int wf_peer_alloc(long long bytes, void** ptr, void* handle64) {
  WF_REQUIRE(bytes > 0 && ptr && handle64, "wf_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  WF_CUDA_OK(cudaMalloc(&p, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return wf::fail(WF_ECUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return WF_OK;
}
extern "C" int wf_peer_open(const void* handle64, void** ptr) {
  WF_REQUIRE(handle64 && ptr, "wf_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  WF_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return WF_OK;
}
extern "C" int wf_peer_close(void* ptr) {
  WF_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return WF_OK;
}
extern "C" int wf_peer_free(void* ptr) {
  WF_CUDA_OK(cudaFree(ptr));
  return WF_OK;
}
