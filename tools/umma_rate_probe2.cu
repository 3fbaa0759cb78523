// Development probe (not part of the library): clocks per tcgen05.mma for the issue PATTERNS the kernels use - fully
// unrolled issue loops from one elected lane of warp 1, BURST consecutive MMAs into one accumulator (the K steps of one
// 128-byte operand row) before moving to the next of CHAINS accumulators, optionally a tcgen05.commit after every burst.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I worldforge_b200/csrc tools/umma_rate_probe2.cu -o build/umma_rate2
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace wf;

template <int KIND, int N, int CHAINS, int BURST, bool COMMIT, bool SHARE_A, int SBO = 1024, int ROW0 = 0, int BULK = 0>
__global__ void probe(int reps, long long* out, const float* gsrc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); slot[1] = 0; fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (BULK > 0 && threadIdx.x / 32 == 2 && elect_one()) {
    // free-running global -> shared bulk copies (TMA engine + shared-memory write port busy), BULK KB each, until the issuer is done
    uint32_t ph = 0;
    volatile uint32_t* done = reinterpret_cast<volatile uint32_t*>(slot + 1);
    while (!*done) {
      mbar_arrive_expect_tx(bar + 2, BULK * 1024);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem + 164 * 1024)), "l"(gsrc + (clock64() & 0xffff) * 64), "r"(BULK * 1024), "r"(smem_u32(bar + 2)) : "memory");
      mbar_wait(bar + 2, ph); ph ^= 1;
    }
  }
  if (threadIdx.x / 32 == 1 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc(KIND, 128, N, 0, 0);
    const uint64_t a0 = umma_desc_sw128(smem_u32(smem) + ROW0 * 128, 16, SBO), b0 = umma_desc_sw128(smem_u32(smem + 128 * 1024), 16, 1024);
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
#pragma unroll
        for (int k = 0; k < BURST; ++k) {
          const uint64_t da = a0 + (SBO != 1024 ? ((c / 2) * 16 * (SBO / 128) + (c % 2) * 8 + (k >> 2) * ((k >> 2) % 3 + (SBO / 128))) * 8 : (SHARE_A ? 0 : c * 1024) + (k >> 2) * 1024 * (SHARE_A ? 1 : CHAINS)) + ((k & 3) * 2);
          const uint64_t db = b0 + ((k & 3) * 2) + (k >> 2) * 2048;
          if (KIND == 2) umma_tf32_ss(tmem + c * (512 / CHAINS), da, db, idesc, 1);
          else umma_f16_ss(tmem + c * (512 / CHAINS), da, db, idesc, 1);
        }
        if (COMMIT) umma_commit(bar + 1);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[0] = clock64() - t0;
    *reinterpret_cast<volatile uint32_t*>(slot + 1) = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

static long long* d_out;
static float* d_src;
template <int KIND, int N, int CHAINS, int BURST, bool COMMIT, bool SHARE_A, int SBO = 1024, int ROW0 = 0, int BULK = 0>
void run(int grid = 1) {
  if (CHAINS * N > 512) return;
  const int smem = 200 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(probe<KIND, N, CHAINS, BURST, COMMIT, SHARE_A, SBO, ROW0, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 1000;
  probe<KIND, N, CHAINS, BURST, COMMIT, SHARE_A, SBO, ROW0, BULK><<<grid, 128, smem>>>(reps, d_out, d_src);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%s N=%3d chains=%d burst=%d commit=%d shareA=%d sbo=%d row0=%d bulkKB=%d grid=%d: %.1f clocks per MMA (floor %d)\n", KIND == 2 ? "tf32" : "bf16", N, CHAINS,
         BURST, int(COMMIT), int(SHARE_A), SBO, ROW0, BULK, grid, double(c) / (double(reps) * CHAINS * BURST), N / 2);
}

template <int KIND, int N> void sweep() {
  run<KIND, N, 1, 1, false, false>(); run<KIND, N, 1, 4, false, false>(); run<KIND, N, 1, 8, false, false>();
  run<KIND, N, 2, 1, false, false>(); run<KIND, N, 2, 4, false, false>(); run<KIND, N, 2, 8, false, false>();
  run<KIND, N, 4, 1, false, false>(); run<KIND, N, 4, 4, false, false>(); run<KIND, N, 4, 8, false, false>();
  run<KIND, N, 2, 4, true, false>(); run<KIND, N, 4, 4, true, false>(); run<KIND, N, 4, 4, false, true>();
}

int main() {
  cudaMalloc(&d_out, 8);
  cudaMalloc(&d_src, 64 << 20); cudaMemset(d_src, 0, 64 << 20);
  // the halo kernel's operand geometry: 18-pixel rows (SBO = 18*128... the kernel uses PW = 18 -> 2304 B), start rows 0 / 19 / 38
  run<2, 96, 4, 4, false, true>(); run<2, 96, 4, 4, false, true>(148);
  run<2, 96, 4, 4, false, true, 2304, 0>(); run<2, 96, 4, 4, false, true, 2304, 19>(); run<2, 96, 4, 4, false, true, 2304, 38>();
  run<2, 96, 4, 8, false, true, 2304, 19>(); run<2, 96, 4, 8, false, true, 2304, 19>(148);
  run<2, 96, 4, 4, false, true, 1024, 0, 12>(); run<2, 96, 4, 4, false, true, 1024, 0, 24>(); run<2, 96, 4, 4, false, true, 2304, 19, 12>();
  run<2, 96, 4, 4, false, true, 2304, 19, 12>(148); run<2, 96, 4, 4, true, true, 2304, 19, 12>(148);
  run<2, 192, 2, 4, false, true, 2304, 19>(); run<2, 192, 2, 4, false, true, 2304, 19, 24>(148);
  if (getenv("WF_SWEEP")) { sweep<2, 96>(); sweep<2, 128>(); sweep<2, 192>(); sweep<2, 256>(); sweep<1, 64>(); sweep<1, 128>(); sweep<1, 256>(); }
  return 0;
}
