// Development probe (not part of the library): clocks per tcgen05.mma issued back to back from shared memory, by kind and
// tile width N (M = 128; K = 8 for tf32, 16 for bf16; operands K-major SWIZZLE_128B, contents irrelevant).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I worldforge_b200/csrc tools/umma_rate_probe.cu -o build/umma_rate
#include <cstdio>
#include "common.cuh"
#include "attn_math.cuh"
using namespace wf;

__global__ void probe(int kind, int N, int reps, int chains, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(kind == 3 ? 1 : kind, 128, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t da = umma_desc_sw128(a0 + k * 32, 16, 1024), db = umma_desc_sw128(b0 + k * 32, 16, 1024);
        for (int c = 0; c < chains; ++c) {        // independent accumulators, round-robin
          const uint32_t d = tmem + c * (512 / chains);
          if (kind == 2) umma_tf32_ss(d, da + c * 1024, db, idesc, 1);
          else if (kind == 1) umma_f16_ss(d, da + c * 1024, db, idesc, 1);
          else umma_f16_ts(d, tmem + 448 + k * 8, db, umma_idesc(1, 128, N, 0, 0), 1);     // kind 3: bf16, A from TMEM
        }
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 160 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2000;
  for (int kind : {2, 1, 3})
    for (int N : {64, 96, 128, 192, 256})
      for (int da : {1, 2, 4}) {
        if (da * N > (kind == 3 ? 448 : 512)) continue;
        probe<<<1, 128, smem>>>(kind, N, reps, da, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%s N=%3d chains=%d: %.1f clocks per MMA (floor %d)\n", kind == 2 ? "tf32 K=8  SS" : kind == 1 ? "bf16 K=16 SS" : "bf16 K=16 TS", N, da, double(c) / (reps * 4 * da), N / 2);
      }
  return 0;
}
