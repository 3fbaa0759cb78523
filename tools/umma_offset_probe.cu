// Development probe (not part of the library): does a SWIZZLE_128B K-major UMMA operand work when its start address is
// NOT 1024-byte aligned - i.e. rows k .. k+127 of a taller tile whose swizzle phase is fixed by absolute shared-memory
// address bits (as TMA writes it)?  And with a row-group stride (SBO) other than 1024?  That is what a convolution needs
// to read all nine in-plane taps out of ONE halo tile instead of re-fetching the tile per tap.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I worldforge_b200/csrc tools/umma_offset_probe.cu -o gpurun_out/umma_probe
//   run:   gpurun_out/umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace wf;

constexpr int ROWS = 320;   // rows in the tall A buffer (128 B each: 32 tf32)

__device__ __forceinline__ uint64_t desc_sw128_bo(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_offset) {
  uint64_t d = umma_desc_sw128(addr, lbo, sbo);
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  return d;
}

// A_full [ROWS][32] fp32, B [64][32] fp32 (K-major both).  D[m][n] = sum_k A_full[row(m)][k] * B[n][k] with
// row(m) = shift + (m / 8) * group_rows + (m % 8): groups of 8 consecutive rows, group pitch group_rows rows.
__global__ void probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int shift, int group_rows,
                      int base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // ROWS * 128 B
  uint8_t* sb = smem + ROWS * 128;          // 64 * 128 B (ROWS*128 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  // write both tiles the way TMA does for SWIZZLE_128B: 16-byte chunk index XOR (row & 7), row taken from ABSOLUTE address bits 7..9
  for (int i = tid; i < ROWS * 8; i += blockDim.x) {
    const int r = i / 8, ch = i % 8;
    const uint32_t off = r * 128 + ch * 16;
    const uint32_t a = smem_u32(sa) + off;
    const uint32_t sw = a ^ (((a >> 7) & 7) << 4);
    *reinterpret_cast<float4*>(sa + (sw - smem_u32(sa))) = *reinterpret_cast<const float4*>(A + r * 32 + ch * 4);
  }
  for (int i = tid; i < 64 * 8; i += blockDim.x) {
    const int r = i / 8, ch = i % 8;
    const uint32_t a = smem_u32(sb) + r * 128 + ch * 16;
    const uint32_t sw = a ^ (((a >> 7) & 7) << 4);
    *reinterpret_cast<float4*>(sb + (sw - smem_u32(sb))) = *reinterpret_cast<const float4*>(B + r * 32 + ch * 4);
  }
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc(2, 128, 64, 0, 0);
    for (int ks = 0; ks < 4; ++ks) {          // K = 32 tf32 = 4 steps of 8
      const uint64_t da = desc_sw128_bo(smem_u32(sa) + shift * 128 + ks * 32, 16, group_rows * 128, base_offset);
      const uint64_t db = umma_desc_sw128(smem_u32(sb) + ks * 32, 16, 1024);
      umma_tf32_ss(tmem, da, db, idesc, ks != 0);
    }
    umma_commit(bar);
  }
  __syncthreads();
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = tid >> 5;
  uint32_t r[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld_32x32b_x32(tmem + ((warp * 32u) << 16) + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[(warp * 32 + (tid & 31)) * 64 + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

static float tf32(float x) {
  uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y;
}

int main() {
  std::vector<float> A(ROWS * 32), B(64 * 32), D(128 * 64);
  srand(1);
  for (auto& v : A) v = tf32((rand() % 2001 - 1000) / 500.0f);
  for (auto& v : B) v = tf32((rand() % 2001 - 1000) / 500.0f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = ROWS * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int groups[] = {8, 16, 18};
  for (int g : groups)
    for (int shift = 0; shift < 10; ++shift)
      for (int bo = 0; bo < 8; ++bo) {
        if (shift + 15 * g + 8 > ROWS) continue;
        if (bo != 0 && bo != (shift & 7)) continue;       // candidates: 0 and (start >> 7) & 7
        cudaMemset(dD, 0, D.size() * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, shift, g, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("group_rows=%d shift=%d base_offset=%d: CUDA error %s\n", g, shift, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            const int row = shift + (m / 8) * g + (m % 8);
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += static_cast<double>(A[row * 32 + k]) * B[n * 32 + k];
            const double err = fabs(ref - D[m * 64 + n]);
            if (err > maxerr) maxerr = err;
          }
        printf("group_rows=%2d shift=%d base_offset=%d: max err %.3g %s\n", g, shift, bo, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
      }
  return 0;
}
