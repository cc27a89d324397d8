// Micro-benchmark (B200): rate of tcgen05.mma kind::f16 SS-mode (both operands in 128B-swizzled K-major shared memory)
// as a function of N, operands resident, optionally with (a) operand starts that are not aligned to the 1024-byte
// swizzle atom (the row-box convolution reads tap s at +128*s bytes) and (b) a concurrent TMA load stream into other
// shared-memory buffers (what the producer of a real kernel does).  Prints cycles per M=128,K=16 MMA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/mma_rate.cu -o tools/bin/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../gan-heightmaps_b200/csrc/tc_ptx.cuh"
using namespace hm::ptx;

// mode bit0: alternate between two accumulators; bit1: A start offset +128*(it%5) bytes (unaligned to the atom)
// fill: 0 = no TMA traffic; k>0 = one 16 KB TMA box per k MMA groups of 4 (k=1: 16 KB per 4 MMAs)
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(const __grid_constant__ CUtensorMap tm, int N, int iters, int mode,
                                                          int fill, int commit_every, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar, fbar[4], cbar[8];
  __shared__ uint32_t tmem_slot;
  const uint32_t a_addr = base;                 // A region: 2 x 17 KB
  const uint32_t b_addr = base + 5 * 16384;     // B tiles: 2 x (up to 256 rows x 128 B = 32 KB)
  const uint32_t f_addr = b_addr + 65536;       // TMA fill ring: 4 x 16 KB
  for (int i = threadIdx.x; i < (5 * 16384 + 65536) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] =
        (mode & 8) ? (0x38003800u ^ ((uint32_t)(i * 2654435761u) & 0x83ff83ffu)) : 0x3c003c00u;   // random-ish / 1.0
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 4; i++) mbar_init(smem_u32(&fbar[i]), 1);
    for (int i = 0; i < 8; i++) mbar_init(smem_u32(&cbar[i]), 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    // the whole warp runs the loop (warp-uniform control flow, operands in uniform registers); one elected lane issues
    const uint32_t idesc = idesc_f16(N);
    const uint64_t bd = desc_k_sw128(b_addr);
    long long t0 = clock64();
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    // one elected lane issues everything back to back; a commit onto a ring of dummy barriers after every
    // `commit_every` groups of 4 MMAs (power of two; 0 = never)
    if (elect_one()) {
      const uint32_t cmask = commit_every > 0 ? (uint32_t)commit_every - 1u : 0xffffffffu;
      for (int g = 0; g < iters; g++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t d = tmem + ((mode & 1) ? (k & 1) * 256 : 0);
          // mode bit2: every group reads a different A tile (6 tiles) and B tile (2 tiles): streaming operands
          const uint32_t ao = (mode & 4) ? (uint32_t)(g % 5) * 16384u : 0u;
          const uint32_t bo = (mode & 4) ? (uint32_t)(g & 1) * 32768u : 0u;
          if (mode & 16)        // MN-major operands (the weight-gradient kernels): K = 16 rows of 128 B, +2048 B per step
            tc_mma_f16(d, desc_mn_sw128(a_addr + ao + k * 2048, 16384), desc_mn_sw128(b_addr + bo + k * 2048, 16384),
                       idesc_f16(N, 1, 1), 1);
          else
            tc_mma_f16(d, desc_k_sw128(a_addr + ao) + 2 * k, desc_k_sw128(b_addr + bo) + 2 * k, idesc, 1);
        }
        if (((uint32_t)g & cmask) == cmask) tc_commit(smem_u32(&cbar[(g >> 2) & 7]));
      }
    }
    __syncwarp();
    if (elect_one()) tc_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; out[148 + blockIdx.x] = (long long)(g1 - g0); }
  } else if (threadIdx.x == 64 && fill > 0) {
    // concurrent TMA stream: iters/fill boxes of 16 KB, 4 in flight
    const int nbox = iters / fill;
    uint32_t ph[4] = {0, 0, 0, 0};
    for (int i = 0; i < nbox; i++) {
      const int s = i & 3;
      if (i >= 4) { mbar_wait(smem_u32(&fbar[s]), ph[s]); ph[s] ^= 1; }
      mbar_expect_tx(smem_u32(&fbar[s]), 16384);
      tma_load_2d(&tm, f_addr + s * 16384, smem_u32(&fbar[s]), 0, (int)((blockIdx.x * 131 + i * 128) % 65536));
    }
    for (int s = 0; s < 4 && s < nbox; s++) mbar_wait(smem_u32(&fbar[s]), ph[s]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  long long* d;
  cudaMalloc(&d, 296 * sizeof(long long));
  void* g;
  cudaMalloc(&g, 65536 * 2 * 128);                       // [131072 rows][64 halves]
  cudaMemset(g, 0, 65536 * 2 * 128);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, 131072};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncFn)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 40000;
  for (int grid : {148})
    for (int mode : {12, 28})
      for (int fill : {0, 1})
      for (int ce : {0})
        for (int N : {64, 128}) {
          mma_rate_kernel<<<grid, 128, 216 * 1024>>>(tm, N, iters, mode, fill, ce, d);
          cudaError_t e = cudaGetLastError();
          if (e == cudaSuccess) e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[296];
          cudaMemcpy(h, d, 296 * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
          double clk = (double)mx / (iters * 4.0);
          double mhz = (double)h[0] / (double)h[148] * 1e3;
          printf("commit every %d groups: grid %3d mode %d  TMA fill %5.1f B/clk-at-ideal  N %3d: %6.1f clk per MMA (ideal %3d)  fill achieved %.1f B/clk   SM clock %.0f MHz  -> %.0f TFLOP/s chip\n",
                 ce, grid, mode, fill ? 16384.0 / (fill * 4 * (N / 2.0)) : 0.0, N, clk, N / 2,
                 fill ? 16384.0 / (fill * 4 * clk) : 0.0, mhz, 2.0 * 128 * N * 16 / clk * mhz * 1e6 * 148 / 1e12);
        }
  return 0;
}
