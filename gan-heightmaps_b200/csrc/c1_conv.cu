// One-channel-source convolutions on tcgen05 with the im2col operand built in shared memory (sm_100a, fast mode).
//
// The two ends of the DCGAN pair touch 512x512 one-channel images next to 64-channel tensors:
//   * the discriminator's first layer  conv5x5(1 -> 64) + LeakyReLU + 2x2 max-pool  (reference
//     architectures/dcgan.py:42-47), whose un-pooled output (64 ch @ 512^2) is 4x the bytes of anything else in the step;
//   * the input gradient of the generator's last layer  nearest-2x -> conv5x5(64 -> 1)  (dcgan.py:31-32).
// Both are the SAME gather: every output element q of the half-resolution grid reads the 6x6 patch
//     A[q][u*6+v] = img[2*qy - 2 + u][2*qx - 2 + v]          (zero outside the image)
// of a one-channel image and multiplies it with a [N][36] matrix:
//   * pooled first layer: N = 4*64 = (window position d, co); Wk[(d,co)][(u,v)] = Wf[co][u-dy][v-dx] (pack mode 15);
//     the epilogue takes the max over d (and its arg), adds the bias, applies the activation and writes the POOLED
//     tensor + the argmax bytes -- the full-resolution activation never exists;
//   * last-layer input gradient: N = 64 = ci; Wk[ci][(u,v)] = the four 3x3 phase filters of pack mode 8, transposed
//     (pack mode 14); dy[512^2] goes straight to dx_low[256^2][64].
// A CTA tile is 128 output elements (bw x bh of the half-res grid).  Builder warps stage the (2bh+4) x (2bw+4) image
// patch in shared memory (registers prefetch the next tile's patch), each builder thread then writes ITS row of the
// [128][64] fp16 operand tile in the 128B-swizzled K-major layout tcgen05 reads (K = 36, zero-padded to 48), one
// elected thread issues 3 tcgen05.mma (M=128, N, K=16) into a double-buffered TMEM accumulator, four epilogue warps
// drain it.  Weights ([N][64] fp16, <= 32 KB) are loaded once per CTA by TMA.  The epilogue stages its [128 px][64 ch]
// output tile (and the argmax bytes) in shared memory, 128B- / 64B-swizzled, and ONE thread hands each tile to the TMA
// unit (cp.async.bulk.tensor store, clipped at the image border): per-thread 16-byte stores of NHWC rows touch 32
// different 128-byte lines per warp instruction, and at 24 KB of output per 384 cycles of MMA work the LSU transaction
// rate, not HBM, was what bounded this kernel (round-2 profile: 1.9 TB/s, unchanged by a second epilogue warp group).
#include <cuda.h>
#include <stdlib.h>

#include "hm_common.cuh"
#include "tc_ptx.cuh"

namespace hm {
using namespace ptx;

constexpr int C1_THREADS = 288;        // warp 0: weights TMA + MMA issuer; warps 1-4: builders; warps 5-8: epilogue
constexpr int C1_STAGES = 3;
constexpr int C1_A_BYTES = 128 * 128;  // one [128 rows][64 k] fp16 operand tile
constexpr int C1_PATCH_WORDS = 784;    // >= max over (bw,bh) of (2bh+4)*(bw+2) 32-bit words (780 at bw=128 or bw=1)
constexpr int C1_PRE = 7;              // patch words per builder thread
constexpr int C1_Y_STAGE = 128 * 128;  // output staging: [128 px][64 ch] fp16, SWIZZLE_128B
constexpr int C1_I_STAGE = 128 * 64;   // argmax staging:  [128 px][64 ch] uint8, SWIZZLE_64B

struct C1Params {
  int B, H, W, Hq, Wq;
  int bw, bh, tiles_x, tiles_y, n_tiles;
  int N;                // GEMM columns: 64 (plain) or 256 (pooled: 4 window positions x 64 channels)
  int pool;
  int act;
  float slope;
  const __half* x;      // [B,H,W] one channel
  const float* bias;    // [64] or null
  __half* y;            // [B,Hq,Wq,64]
  uint8_t* idx;         // [B,Hq,Wq,64] argmax (pooled form)
};

template <int ACT>
__device__ __forceinline__ float c1_act(float v, float slope) {
  if (ACT == HM_ACT_LRELU) return fmaxf(v, 0.f) + slope * fminf(v, 0.f);
  if (ACT == HM_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// The four epilogue warps (128 threads, named barrier 2).  stage_s / stage_g: shared-window / generic address of the
// output staging area: two [y tile | idx tile] buffers, alternating per tile.
template <int ACT>
__device__ __forceinline__ void c1_epilogue(const C1Params& p, const CUtensorMap* tmY, const CUtensorMap* tmI,
                                            uint32_t tmem_base, uint32_t t_full0, uint32_t t_empty0,
                                            const float* bias_s, uint32_t stage_s, uint8_t* stage_g, int warp, int lane) {
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int per_img = p.tiles_x * p.tiles_y;
  const int accstride = p.pool ? 256 : 64;
  const bool issuer = q == 0 && lane == 0;
  int it = 0;
  for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, it++) {
    const int acc = it & 1;
    uint8_t* ybuf = stage_g + (it & 1) * (C1_Y_STAGE + C1_I_STAGE);
    uint8_t* ibuf = ybuf + C1_Y_STAGE;
    // the TMA store of tile it-2 has finished READING this staging buffer
    if (issuer) bulk_wait_group_read<1>();
    named_bar_sync(2, 128);
    mbar_wait(t_full0 + 8u * acc, (it >> 1) & 1);
    tc_fence_after();
    const int b = t / per_img, rem = t - b * per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * accstride;
    if (p.pool) {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v0[16], v1[16], v2[16], v3[16];
        tmem_ld16(taddr + c0, v0);
        tmem_ld16(taddr + 64 + c0, v1);
        tmem_ld16(taddr + 128 + c0, v2);
        tmem_ld16(taddr + 192 + c0, v3);
        tmem_ld_wait();
        uint32_t packed[8], kb[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float r2[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            float m = __uint_as_float(v0[j + e]);
            uint32_t k = 0;
            const float a1 = __uint_as_float(v1[j + e]), a2 = __uint_as_float(v2[j + e]), a3 = __uint_as_float(v3[j + e]);
            if (a1 > m) { m = a1; k = 1; }
            if (a2 > m) { m = a2; k = 2; }
            if (a3 > m) { m = a3; k = 3; }
            r2[e] = c1_act<ACT>(m + bias_s[c0 + j + e], p.slope);
            kb[(j + e) >> 2] |= k << (8 * ((j + e) & 3));
          }
          __half2 h = __floats2half2_rn(r2[0], r2[1]);
          packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
          // bit 2 of the argmax byte: the STORED fp16 value is on the slope-1 side of the activation (>= 0; > 0 for ReLU),
          // so that the backward pass does not have to read the pooled tensor for act'
          const uint32_t pm = ACT == HM_ACT_RELU ? __hgt2_mask(h, __float2half2_rn(0.f)) : __hge2_mask(h, __float2half2_rn(0.f));
          kb[j >> 2] |= ((pm & 0x4u) << (8 * (j & 3))) | ((pm & 0x40000u) >> 16 << (8 * ((j + 1) & 3)));
        }
        const int ch = c0 >> 3;                                // 16-byte chunk of the row's 128 bytes
        *reinterpret_cast<uint4*>(ybuf + sw128_off(row, ch)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *reinterpret_cast<uint4*>(ybuf + sw128_off(row, ch + 1)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        // SWIZZLE_64B: 16-byte chunk index (address bits 4-5) XOR address bits 7-8 = (row >> 1) & 3 at a 64-byte pitch
        *reinterpret_cast<uint4*>(ibuf + row * 64 + ((((c0 >> 4) ^ (row >> 1)) & 3) << 4)) =
            make_uint4(kb[0], kb[1], kb[2], kb[3]);
      }
    } else {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          __half2 h = __floats2half2_rn(c1_act<ACT>(__uint_as_float(v[j]) + bias_s[c0 + j], p.slope),
                                        c1_act<ACT>(__uint_as_float(v[j + 1]) + bias_s[c0 + j + 1], p.slope));
          packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
          *reinterpret_cast<uint4*>(ybuf + sw128_off(row, (c0 >> 3) + j)) =
              make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(t_empty0 + 8u * acc);              // the accumulator is drained
    fence_proxy_async();                                          // staging writes -> visible to the TMA unit
    named_bar_sync(2, 128);
    if (issuer) {
      const uint32_t ys = stage_s + (it & 1) * (C1_Y_STAGE + C1_I_STAGE);
      tma_store_4d(tmY, ys, 0, tx * p.bw, ty * p.bh, b);
      if (p.pool) tma_store_4d(tmI, ys + C1_Y_STAGE, 0, tx * p.bw, ty * p.bh, b);
      bulk_commit_group();
    }
  }
  if (issuer) bulk_wait_group<0>();                               // every store has landed before the CTA exits
}

__global__ void __launch_bounds__(C1_THREADS, 1)
    c1s2_conv_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY,
                     const __grid_constant__ CUtensorMap tmI, const C1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_bytes = (uint32_t)p.N * 128u;
  const uint32_t a_off = w_bytes;
  const uint32_t stage_off = a_off + C1_STAGES * C1_A_BYTES;             // 1024-aligned: two [y | idx] output buffers
  const uint32_t patch_off = stage_off + 2 * (C1_Y_STAGE + C1_I_STAGE);
  const uint32_t ctrl_off = patch_off + 2 * C1_PATCH_WORDS * 4;
  const uint32_t ctrl = base + ctrl_off;
  const uint32_t w_full = ctrl;
  auto a_full = [&](int s) { return ctrl + 8u * (1 + s); };
  auto a_empty = [&](int s) { return ctrl + 8u * (1 + C1_STAGES + s); };
  auto t_full = [&](int a) { return ctrl + 8u * (1 + 2 * C1_STAGES + a); };
  auto t_empty = [&](int a) { return ctrl + 8u * (3 + 2 * C1_STAGES + a); };
  const uint32_t tmem_slot = ctrl + 8u * (5 + 2 * C1_STAGES);
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(gbase + ctrl_off + 8 * (5 + 2 * C1_STAGES));
  float* bias_s = (float*)(gbase + ctrl_off + 256);
  uint32_t* patch = (uint32_t*)(gbase + patch_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = p.pool ? 512u : 128u;
  if (threadIdx.x < 64) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < C1_STAGES; s++) {
      mbar_init(a_full(s), 128);
      mbar_init(a_empty(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(t_full(a), 1);
      mbar_init(t_empty(a), 4);
    }
    mbar_fence_init();
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmY);
    prefetch_tensormap(&tmI);
  }
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (warp >= 1 && warp <= 4) {            // zero the operand stages once: chunk 5 and the tail of chunk 4 stay zero
    uint4* a4 = reinterpret_cast<uint4*>(gbase + a_off);
    for (int i = threadIdx.x - 32; i < C1_STAGES * C1_A_BYTES / 16; i += 128) a4[i] = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  const int per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ===================== weights + MMA issuer =====================
    if (elect_one()) {
      mbar_expect_tx(w_full, w_bytes);
      tma_load_2d(&tmW, base, w_full, 0, 0);
      mbar_wait(w_full, 0);
      const uint32_t idesc = idesc_f16(p.N);
      const uint64_t bd = desc_k_sw128(base);
      const int accstride = p.pool ? 256 : 64;
      int it = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, it++) {
        const int s = it % C1_STAGES, ph = (it / C1_STAGES) & 1, acc = it & 1;
        mbar_wait(t_empty(acc), ((it >> 1) & 1) ^ 1);
        mbar_wait(a_full(s), ph);
        tc_fence_after();
        const uint64_t ad = desc_k_sw128(base + a_off + s * C1_A_BYTES);
#pragma unroll
        for (int k = 0; k < 3; k++)            // K = 48 = taps 0..35 + zero padding; +32 B per K=16 slice
          tc_mma_f16(tmem_base + acc * accstride, ad + 2 * k, bd + 2 * k, idesc, k != 0);
        tc_commit(a_empty(s));
        tc_commit(t_full(acc));
      }
    }
  } else if (warp <= 4) {
    // ===================== builders =====================
    const int tb = threadIdx.x - 32;
    const int iy = tb / p.bw, ix = tb - iy * p.bw;
    const int PWW = p.bw + 2, n_words = (2 * p.bh + 4) * PWW;
    uint32_t pre[C1_PRE];
    auto prefetch = [&](int t) {
      const int b = t / per_img, rem = t - b * per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int Y0 = 2 * ty * p.bh - 2, X0 = 2 * tx * p.bw - 2;
      const __half* img = p.x + (size_t)b * p.H * p.W;
#pragma unroll
      for (int j = 0; j < C1_PRE; j++) {
        const int i = tb + 128 * j;
        uint32_t v = 0;
        if (i < n_words) {
          const int pr = i / PWW, pw = i - pr * PWW;
          const int Y = Y0 + pr, X = X0 + 2 * pw;
          if (Y >= 0 && Y < p.H && X >= 0 && X < p.W) v = *reinterpret_cast<const uint32_t*>(img + (size_t)Y * p.W + X);
        }
        pre[j] = v;
      }
    };
    int t = blockIdx.x;
    if (t < p.n_tiles) prefetch(t);
    int it = 0;
    for (; t < p.n_tiles; t += gridDim.x, it++) {
      uint32_t* pb = patch + (it & 1) * C1_PATCH_WORDS;
#pragma unroll
      for (int j = 0; j < C1_PRE; j++) {
        const int i = tb + 128 * j;
        if (i < n_words) pb[i] = pre[j];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (t + (int)gridDim.x < p.n_tiles) prefetch(t + gridDim.x);
      const int s = it % C1_STAGES, ph = (it / C1_STAGES) & 1;
      mbar_wait(a_empty(s), ph ^ 1);
      uint32_t wd[20];
#pragma unroll
      for (int u = 0; u < 6; u++)
#pragma unroll
        for (int j = 0; j < 3; j++) wd[u * 3 + j] = pb[(2 * iy + u) * PWW + ix + j];
      wd[18] = wd[19] = 0;
      uint8_t* arow = gbase + a_off + s * C1_A_BYTES;
#pragma unroll
      for (int c = 0; c < 5; c++)
        *reinterpret_cast<uint4*>(arow + sw128_off(tb, c)) = make_uint4(wd[4 * c], wd[4 * c + 1], wd[4 * c + 2], wd[4 * c + 3]);
      fence_proxy_async();
      mbar_arrive(a_full(s));
    }
  } else {
    // ===================== epilogue (warps 5..8) =====================
    const uint32_t ss = base + stage_off;
    uint8_t* sg = gbase + stage_off;
    switch (p.act) {
      case HM_ACT_LRELU: c1_epilogue<HM_ACT_LRELU>(p, &tmY, &tmI, tmem_base, t_full(0), t_empty(0), bias_s, ss, sg, warp, lane); break;
      case HM_ACT_RELU: c1_epilogue<HM_ACT_RELU>(p, &tmY, &tmI, tmem_base, t_full(0), t_empty(0), bias_s, ss, sg, warp, lane); break;
      default: c1_epilogue<HM_ACT_LINEAR>(p, &tmY, &tmI, tmem_base, t_full(0), t_empty(0), bias_s, ss, sg, warp, lane); break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

typedef CUresult (*C1EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static C1EncodeTiledFn c1_encode_fn() {
  static C1EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = (C1EncodeTiledFn)q;
  }
  return fn;
}

}  // namespace hm

using namespace hm;

// y[B,H/2,W/2,64] = act( A . Wk^T + bias ),  A[q][u*6+v] = x[2qy-2+u][2qx-2+v]  (x: one-channel fp16 image [B,H,W]).
//   ncols == 64 : plain form, wk = [64][64] fp16 (columns >= 36 ignored), idx must be NULL;
//   ncols == 256: pooled form, wk = [(d,co)][64]; y = act(max_d + bias[co]), idx = argmax d (first maximum).
extern "C" int hm_c1s2_conv(const void* x, const void* wk, const float* bias, void* y, uint8_t* idx, int B, int H, int W,
                            int ncols, int act, float slope, void* stream) {
  HM_CHECK_ARG(x && wk && y && B > 0 && H > 0 && W > 0, "hm_c1s2_conv: bad argument");
  HM_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "hm_c1s2_conv: the image must have even height and width (%dx%d)", H, W);
  HM_CHECK_ARG((ncols == 64 && !idx) || (ncols == 256 && idx), "hm_c1s2_conv: ncols must be 64 (idx NULL) or 256 (pooled, idx given)");
  HM_CHECK_ARG(act == HM_ACT_LINEAR || act == HM_ACT_LRELU || act == HM_ACT_RELU, "hm_c1s2_conv: activation %d is not supported", act);
  if ((((uintptr_t)x) & 3) || (((uintptr_t)wk | (uintptr_t)y | (uintptr_t)idx) & 15)) {
    set_error("hm_c1s2_conv: x must be 4-byte, wk / y / idx 16-byte aligned");
    return HM_ERR_ALIGN;
  }
  if (!c1_encode_fn()) {
    set_error("hm_c1s2_conv: cuTensorMapEncodeTiled is not available from this driver");
    return HM_ERR_CUDA;
  }
  C1Params p;
  p.B = B; p.H = H; p.W = W; p.Hq = H / 2; p.Wq = W / 2;
  int bw = 1;
  while (bw * 2 <= p.Wq && bw < 128) bw *= 2;
  p.bw = bw; p.bh = 128 / bw;
  p.tiles_x = (p.Wq + p.bw - 1) / p.bw;
  p.tiles_y = (p.Hq + p.bh - 1) / p.bh;
  p.n_tiles = B * p.tiles_x * p.tiles_y;
  p.N = ncols; p.pool = idx ? 1 : 0; p.act = act; p.slope = slope;
  p.x = (const __half*)x; p.bias = bias; p.y = (__half*)y; p.idx = idx;
  CUtensorMap tmW;
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)ncols};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)ncols};
    cuuint32_t es[2] = {1, 1};
    CUresult r = c1_encode_fn()(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(wk), dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("hm_c1s2_conv: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
      return HM_ERR_CUDA;
    }
  }
  // output tiles leave through TMA stores: y[B,Hq,Wq,64] fp16 (128-byte rows, SWIZZLE_128B) and idx (64-byte rows,
  // SWIZZLE_64B), box {64 ch, bw, bh, 1}; the part of a tile beyond the image is clipped by the TMA unit
  CUtensorMap tmY, tmI;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)p.Wq, (cuuint64_t)p.Hq, (cuuint64_t)B};
    cuuint32_t box[4] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    cuuint64_t sy[3] = {128, (cuuint64_t)p.Wq * 128, (cuuint64_t)p.Hq * p.Wq * 128};
    CUresult r = c1_encode_fn()(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, y, dims, sy, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && idx) {
      cuuint64_t si[3] = {64, (cuuint64_t)p.Wq * 64, (cuuint64_t)p.Hq * p.Wq * 64};
      r = c1_encode_fn()(&tmI, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, idx, dims, si, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      tmI = tmY;
    }
    if (r != CUDA_SUCCESS) {
      set_error("hm_c1s2_conv: cuTensorMapEncodeTiled failed for the output (CUresult %d)", (int)r);
      return HM_ERR_CUDA;
    }
  }
  // >= 120 KB so that only one CTA (which may own all 512 TMEM columns) is resident per SM
  size_t smem = 1024 + (size_t)ncols * 128 + C1_STAGES * C1_A_BYTES + 2 * (C1_Y_STAGE + C1_I_STAGE) +
                2 * C1_PATCH_WORDS * 4 + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(c1s2_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) {
      set_error("hm_c1s2_conv: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
      return HM_ERR_CUDA;
    }
    attr = true;
  }
  int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  c1s2_conv_kernel<<<grid, C1_THREADS, smem, (cudaStream_t)stream>>>(tmW, tmY, tmI, p);
  HM_CHECK_LAUNCH("hm_c1s2_conv");
  return HM_OK;
}

// =================================================================================================================
// Backward of the pooled first layer (conv5x5(1 -> 64) + activation + 2x2 max-pool), straight from the gradient of
// the POOLED tensor: the full-resolution gradient (64 ch @ 512^2, 3/4 of it zeros) is never written.
//   G'[w][(d,co)] = g[w][co] * act'(p[w][co]) * [idx[w][co] == d]            (w = pooling window, d = position in it)
//   weight/bias gradient:  dWk[(d,co)][k] += sum_w G'[w][(d,co)] * A[w][k]    A = the 6x6 patch operand of the forward
//                          pass with A[w][36] = 1 (so column 36 collects the bias gradient);  MN-major tcgen05 operands
//   input gradient:        U[w][k] = sum_(d,co) G'[w][(d,co)] * Wk[(d,co)][k]  (K-major operands), then hm_c1s2_col2im
//                          folds the 36 patch contributions of every window onto the image.
// Two builder groups (128 threads each) alternate tiles so that the global-load latency of one tile hides behind the
// other's build; each builder thread owns one window: it reads g, p, idx (320 B), writes its row of the four G' tiles
// (and of the patch tile) in the 128B-swizzled layout, and arrives on the stage's mbarrier.
// =================================================================================================================
namespace hm {

constexpr int CB_THREADS = 416;         // warp 0: MMA (+ weights TMA); warps 1-4, 5-8: builder groups; warps 9-12: epilogue
constexpr int CB_G_BYTES = 4 * C1_A_BYTES;            // four [128 w][64 co] tiles
constexpr int CB_STAGE_BYTES = CB_G_BYTES + C1_A_BYTES;

struct CbParams {
  int B, H, W, Hq, Wq;
  int bw, bh, tiles_x, tiles_y, n_tiles;
  int act;
  float slope;
  int want_dw, want_u;
  const __half* x;      // [B,H,W]
  const __half* g;      // [B,Hq,Wq,64] gradient of the pooled output
  const __half* pl;     // [B,Hq,Wq,64] pooled output (for act')
  const uint8_t* idx;   // [B,Hq,Wq,64]
  float* dwk;           // [256][64] fp32, atomically accumulated
  __half* u;            // [B,Hq,Wq,64]
  const float* img_scale;   // [B] or null: g is multiplied by img_scale[image] (per-sample weighting of dW, db)
  int plain;            // hm_c1s2_wgrad: g IS the 64-row operand (no activation derivative, no argmax routing, d = 0 only)
};

__global__ void __launch_bounds__(CB_THREADS, 1)
    c1s2_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const CbParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_bytes = 256u * 128u;                              // Wk2: four [64 k][64 co] tiles
  const uint32_t st_off = w_bytes;                                   // stages: [G' tiles d=0..3][patch tile]
  const uint32_t patch_off = st_off + 2 * CB_STAGE_BYTES;
  const uint32_t ctrl_off = patch_off + 2 * C1_PATCH_WORDS * 4;
  const uint32_t ctrl = base + ctrl_off;
  const uint32_t w_full = ctrl;
  auto full = [&](int s) { return ctrl + 8u * (1 + s); };
  auto empty = [&](int s) { return ctrl + 8u * (3 + s); };
  auto u_full = [&](int a) { return ctrl + 8u * (5 + a); };
  auto u_empty = [&](int a) { return ctrl + 8u * (7 + a); };
  const uint32_t dw_full = ctrl + 8u * 9;
  const uint32_t tmem_slot = ctrl + 8u * 10;
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(gbase + ctrl_off + 80);
  uint32_t* patch = (uint32_t*)(gbase + patch_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; s++) {
      mbar_init(full(s), 128);
      mbar_init(empty(s), 1);
      mbar_init(u_full(s), 1);
      mbar_init(u_empty(s), 4);
    }
    mbar_init(dw_full, 1);
    mbar_fence_init();
    prefetch_tensormap(&tmW);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (warp >= 1 && warp <= 8) {          // zero the stages once (patch-tile chunks 5..7 are never written afterwards)
    uint4* a4 = reinterpret_cast<uint4*>(gbase + st_off);
    for (int i = threadIdx.x - 32; i < 2 * CB_STAGE_BYTES / 16; i += 256) a4[i] = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  const int per_img = p.tiles_x * p.tiles_y;
  const int my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      if (p.want_u) {
        mbar_expect_tx(w_full, w_bytes);
        tma_load_2d(&tmW, base, w_full, 0, 0);
        mbar_wait(w_full, 0);
      }
      const uint32_t id_mn = idesc_f16(64, 1, 1), id_k = idesc_f16(64);
      for (int it = 0; it < my_tiles; it++) {
        const int s = it & 1, n = it >> 1;
        const uint32_t g_addr = base + st_off + s * CB_STAGE_BYTES, a_addr = g_addr + CB_G_BYTES;
        mbar_wait(full(s), n & 1);
        tc_fence_after();
        if (p.want_dw) {
          const int nh = p.plain ? 1 : 2;                  // plain: rows 0..63 = the operand, rows 64..127 stay zero
#pragma unroll 1
          for (int h = 0; h < nh; h++)
#pragma unroll
            for (int kk = 0; kk < 8; kk++)                 // 16 windows (two 8-row swizzle atoms) per MMA
              tc_mma_f16(tmem_base + h * 64, desc_mn_sw128(g_addr + h * 2 * C1_A_BYTES + kk * 2048, C1_A_BYTES),
                         desc_mn_sw128(a_addr + kk * 2048, C1_A_BYTES), id_mn, (it > 0 || kk > 0) ? 1u : 0u);
        }
        if (p.want_u) {
          const int acc = it & 1;
          mbar_wait(u_empty(acc), (n & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int d = 0; d < 4; d++)
#pragma unroll
            for (int k = 0; k < 4; k++)
              tc_mma_f16(tmem_base + 128 + acc * 64, desc_k_sw128(g_addr + d * C1_A_BYTES) + 2 * k,
                         desc_k_sw128(base + d * 8192) + 2 * k, id_k, (d | k) != 0);
          tc_commit(empty(s));
          tc_commit(u_full(acc));
        } else {
          tc_commit(empty(s));
        }
      }
      if (p.want_dw) tc_commit(dw_full);
    }
  } else if (warp <= 8) {
    // ===================== builders: group 0 = warps 1..4 (even tiles), group 1 = warps 5..8 (odd tiles) ==========
    const int grp = (warp - 1) >> 2;
    const int tb = threadIdx.x - 32 - grp * 128;
    const int iy = tb / p.bw, ix = tb - iy * p.bw;
    const int PWW = p.bw + 2, n_words = (2 * p.bh + 4) * PWW;
    uint32_t* pb = patch + grp * C1_PATCH_WORDS;
    uint8_t* stg = gbase + st_off + grp * CB_STAGE_BYTES;
    const float neg = p.act == HM_ACT_LRELU ? p.slope : (p.act == HM_ACT_RELU ? 0.f : 1.f);
    for (int it = grp; it < my_tiles; it += 2) {
      const int t = blockIdx.x + it * gridDim.x;
      const int b = t / per_img, rem = t - b * per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const float isc = p.img_scale ? p.img_scale[b] : 1.f;
      // issue every global load of this tile up front (24 gradient / activation / argmax loads + the patch words): one
      // round trip per tile instead of three.  The G' tiles are built chunk-wise -- (window r, 8-channel chunk cc) needs
      // only g, p, idx at that position -- so the loads are laid out for coalescing, not per window: in pass k a warp
      // covers 4 consecutive windows x 8 chunks = 512 contiguous bytes of g and of p (256 of idx), i.e. 4 + 4 + 2
      // 128-byte lines per three instructions.  (One thread per window row touched 32 lines per instruction; the LSU
      // transaction rate, not HBM, bounded the kernel: round-2 profile.)  Tried and reverted: issuing the NEXT tile's loads
      // pass by pass into the registers each pass of the build has just consumed (a tile in flight per group while it
      // builds) made all three uses 30-45 % slower (0.479 -> 0.644 ms): the 24 loads must go out back to back.
      const int sub = tb >> 3, cc = tb & 7;
      uint4 gv[8], pv[8];
      uint2 kv[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int r = k * 16 + sub;
        const int riy = r / p.bw, rix = r - riy * p.bw;
        const int rwy = ty * p.bh + riy, rwx = tx * p.bw + rix;
        if (rwy < p.Hq && rwx < p.Wq) {
          const size_t ro = (((size_t)b * p.Hq + rwy) * p.Wq + rwx) * 64 + cc * 8;
          gv[k] = *reinterpret_cast<const uint4*>(p.g + ro);
          if (!p.plain) {
            pv[k] = p.pl ? *reinterpret_cast<const uint4*>(p.pl + ro) : make_uint4(0, 0, 0, 0);
            kv[k] = *reinterpret_cast<const uint2*>(p.idx + ro);
          } else {
            pv[k] = make_uint4(0, 0, 0, 0);
            kv[k] = make_uint2(0, 0);
          }
        } else {
          gv[k] = pv[k] = make_uint4(0, 0, 0, 0);
          kv[k] = make_uint2(0, 0);
        }
      }
      uint32_t pre[C1_PRE];
      if (p.want_dw) {
        const int Y0 = 2 * ty * p.bh - 2, X0 = 2 * tx * p.bw - 2;
        const __half* img = p.x + (size_t)b * p.H * p.W;
#pragma unroll
        for (int j = 0; j < C1_PRE; j++) {
          const int i = tb + 128 * j;
          uint32_t v = 0;
          if (i < n_words) {
            const int pr = i / PWW, pw = i - pr * PWW;
            const int Y = Y0 + pr, X = X0 + 2 * pw;
            if (Y >= 0 && Y < p.H && X >= 0 && X < p.W) v = *reinterpret_cast<const uint32_t*>(img + (size_t)Y * p.W + X);
          }
          pre[j] = v;
        }
      }
      mbar_wait(empty(grp), ((it >> 1) & 1) ^ 1);
      if (p.want_dw) {
#pragma unroll
        for (int j = 0; j < C1_PRE; j++) {
          const int i = tb + 128 * j;
          if (i < n_words) pb[i] = pre[j];
        }
      }
      // g' = g * act'(p) * isc in half2 arithmetic: the factor is one of two constants selected by the sign of p
      const __half2 zero2 = __float2half2_rn(0.f);
      const uint32_t f_pos = __half_as_ushort(__float2half_rn(isc)) * 0x00010001u;
      const uint32_t f_neg = __half_as_ushort(__float2half_rn(neg * isc)) * 0x00010001u;
      if (p.plain) {
#pragma unroll
        for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(stg + sw128_off(c * 16 + sub, cc)) = gv[c];
      }
      if (!p.plain) {
#pragma unroll
      for (int c = 0; c < 8; c++) {         // pass c: window c*16 + sub, chunk cc
        const uint32_t gw[4] = {gv[c].x, gv[c].y, gv[c].z, gv[c].w}, pw4[4] = {pv[c].x, pv[c].y, pv[c].z, pv[c].w};
        // without the pooled tensor: bit 2 of the argmax bytes (written by hm_c1s2_conv) says which side of the
        // activation the stored value is on; bytes spread to 0xffff half lanes like the position compare below
        const uint32_t s0 = __vcmpeq4(kv[c].x & 0x04040404u, 0x04040404u), s1 = __vcmpeq4(kv[c].y & 0x04040404u, 0x04040404u);
        const uint32_t sm[4] = {__byte_perm(s0, 0, 0x1100), __byte_perm(s0, 0, 0x3322), __byte_perm(s1, 0, 0x1100),
                                __byte_perm(s1, 0, 0x3322)};
        uint32_t hv[4];                     // 8 halves
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const __half2 p2 = *reinterpret_cast<const __half2*>(&pw4[j]);
          const uint32_t pm = !p.pl ? sm[j]
                                    : (p.act == HM_ACT_RELU ? __hgt2_mask(p2, zero2) : __hge2_mask(p2, zero2));   // 0xffff per lane
          const uint32_t fb = (f_pos & pm) | (f_neg & ~pm);
          const __half2 h = __hmul2(*reinterpret_cast<const __half2*>(&gw[j]), *reinterpret_cast<const __half2*>(&fb));
          hv[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        // per window position d: keep the channels whose argmax byte equals d (byte compare, bytes spread to halves)
#pragma unroll
        for (int d = 0; d < 4; d++) {
          const uint32_t e0 = __vcmpeq4(kv[c].x & 0x03030303u, 0x01010101u * (uint32_t)d),
                         e1 = __vcmpeq4(kv[c].y & 0x03030303u, 0x01010101u * (uint32_t)d);
          const uint32_t m0 = hv[0] & __byte_perm(e0, 0, 0x1100), m1 = hv[1] & __byte_perm(e0, 0, 0x3322);
          const uint32_t m2 = hv[2] & __byte_perm(e1, 0, 0x1100), m3 = hv[3] & __byte_perm(e1, 0, 0x3322);
          *reinterpret_cast<uint4*>(stg + d * C1_A_BYTES + sw128_off(c * 16 + sub, cc)) = make_uint4(m0, m1, m2, m3);
        }
      }
      }
      if (p.want_dw) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        uint32_t wd[20];
#pragma unroll
        for (int u = 0; u < 6; u++)
#pragma unroll
          for (int j = 0; j < 3; j++) wd[u * 3 + j] = pb[(2 * iy + u) * PWW + ix + j];
        wd[18] = 0x00003C00u;                 // A[w][36] = 1: column 36 of the weight gradient is the bias gradient
        wd[19] = 0;
#pragma unroll
        for (int c = 0; c < 5; c++)
          *reinterpret_cast<uint4*>(stg + CB_G_BYTES + sw128_off(tb, c)) =
              make_uint4(wd[4 * c], wd[4 * c + 1], wd[4 * c + 2], wd[4 * c + 3]);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");      // patch buffer free for the next tile
      }
      fence_proxy_async();
      mbar_arrive(full(grp));
    }
  } else {
    // ===================== epilogue (warps 9..12) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int iy = row / p.bw, ix = row - iy * p.bw;
    if (p.want_u) {
      for (int it = 0; it < my_tiles; it++) {
        const int acc = it & 1;
        mbar_wait(u_full(acc), (it >> 1) & 1);
        tc_fence_after();
        const int t = blockIdx.x + it * gridDim.x;
        const int b = t / per_img, rem = t - b * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int wy = ty * p.bh + iy, wx = tx * p.bw + ix;
        const bool valid = wy < p.Hq && wx < p.Wq;
        const size_t o = (((size_t)b * p.Hq + wy) * p.Wq + wx) * 64;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 128 + acc * 64;
        uint32_t v[32], v2[16];
        tmem_ld32(taddr, v);
        tmem_ld16(taddr + 32, v2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(u_empty(acc));
        if (valid) {
          uint32_t pk[20];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            __half2 h = __floats2half2_rn(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            __half2 h = __floats2half2_rn(__uint_as_float(v2[j]), __uint_as_float(v2[j + 1]));
            pk[16 + (j >> 1)] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* d4 = reinterpret_cast<uint4*>(p.u + o);
#pragma unroll
          for (int j = 0; j < 5; j++) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      }
    }
    if (p.want_dw && my_tiles > 0 && !(p.plain && q >= 2)) {      // plain: rows 0..63 of the first accumulator only
      mbar_wait(dw_full, 0);
      tc_fence_after();
      const int nh = p.plain ? 1 : 2;
#pragma unroll 1
      for (int h = 0; h < nh; h++) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * 64;
        uint32_t v[32], v2[16];
        tmem_ld32(taddr, v);
        tmem_ld16(taddr + 32, v2);
        tmem_ld_wait();
        float* dst = p.dwk + (size_t)(h * 128 + row) * 64;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                       "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                       : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 32), "f"(__uint_as_float(v2[0])),
                     "f"(__uint_as_float(v2[1])), "f"(__uint_as_float(v2[2])), "f"(__uint_as_float(v2[3]))
                     : "memory");
        atomicAdd(dst + 36, __uint_as_float(v2[4]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// dwk[(d,co)][k] -> master gradient dW[co][0][a][b] (adjoint of pack mode 15) and db[co] (column 36)
__global__ void c1s2_bwd_fold_kernel(const float* __restrict__ dwk, float* __restrict__ dw, float* __restrict__ db,
                                     int cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cout * 25) {
    const int co = i / 25, a = (i % 25) / 5, bb = i % 5;
    const int r = 4 - a, s = 4 - bb;
    float acc = 0.f;
    for (int d = 0; d < 4; d++) acc += dwk[(size_t)(d * cout + co) * 64 + (r + (d >> 1)) * 6 + (s + (d & 1))];
    dw[i] = acc;
  } else if (i < cout * 26 && db) {
    const int co = i - cout * 25;
    float acc = 0.f;
    for (int d = 0; d < 4; d++) acc += dwk[(size_t)(d * cout + co) * 64 + 36];
    db[co] = acc;
  }
}

// dx[b][Y][X] = sum over the (up to 9) windows whose 6x6 patch covers pixel (Y,X) of U[window][patch position];
// one thread per window writes its 2x2 pixels.
__global__ void c1s2_col2im_kernel(const __half* __restrict__ u, __half* __restrict__ dx, int B, int Hq, int Wq) {
  const long long n = (long long)B * Hq * Wq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int wx = (int)(i % Wq);
    long long t = i / Wq;
    const int wy = (int)(t % Hq);
    const int b = (int)(t / Hq);
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int yy = wy + 1 - a;
      if (yy < 0 || yy >= Hq) continue;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const int xx = wx + 1 - c;
        if (xx < 0 || xx >= Wq) continue;
        const __half* row = u + (((size_t)b * Hq + yy) * Wq + xx) * 64;
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(row + (2 * a) * 6 + 2 * c));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(row + (2 * a + 1) * 6 + 2 * c));
        s00 += lo.x; s01 += lo.y; s10 += hi.x; s11 += hi.y;
      }
    }
    const int W = 2 * Wq;
    __half* o = dx + ((size_t)b * 2 * Hq + 2 * wy) * W + 2 * wx;
    *reinterpret_cast<__half2*>(o) = __floats2half2_rn(s00, s01);
    *reinterpret_cast<__half2*>(o + W) = __floats2half2_rn(s10, s11);
  }
}

}  // namespace hm

// geometry, the (optional) weight tensor map and the launch shared by hm_c1s2_bwd and hm_c1s2_wgrad
static int c1s2_bwd_launch(hm::CbParams& p, int B, int H, int W, const void* wk2, void* stream, const char* who) {
  using namespace hm;
  p.B = B; p.H = H; p.W = W; p.Hq = H / 2; p.Wq = W / 2;
  int bw = 1;
  while (bw * 2 <= p.Wq && bw < 128) bw *= 2;
  p.bw = bw; p.bh = 128 / bw;
  p.tiles_x = (p.Wq + p.bw - 1) / p.bw;
  p.tiles_y = (p.Hq + p.bh - 1) / p.bh;
  p.n_tiles = B * p.tiles_x * p.tiles_y;
  CUtensorMap tmW;
  {
    cuuint64_t dims[2] = {64, 256};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 256};
    cuuint32_t es[2] = {1, 1};
    void* wp = const_cast<void*>(wk2 ? wk2 : (const void*)p.g);     // unused (never dereferenced) without the input gradient
    CUresult r = c1_encode_fn()(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wp, dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d)", who, (int)r);
      return HM_ERR_CUDA;
    }
  }
  const size_t smem = 1024 + 256 * 128 + 2 * CB_STAGE_BYTES + 2 * C1_PATCH_WORDS * 4 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(c1s2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cannot raise dynamic shared memory: %s", who, cudaGetErrorString(e));
      return HM_ERR_CUDA;
    }
    attr = true;
  }
  int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  c1s2_bwd_kernel<<<grid, CB_THREADS, smem, (cudaStream_t)stream>>>(tmW, p);
  HM_CHECK_LAUNCH(who);
  return HM_OK;
}

// Backward of hm_c1s2_conv's pooled form from the gradient g of the pooled tensor (see the kernel comment).
//   dwk != NULL: dwk[256][64] fp32 += weight-gradient partials (caller zeroes; fold with hm_c1s2_bwd_fold);
//   u   != NULL: u[B,H/2,W/2,64] = patch-space input gradient (needs wk2 = pack mode 16; fold with hm_c1s2_col2im);
//   img_scale (optional, [B]): g of image b is multiplied by img_scale[b] first.
extern "C" int hm_c1s2_bwd(const void* x, const void* g, const void* pooled, const uint8_t* idx, const void* wk2,
                           float* dwk, void* u, const float* img_scale, int B, int H, int W, int act, float slope,
                           void* stream) {
  HM_CHECK_ARG(g && idx && B > 0 && H > 0 && W > 0 && (dwk || u), "hm_c1s2_bwd: bad argument");
  HM_CHECK_ARG(!dwk || x, "hm_c1s2_bwd: the weight gradient needs the source image");
  HM_CHECK_ARG(!u || wk2, "hm_c1s2_bwd: the input gradient needs the weights (pack mode 16)");
  HM_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "hm_c1s2_bwd: the image must have even height and width (%dx%d)", H, W);
  HM_CHECK_ARG(act == HM_ACT_LINEAR || act == HM_ACT_LRELU || act == HM_ACT_RELU, "hm_c1s2_bwd: activation %d is not supported", act);
  if ((((uintptr_t)x) & 3) || (((uintptr_t)g | (uintptr_t)pooled | (uintptr_t)wk2 | (uintptr_t)u) & 15) || (((uintptr_t)idx) & 7)) {
    set_error("hm_c1s2_bwd: x must be 4-byte, idx 8-byte, g / pooled / wk2 / u 16-byte aligned");
    return HM_ERR_ALIGN;
  }
  if (!c1_encode_fn()) {
    set_error("hm_c1s2_bwd: cuTensorMapEncodeTiled is not available from this driver");
    return HM_ERR_CUDA;
  }
  CbParams p;
  p.act = act; p.slope = slope; p.want_dw = dwk ? 1 : 0; p.want_u = u ? 1 : 0;
  p.x = (const __half*)x; p.g = (const __half*)g; p.pl = (const __half*)pooled; p.idx = idx; p.dwk = dwk; p.u = (__half*)u;
  p.img_scale = img_scale;
  p.plain = 0;
  return c1s2_bwd_launch(p, B, H, W, wk2, stream, "hm_c1s2_bwd");
}

// Weight gradient of (nearest-2x -> conv5x5 'same', 64 -> 1 channel) from the one-channel dy: the same kernel with the
// roles swapped -- the 6x6 stride-2 patches are gathered from dy, the 64-row operand is the low-res source itself:
//   dwk[ci][u*6+v] += sum_q x[q][ci] * dy[2q-2+(u,v)]      (the gradient of pack mode 14's Wk; unpack mode 14 folds it
// onto the 5x5 filter).  x is read once (128 B per source pixel) and nothing else is materialised.
extern "C" int hm_c1s2_wgrad(const void* dy, const void* x, float* dwk, int B, int H, int W, void* stream) {
  HM_CHECK_ARG(dy && x && dwk && B > 0 && H > 0 && W > 0, "hm_c1s2_wgrad: bad argument");
  HM_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "hm_c1s2_wgrad: dy must have even height and width (%dx%d)", H, W);
  if ((((uintptr_t)dy) & 3) || (((uintptr_t)x) & 15)) {
    set_error("hm_c1s2_wgrad: dy must be 4-byte, x 16-byte aligned");
    return HM_ERR_ALIGN;
  }
  if (!c1_encode_fn()) {
    set_error("hm_c1s2_wgrad: cuTensorMapEncodeTiled is not available from this driver");
    return HM_ERR_CUDA;
  }
  CbParams p;
  p.act = HM_ACT_LINEAR; p.slope = 0.f; p.want_dw = 1; p.want_u = 0;
  p.x = (const __half*)dy; p.g = (const __half*)x; p.pl = nullptr; p.idx = nullptr; p.dwk = dwk; p.u = nullptr;
  p.img_scale = nullptr;
  p.plain = 1;
  return c1s2_bwd_launch(p, B, H, W, nullptr, stream, "hm_c1s2_wgrad");
}

extern "C" int hm_c1s2_bwd_fold(const float* dwk, float* dw, float* db, int cout, void* stream) {
  HM_CHECK_ARG(dwk && dw && cout == 64, "hm_c1s2_bwd_fold: bad argument (cout must be 64)");
  c1s2_bwd_fold_kernel<<<(cout * 26 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dwk, dw, db, cout);
  HM_CHECK_LAUNCH("hm_c1s2_bwd_fold");
  return HM_OK;
}

extern "C" int hm_c1s2_col2im(const void* u, void* dx, int B, int H, int W, void* stream) {
  HM_CHECK_ARG(u && dx && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "hm_c1s2_col2im: bad argument");
  HM_CHECK_ARG((((uintptr_t)u) & 15) == 0 && (((uintptr_t)dx) & 3) == 0, "hm_c1s2_col2im: misaligned pointer");
  const long long n = (long long)B * (H / 2) * (W / 2);
  long long blocks = (n + 255) / 256, cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  c1s2_col2im_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)u, (__half*)dx, B, H / 2, W / 2);
  HM_CHECK_LAUNCH("hm_c1s2_col2im");
  return HM_OK;
}
