// tcgen05 implicit-GEMM convolution for sm_100a (fast mode: fp16 operands, fp32 accumulate in TMEM).
//
// Stands in for the cuDNN/CorrMM convolution that Theano dispatches for lasagne Conv2DLayer
// (reference architectures/dcgan.py:22,42; architectures/p2p.py:20-21,208-209) and, with the
// rotated weight pack, for its input gradient (CorrMM_gradInputs).
//
// GEMM view:  D[M = 128 output pixels][N = Cout tile] += A[M][K] * B[N][K],  K = taps x Cin.
//   * A is never materialised (no im2col buffer): for every filter tap (r,s) and every 64-channel
//     slice, ONE tiled TMA load of the box {64 ch, bw, bh, bn} (bw*bh*bn = 128 pixels) at the
//     tap-shifted coordinate lands the 128x64 operand tile in shared memory in the 128B-swizzled
//     K-major layout tcgen05 reads; rows that fall into the zero padding are zero-filled by the
//     TMA unit itself (out-of-bounds box elements, negative coordinates included).
//   * B is the per-tap weight slice [Cout tile][64 ch] of the pack [tap][Cout][Cin] (K-major).
//   * One elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N<=256, K=16 x4 per
//     stage); tcgen05.commit releases the shared-memory stage and, at the end of a tile, publishes
//     the TMEM accumulator to the four epilogue warps.  Two accumulators (2 x N columns) let the
//     epilogue of tile i overlap the MMAs of tile i+1.  CTAs are persistent over the tile list.
//   * Epilogue: tcgen05.ld 32 lanes x 32 columns -> bias, activation -> fp16 -> 16-byte stores
//     into the NHWC output.
#include <cuda.h>
#include <stdlib.h>

#include "hm_common.cuh"

namespace hm {

constexpr int TC_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue
constexpr int TILE_M = 128;
constexpr int KCH = 64;                      // channels per pipeline stage = one 128-byte swizzle row
constexpr int A_BYTES = TILE_M * KCH * 2;    // 16 KB

struct TcParams {
  int B, Ho, Wo;           // output grid
  int Cin, C1, Cout;       // K per tap (C1 from source 1, Cin-C1 from source 2), total N
  int kh, kw, pad;
  int stride;              // 1, or 2 (forward only: the A box is loaded with TMA element strides of 2)
  int bw, bh, bn;          // pixel box of one M tile
  int tiles_x, tiles_y, tiles_n, n_mtiles;
  int ntile, n_ntiles;     // N tile (<=256, multiple of 16)
  int S, n_super;          // S consecutive M tiles share every B (weight) stage: S*ntile*2 <= 512 TMEM columns
  int stages;
  int a_slots, b_slots, rb_bytes;   // row-box variant: A ring (one box per filter row), B ring (one slot per tap)
  int act;
  float slope;
  const float* bias;       // [Cout] or null
  __half* y;               // [B,Ho,Wo,split]       channels [0,split)
  __half* y2;              // [B,Ho,Wo,Cout-split]  channels [split,Cout)   (concat sources of an input gradient)
  int split, accumulate;   // accumulate bit0: y += , bit1: y2 +=
  int vec_store;           // 1: 16-byte stores (Cout, split multiples of 8); 0: scalar stores of the real columns
  int d2s;                 // 1: nearest-2x + 5x5 as four 3x3 phase convolutions: N = (phase, co), the epilogue scatters
                           //    phase (py,px) of low-res pixel (qy,qx) to output pixel (2qy+py, 2qx+px)
  int cph;                 // channels per phase (= real Cout) when d2s
  int pool;                // row-box kernel, S == 2: 2x2 max-pool fused into the epilogue; the two sub-tiles of a super tile
                           // are the image rows 2yp, 2yp+1 of one 128-pixel segment, y is the POOLED tensor
  uint8_t* idx;            // [B, Ho/2, Wo/2, Cout] argmax bytes (pool)
  int bf16;                // operands are bfloat16 (HM_BF16X3: hi/lo splits of fp32 tensors) instead of fp16
  int out32;               // y / y2 are float tensors (HM_BF16X3): the fp32 accumulator is stored unrounded
};

// ---- PTX wrappers ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// true in exactly one lane of a converged warp.  The role loops below run WARP-UNIFORMLY (all 32 lanes wait on the
// barriers and compute the same descriptors, which therefore live in uniform registers) and only the instructions that
// must be issued once sit under elect_one(): nvcc then emits ONE predicated UTCHMMA / UTMALDG.  Issued from inside an
// `if (lane == 0)` region the same inline asm compiles to an ELECT / BRA.U.ANY waterfall loop per instruction, which
// made the single issuing thread the bottleneck (134 cycles per M=128,N=128,K=16 MMA instead of 64; tools/mma_rate.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address      bits [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset  bits [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: A,B = F16 (K-major), D = F32, M = 128, N = n.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
// the same with A,B = BF16 (a_format, b_format = 1 at bits 7 and 10) when bf16 != 0
__host__ __device__ constexpr uint32_t umma_idesc_16(int n, int bf16) {
  return umma_idesc_f16(n) | (bf16 ? ((1u << 7) | (1u << 10)) : 0u);
}

template <int ACT>
__device__ __forceinline__ float act_t(float v, float slope) {
  if (ACT == HM_ACT_LRELU) return fmaxf(v, 0.f) + slope * fminf(v, 0.f);
  if (ACT == HM_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == HM_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  if (ACT == HM_ACT_TANH) return tanhf(v);
  return v;
}

// 32 accumulator columns -> +bias (shared memory, float4 reads) -> activation -> 16 packed half2
template <int ACT>
__device__ __forceinline__ void epi_pack32(const uint32_t* v, const float* bias32, float slope, uint32_t* packed) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias32 + j);
    const float a0 = act_t<ACT>(__uint_as_float(v[j]) + b.x, slope);
    const float a1 = act_t<ACT>(__uint_as_float(v[j + 1]) + b.y, slope);
    const float a2 = act_t<ACT>(__uint_as_float(v[j + 2]) + b.z, slope);
    const float a3 = act_t<ACT>(__uint_as_float(v[j + 3]) + b.w, slope);
    __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(a2, a3);
    packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h0);
    packed[(j >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
  }
}

// M tile -> (x segment, image row block, image block).  Pooled layers enumerate their tiles so that the consecutive
// tiles 2k, 2k+1 -- the two sub-tiles of a super tile (S == 2) -- are the rows 2yp and 2yp+1 of the same segment.
__device__ __forceinline__ void mtile_coords(const TcParams& p, int mt, int& tx, int& ty, int& tn) {
  if (p.pool) {
    const int i = mt & 1, pt = mt >> 1, hy = p.tiles_y >> 1;
    tx = pt % p.tiles_x;
    ty = 2 * ((pt / p.tiles_x) % hy) + i;
    tn = pt / (p.tiles_x * hy);
  } else {
    tx = mt % p.tiles_x;
    ty = (mt / p.tiles_x) % p.tiles_y;
    tn = mt / (p.tiles_x * p.tiles_y);
  }
}

// HM_BF16X3 epilogue: up to 32 accumulator columns -> (+ stored value) + bias -> activation -> FLOAT stores (the fp32
// accumulator is kept unrounded: the operands were bf16 hi/lo splits of fp32 tensors, see hm_split_bf16x3)
template <int ACT>
__device__ __forceinline__ void epi_store32_f32(float* dst, const uint32_t* v, const float* bias32, float slope,
                                                int ncols, bool accum, bool vec) {
  if (vec) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < ncols) {
        const float4 b = *reinterpret_cast<const float4*>(bias32 + j);
        float4 o = make_float4(__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y,
                               __uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w);
        if (accum) {
          const float4 pv = *reinterpret_cast<const float4*>(dst + j);
          o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
        }
        o.x = act_fwd(o.x, ACT, slope); o.y = act_fwd(o.y, ACT, slope);
        o.z = act_fwd(o.z, ACT, slope); o.w = act_fwd(o.w, ACT, slope);
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j++) {
      if (j < ncols) {
        float a = __uint_as_float(v[j]) + bias32[j];
        if (accum) a += dst[j];
        dst[j] = act_fwd(a, ACT, slope);
      }
    }
  }
}

// depth-to-space store of a thin phase-decomposed layer: columns (phase, co), co < CPH <= 4; phase (py, px) of low-res
// pixel (oy, ox) goes to output pixel (2oy+py, 2ox+px).  row0 = pixel index of (2oy, 2ox) in the [B, 2Ho, 2Wo] grid.
template <int ACT, int CPH>
__device__ __forceinline__ void d2s_thin_store(const TcParams& p, const uint32_t* v, const float* bias_t, size_t row0) {
#pragma unroll
  for (int py = 0; py < 2; py++) {
    __half* dst = p.y + (row0 + (size_t)py * (2 * p.Wo)) * CPH;       // pixels (2oy+py, 2ox) and (2oy+py, 2ox+1): 2*CPH halves
    float a[2 * CPH];
#pragma unroll
    for (int px = 0; px < 2; px++)
#pragma unroll
      for (int co = 0; co < CPH; co++) {
        const int col = (py * 2 + px) * CPH + co;
        a[px * CPH + co] = __uint_as_float(v[col]) + bias_t[col];
        if (p.accumulate & 1) a[px * CPH + co] += __half2float(dst[px * CPH + co]);
        a[px * CPH + co] = act_t<ACT>(a[px * CPH + co], p.slope);
      }
    if (CPH % 2 == 0 || CPH == 1) {            // 4-byte aligned pairs (2ox is even): half2 stores
#pragma unroll
      for (int e = 0; e < 2 * CPH; e += 2)
        *reinterpret_cast<__half2*>(dst + e) = __floats2half2_rn(a[e], a[e + 1]);
    } else {
#pragma unroll
      for (int e = 0; e < 2 * CPH; e++) dst[e] = __float2half_rn(a[e]);
    }
  }
}

// The whole epilogue role: for every tile of this CTA (work items blockIdx.x, +gridDim.x, ...) wait for the accumulator,
// drain it, release it.
template <int ACT, bool OUT32 = false>
__device__ __forceinline__ void epilogue_loop(const TcParams& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                              const float* bias_s, int warp, int lane, int total_tiles) {
  const int q = warp & 3;                                   // TMEM lane quarter this warp may access
  const int row = q * 32 + lane;
  const int px_per_img = p.bw * p.bh;
  const int in = row / px_per_img;
  const int rem = row - in * px_per_img;
  const int iy = rem / p.bw, ix = rem - iy * p.bw;
  int it = 0;
  for (int t = (int)blockIdx.x; t < total_tiles; t += (int)gridDim.x, it++) {
    const int st = t / p.n_ntiles, nt = t - st * p.n_ntiles;
    const int mt0 = st * p.S;
    const int nv = max(0, min(p.S, p.n_mtiles - mt0));
    const int acc = it & 1;
    mbar_wait(tfull0 + 8u * acc, (it >> 1) & 1);
    tc_fence_after();
    const int col0 = nt * p.ntile;
    const float* bias_t = bias_s + col0;
    if (!OUT32 && p.pool) {
      // ---- conv + bias + activation + 2x2 max-pool: lane = pixel x of the rows 2yp (accumulator 0) and 2yp+1
      // (accumulator 1); vertical max in registers, horizontal max with the neighbouring lane; the even lanes write the
      // pooled pixel and its argmax (d = 2*dy + dx, first maximum wins, as hm_maxpool2_fwd).  The activation is
      // monotonic (host check), so it commutes with the max.
      int tx, ty, tn;
      mtile_coords(p, mt0, tx, ty, tn);
      const int ox = tx * p.bw + ix;
      const bool wr = !(lane & 1) && tn < p.B && ox < p.Wo;
      const uint32_t ta0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (p.S * p.ntile), ta1 = ta0 + p.ntile;
      const size_t pp = ((size_t)tn * (p.Ho >> 1) + (ty >> 1)) * (size_t)(p.Wo >> 1) + (ox >> 1);
      __half* yo = p.y + pp * p.Cout + col0;
      uint8_t* io = p.idx + pp * p.Cout + col0;
      for (int c0 = 0; c0 < p.ntile; c0 += 32) {
        uint32_t v0[32], v1[32];
        tmem_ld32(ta0 + c0, v0);
        tmem_ld32(ta1 + c0, v1);
        tmem_ld_wait();
        uint32_t packed[16], kb[8];
#pragma unroll
        for (int j = 0; j < 8; j++) kb[j] = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float r2[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const float a = __uint_as_float(v0[j + e]), b = __uint_as_float(v1[j + e]);
            float m = b > a ? b : a;
            uint32_t k = b > a ? 2u : 0u;
            const float pm = __shfl_xor_sync(0xffffffffu, m, 1);
            const uint32_t pk = __shfl_xor_sync(0xffffffffu, k, 1) + 1u;          // the odd lane is dx = 1
            if (pm > m || (pm == m && pk < k)) { m = pm; k = pk; }
            r2[e] = act_t<ACT>(m + bias_t[c0 + j + e], p.slope);
            kb[(j + e) >> 2] |= k << (8 * ((j + e) & 3));
          }
          __half2 h = __floats2half2_rn(r2[0], r2[1]);
          packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (wr) {
          uint4* d4 = reinterpret_cast<uint4*>(yo + c0);
#pragma unroll
          for (int j = 0; j < 4; j++) d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          uint4* i4 = reinterpret_cast<uint4*>(io + c0);
          i4[0] = make_uint4(kb[0], kb[1], kb[2], kb[3]);
          i4[1] = make_uint4(kb[4], kb[5], kb[6], kb[7]);
        }
      }
    } else
   for (int sub = 0; sub < nv; sub++) {
    const int mt = mt0 + sub;
    int tx, ty, tn;
    mtile_coords(p, mt, tx, ty, tn);
    const int n = tn * p.bn + in, oy = ty * p.bh + iy, ox = tx * p.bw + ix;
    const bool valid = n < p.B && oy < p.Ho && ox < p.Wo;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (p.S * p.ntile) + sub * p.ntile;
    if (p.d2s) {
      // ---- depth-to-space: column = (phase, co); phase (py,px) of low-res pixel (oy,ox) -> (2oy+py, 2ox+px)
      for (int c0 = 0; c0 < p.ntile; c0 += 32) {
        uint32_t v[32];
        if (p.ntile - c0 >= 32) {
          tmem_ld32(taddr + c0, v);
        } else {
          tmem_ld16(taddr + c0, v);
#pragma unroll
          for (int j = 16; j < 32; j++) v[j] = 0;
        }
        tmem_ld_wait();
        if (!valid) continue;
        const int gc = col0 + c0;
        if (OUT32) {
          float* y32 = reinterpret_cast<float*>(p.y);
          if (p.cph % 32 == 0) {
            const int ph = gc / p.cph, co = gc - ph * p.cph;
            const size_t op = ((size_t)((size_t)n * 2 * p.Ho + 2 * oy + (ph >> 1)) * (2 * p.Wo) + 2 * ox + (ph & 1));
            epi_store32_f32<ACT>(y32 + op * p.cph + co, v, bias_t + c0, p.slope, min(32, p.ntile - c0),
                                 (p.accumulate & 1) != 0, true);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const int col = gc + j;
              if (col < 4 * p.cph) {
                const int ph = col / p.cph, co = col - ph * p.cph;
                const size_t op = ((size_t)((size_t)n * 2 * p.Ho + 2 * oy + (ph >> 1)) * (2 * p.Wo) + 2 * ox + (ph & 1));
                float a = __uint_as_float(v[j]) + bias_t[c0 + j];
                if (p.accumulate & 1) a += y32[op * p.cph + co];
                y32[op * p.cph + co] = act_fwd(a, ACT, p.slope);
              }
            }
          }
          continue;
        }
        if (p.cph % 32 == 0) {
          const int ph = gc / p.cph, co = gc - ph * p.cph;
          const size_t op = ((size_t)((size_t)n * 2 * p.Ho + 2 * oy + (ph >> 1)) * (2 * p.Wo) + 2 * ox + (ph & 1));
          uint32_t packed[16];
          uint4* d4 = reinterpret_cast<uint4*>(p.y + op * p.cph + co);
          if (p.accumulate & 1) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const uint4 pv = d4[j];
              const uint32_t w4[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&w4[k]));
                v[j * 8 + 2 * k] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * k]) + pf.x);
                v[j * 8 + 2 * k + 1] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * k + 1]) + pf.y);
              }
            }
          }
          epi_pack32<ACT>(v, bias_t + c0, p.slope, packed);
#pragma unroll
          for (int j = 0; j < 4; j++)
            d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        } else if (p.cph <= 4 && gc == 0) {
          // thin layers (1..4 channels per phase: the generator's last layer, the U-Net's output deconvolution): the
          // 4*cph real columns with compile-time phase / channel indices.  (The generic loop below -- 32 predicated
          // iterations with a runtime division each -- WAS the run time of the generator's last layer: 0.35 ms for a
          // 268 MB read, tensor pipe 5 % busy, stall samples spread over its 32 reconvergence points; round-2 profile.)
          const size_t row0 = ((size_t)n * 2 * p.Ho + 2 * oy) * (size_t)(2 * p.Wo) + 2 * ox;
          switch (p.cph) {
            case 1: d2s_thin_store<ACT, 1>(p, v, bias_t, row0); break;
            case 2: d2s_thin_store<ACT, 2>(p, v, bias_t, row0); break;
            case 3: d2s_thin_store<ACT, 3>(p, v, bias_t, row0); break;
            default: d2s_thin_store<ACT, 4>(p, v, bias_t, row0); break;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int col = gc + j;
            if (col < 4 * p.cph) {
              const int ph = col / p.cph, co = col - ph * p.cph;
              const size_t op = ((size_t)((size_t)n * 2 * p.Ho + 2 * oy + (ph >> 1)) * (2 * p.Wo) + 2 * ox + (ph & 1));
              float a = __uint_as_float(v[j]) + bias_t[c0 + j];
              if (p.accumulate & 1) a += __half2float(p.y[op * p.cph + co]);
              p.y[op * p.cph + co] = __float2half_rn(act_t<ACT>(a, p.slope));
            }
          }
        }
      }
    } else {
      const size_t pix = (size_t)((size_t)n * p.Ho + oy) * p.Wo + ox;
      const bool second = col0 >= p.split;
      __half* dst = second ? p.y2 + pix * (p.Cout - p.split) + (col0 - p.split) : p.y + pix * p.split + col0;
      const bool accum = (p.accumulate & (second ? 2 : 1)) != 0;
      const bool store = valid && (second ? p.y2 != nullptr : p.y != nullptr);
      for (int c0 = 0; c0 < p.ntile; c0 += 32) {
        uint32_t v[32];
        if (p.ntile - c0 >= 32) {
          tmem_ld32(taddr + c0, v);
        } else {
          tmem_ld16(taddr + c0, v);
#pragma unroll
          for (int j = 16; j < 32; j++) v[j] = 0;
        }
        tmem_ld_wait();
        if (!store) continue;
        const int ncols = min(32, p.ntile - c0);
        if (OUT32) {
          float* d32 = second ? reinterpret_cast<float*>(p.y2) + pix * (p.Cout - p.split) + (col0 - p.split)
                              : reinterpret_cast<float*>(p.y) + pix * p.split + col0;
          epi_store32_f32<ACT>(d32 + c0, v, bias_t + c0, p.slope, p.vec_store ? ncols : min(ncols, p.Cout - col0 - c0),
                               accum, p.vec_store != 0);
          continue;
        }
        if (p.vec_store) {
          uint32_t packed[16];
          if (accum) {                                        // y += : add the stored fp16 values in fp32 first
#pragma unroll
            for (int j = 0; j < 4; j++) {
              if (j * 8 < ncols) {
                const uint4 pv = reinterpret_cast<const uint4*>(dst + c0)[j];
                const uint32_t w4[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&w4[k]));
                  v[j * 8 + 2 * k] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * k]) + pf.x);
                  v[j * 8 + 2 * k + 1] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * k + 1]) + pf.y);
                }
              }
            }
          }
          epi_pack32<ACT>(v, bias_t + c0, p.slope, packed);
          uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (j * 8 < ncols) d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        } else {  // thin outputs: the N tile is zero-padded by TMA, store only the real columns
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (col0 + c0 + j < p.Cout) {
              float a = __uint_as_float(v[j]) + bias_t[c0 + j];
              if (accum) a += __half2float(dst[c0 + j]);
              dst[c0 + j] = __float2half_rn(act_t<ACT>(a, p.slope));
            }
          }
        }
      }
    }
   }
    tc_fence_before();
    __syncwarp();
    if (elect_one()) mbar_arrive(tempty0 + 8u * acc);
  }
}

// activation / output-type dispatch of the epilogue role
__device__ __forceinline__ void epilogue_dispatch(const TcParams& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                                  const float* bias_s, int warp, int lane, int total_tiles) {
#define HM_EPI(A_, O_) epilogue_loop<A_, O_>(p, tmem_base, tfull0, tempty0, bias_s, warp, lane, total_tiles)
  if (p.out32) {
    switch (p.act) {
      case HM_ACT_LRELU: HM_EPI(HM_ACT_LRELU, true); break;
      case HM_ACT_RELU: HM_EPI(HM_ACT_RELU, true); break;
      case HM_ACT_SIGMOID: HM_EPI(HM_ACT_SIGMOID, true); break;
      case HM_ACT_TANH: HM_EPI(HM_ACT_TANH, true); break;
      default: HM_EPI(HM_ACT_LINEAR, true); break;
    }
  } else {
    switch (p.act) {
      case HM_ACT_LRELU: HM_EPI(HM_ACT_LRELU, false); break;
      case HM_ACT_RELU: HM_EPI(HM_ACT_RELU, false); break;
      case HM_ACT_SIGMOID: HM_EPI(HM_ACT_SIGMOID, false); break;
      case HM_ACT_TANH: HM_EPI(HM_ACT_TANH, false); break;
      default: HM_EPI(HM_ACT_LINEAR, false); break;
    }
  }
#undef HM_EPI
}

// One filter tap of the row-box kernels for NS sub-tiles, fully unrolled: NS*4 back-to-back MMAs.  Descriptors are
// (constant high word, 32-bit low word); a_lo advances by a_sub per sub-tile, the accumulator column by ntile.
constexpr uint32_t UMMA_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_lo(uint32_t lo) { return ((uint64_t)UMMA_DESC_HI << 32) | (uint64_t)lo; }
__device__ __forceinline__ uint32_t umma_lo_of(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | (1u << 16); }
// (Issue order: sub-tile outer, K slice inner.  Interleaving the sub-tiles' accumulators -- K slice outer -- changed nothing,
// neither on the wide layers nor on the thin N = 16 one: measured in round 2.)
template <int NS>
__device__ __forceinline__ void mma_tap(uint32_t d_tmem, uint32_t a_lo, uint32_t a_sub, uint32_t b_lo, uint32_t ntile,
                                        uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int i = 0; i < NS; i++) {
#pragma unroll
    for (int k = 0; k < KCH / 16; k++)
      tc_mma_f16(d_tmem + i * ntile, umma_desc_lo(a_lo + i * a_sub + 2 * k), umma_desc_lo(b_lo + 2 * k), idesc,
                 acc_first | (uint32_t)k);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = p.S * A_BYTES;
  const uint32_t stage_bytes = a_bytes + p.ntile * 128;
  const uint32_t ctrl = base + p.stages * stage_bytes;        // 1024-aligned
  // control block: full[stages] | empty[stages] | tfull[2] | tempty[2] | tmem base address | (+1024) bias[<=1024]
  auto full_bar = [&](int s) { return ctrl + 8u * s; };
  auto empty_bar = [&](int s) { return ctrl + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return ctrl + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return ctrl + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = ctrl + 8u * (2 * p.stages + 4);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(gen_base + (tmem_slot - base));
  float* bias_s = (float*)(gen_base + (ctrl - base) + 1024);     // [n_ntiles * ntile] (<= 8192) floats
  {
    const int ncols = p.n_ntiles * p.ntile;
    const int creal = p.d2s ? p.cph : p.Cout;
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) {
      const int c = p.d2s ? i % p.cph : i;
      bias_s[i] = (p.bias && c < creal && (!p.d2s || i < 4 * p.cph)) ? p.bias[c] : 0.f;
    }
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_cols = p.S * p.ntile;                    // columns of one accumulator set (double-buffered)
  const int tmem_cols = 2 * acc_cols <= 128 ? 128 : (2 * acc_cols <= 256 ? 256 : 512);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;

  const int total_tiles = p.n_super * p.n_ntiles;        // work items: (super tile of S M tiles, N tile)
  const int ksteps_per_tap = p.Cin / KCH;
  const int taps = p.kh * p.kw;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int st = t / p.n_ntiles, nt = t - st * p.n_ntiles;
        const int nv = min(p.S, p.n_mtiles - st * p.S);      // M tiles present in this super tile
        for (int tap = 0; tap < taps; tap++) {
          const int r = tap / p.kw, s = tap - r * p.kw;
          for (int cc = 0; cc < ksteps_per_tap; cc++) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = base + stage * stage_bytes;
            mbar_expect_tx(full_bar(stage), nv * A_BYTES + p.ntile * 128);
            const int c = cc * KCH;
#pragma unroll 1
            for (int i = 0; i < nv; i++) {
              const int mt = st * p.S + i;
              const int ox = (mt % p.tiles_x) * p.bw * p.stride - p.pad + s;
              const int oy = ((mt / p.tiles_x) % p.tiles_y) * p.bh * p.stride - p.pad + r;
              const int on = (mt / (p.tiles_x * p.tiles_y)) * p.bn;
              if (c < p.C1)
                tma_load_4d(&tmA, sa + i * A_BYTES, full_bar(stage), c, ox, oy, on);
              else
                tma_load_4d(&tmA2, sa + i * A_BYTES, full_bar(stage), c - p.C1, ox, oy, on);
            }
            tma_load_3d(&tmB, sa + a_bytes, full_bar(stage), c, nt * p.ntile, tap);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      // single issuing lane; 32-bit descriptor arithmetic; the NEXT stage's barrier is polled right behind the MMAs of
      // the current one (the tensor pipe queues only a few MMAs: tools/mma_rate.cu)
      const uint32_t idesc = umma_idesc_16(p.ntile, p.bf16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int ksteps = taps * ksteps_per_tap;
      if ((int)blockIdx.x < total_tiles) mbar_wait(full_bar(0), 0);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        const int st = t / p.n_ntiles;
        const int nv = min(p.S, p.n_mtiles - st * p.S);
        const bool more_tiles = t + (int)gridDim.x < total_tiles;
        uint32_t first = 0;
        for (int ks = 0; ks < ksteps; ks++) {
          tc_fence_after();
          const uint32_t a_lo = umma_lo_of(base + stage * stage_bytes);
          const uint32_t b_lo = a_lo + (a_bytes >> 4);
          if (nv == 1) mma_tap<1>(d_tmem, a_lo, A_BYTES >> 4, b_lo, p.ntile, idesc, first);
          else if (nv == 2) mma_tap<2>(d_tmem, a_lo, A_BYTES >> 4, b_lo, p.ntile, idesc, first);
          else if (nv == 4) mma_tap<4>(d_tmem, a_lo, A_BYTES >> 4, b_lo, p.ntile, idesc, first);
          else mma_tap<3>(d_tmem, a_lo, A_BYTES >> 4, b_lo, p.ntile, idesc, first);
          first = 1;
          tc_commit(empty_bar(stage));                        // frees the stage when these MMAs retire
          if (ks == ksteps - 1) tc_commit(tfull_bar(acc));    // accumulator complete
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          if (ks + 1 < ksteps || more_tiles) mbar_wait(full_bar(stage), phase);
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    epilogue_dispatch(p, tmem_base, tfull_bar(0), tempty_bar(0), bias_s, warp, lane, total_tiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// NOTE (measured on B200, tools/tc_probe.py *_rb cases): the 128B swizzle is a function of the absolute shared-memory
// address, so a K-major SW128 descriptor may start at ANY 128-byte line of a TMA-written buffer (not only at a
// 1024-byte atom boundary) with base_offset = 0; encoding (addr >> 7) & 7 as base_offset gives wrong results.  The
// row-box kernels rely on this: tap s of a filter row reads the row's box from start + 128*s bytes.

// MMA-issuer role of the row-box kernel with the tap loop fully unrolled (KW taps per filter row, NS sub-tiles) for a
// weight ring whose length is a multiple of KW: the ring slot of tap s is (ring row)*KW + s, so every barrier address and
// descriptor inside a row is the row's base plus a compile-time constant.  Per tap the issuing lane then executes one
// barrier poll, NS*4 UTCHMMA and one commit; the poll for the NEXT tap (or next row's operands) sits behind the MMAs.
template <int KW, int NS>
__device__ __forceinline__ void rb_mma_role_unrolled(const TcParams& p, uint32_t base, uint32_t b_base, uint32_t ctrl,
                                                     uint32_t tmem_base, int total_tiles, int cchunks) {
  const uint32_t a_slot_bytes = p.S * p.rb_bytes, b_slot_bytes = p.ntile * 128;
  const uint32_t afull0 = ctrl, aempty0 = ctrl + 8u * p.a_slots, bfull0 = ctrl + 8u * (2 * p.a_slots),
                 bempty0 = ctrl + 8u * (2 * p.a_slots + p.b_slots), tfull0 = ctrl + 8u * (2 * p.a_slots + 2 * p.b_slots),
                 tempty0 = tfull0 + 16u;
  const uint32_t idesc = umma_idesc_16(p.ntile, p.bf16);
  const uint32_t a_sub = (uint32_t)p.rb_bytes >> 4, b_sub = b_slot_bytes >> 4;
  const uint32_t a_lo_base = umma_lo_of(base), b_lo_base = umma_lo_of(b_base);
  const int ring_rows = p.b_slots / KW;
  const int acc_cols = p.S * p.ntile;
  int as = 0, brow = 0, it = 0;
  uint32_t aph = 0, bph = 0;
  const int steps = p.kh * cchunks;                                   // A slots (filter row x channel chunk) per tile
  // operands of the very first step
  mbar_wait(afull0 + 8u * as, aph);
  mbar_wait(bfull0 + 8u * (brow * KW), bph);
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++) {
    const int acc = it & 1;
    mbar_wait(tempty0 + 8u * acc, ((it >> 1) & 1) ^ 1);
    const uint32_t d_tmem = tmem_base + acc * acc_cols;
    uint32_t first = 0;
    const bool more_tiles = t + (int)gridDim.x < total_tiles;
    for (int st = 0; st < steps; st++) {
      const uint32_t a_lo0 = a_lo_base + (uint32_t)as * (a_slot_bytes >> 4);
      const uint32_t b_lo0 = b_lo_base + (uint32_t)(brow * KW) * b_sub;
      const uint32_t bf = bfull0 + 8u * (brow * KW), be = bempty0 + 8u * (brow * KW);
      const bool last_step = st == steps - 1;
      // next step's ring positions (needed for the look-ahead waits)
      int as_n = as + 1, brow_n = brow + 1;
      uint32_t aph_n = aph, bph_n = bph;
      if (as_n == p.a_slots) { as_n = 0; aph_n ^= 1; }
      if (brow_n == ring_rows) { brow_n = 0; bph_n ^= 1; }
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < KW; s++) {
        mma_tap<NS>(d_tmem, a_lo0 + 8u * s, a_sub, b_lo0 + (uint32_t)s * b_sub, p.ntile, idesc, first);
        first = 1;
        tc_commit(be + 8u * s);
        if (s == KW - 1) {
          tc_commit(aempty0 + 8u * as);
          if (last_step) tc_commit(tfull0 + 8u * acc);
          if (!last_step || more_tiles) {                              // operands of the next step
            mbar_wait(afull0 + 8u * as_n, aph_n);
            mbar_wait(bfull0 + 8u * (brow_n * KW), bph_n);
          }
        } else {
          mbar_wait(bf + 8u * (s + 1), bph);
        }
        tc_fence_after();
      }
      as = as_n; aph = aph_n; brow = brow_n; bph = bph_n;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Row-box variant (tiles that are 128 consecutive pixels of ONE image row, i.e. W >= 128): instead of one A tile per
// tap, ONE box of 128+kw-1 pixel lines per filter ROW is loaded; the kw taps of that row are the same lines read
// from start offsets 0, 128, 256 ... bytes (UMMA descriptor start = box + s*128).  L2->SM traffic for A drops from
// kh*kw to ~kh tiles per 64-channel slice; the weight slices stream through their own ring.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_conv_rb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                      const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_slot_bytes = p.S * p.rb_bytes;               // rb_bytes is a multiple of 1024
  const uint32_t b_slot_bytes = p.ntile * 128;
  const uint32_t b_base = base + p.a_slots * a_slot_bytes;
  const uint32_t ctrl = b_base + p.b_slots * b_slot_bytes;
  auto afull = [&](int s) { return ctrl + 8u * s; };
  auto aempty = [&](int s) { return ctrl + 8u * (p.a_slots + s); };
  auto bfull = [&](int s) { return ctrl + 8u * (2 * p.a_slots + s); };
  auto bempty = [&](int s) { return ctrl + 8u * (2 * p.a_slots + p.b_slots + s); };
  auto tfull_bar = [&](int a) { return ctrl + 8u * (2 * p.a_slots + 2 * p.b_slots + a); };
  auto tempty_bar = [&](int a) { return ctrl + 8u * (2 * p.a_slots + 2 * p.b_slots + 2 + a); };
  const uint32_t tmem_slot = ctrl + 8u * (2 * p.a_slots + 2 * p.b_slots + 4);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(gen_base + (tmem_slot - base));
  float* bias_s = (float*)(gen_base + (ctrl - base) + 1024);
  {
    const int ncols = p.n_ntiles * p.ntile;
    const int creal = p.d2s ? p.cph : p.Cout;
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) {
      const int c = p.d2s ? i % p.cph : i;
      bias_s[i] = (p.bias && c < creal && (!p.d2s || i < 4 * p.cph)) ? p.bias[c] : 0.f;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_cols = p.S * p.ntile;
  const int tmem_cols = 2 * acc_cols <= 128 ? 128 : (2 * acc_cols <= 256 ? 256 : 512);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_slots; s++) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < p.b_slots; s++) {
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  const int total_tiles = p.n_super * p.n_ntiles;
  const int cchunks = p.Cin / KCH;
  const uint32_t box_bytes = (uint32_t)(TILE_M + p.kw - 1) * 128u;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int st = t / p.n_ntiles, nt = t - st * p.n_ntiles;
      const int nv = min(p.S, p.n_mtiles - st * p.S);
      for (int r = 0; r < p.kh; r++) {
        for (int cc = 0; cc < cchunks; cc++) {
          const int c = cc * KCH;
          mbar_wait(aempty(as), aph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(afull(as), nv * box_bytes);
            // NOTE: keep this loop rolled and free of local arrays: with the tile coordinates precomputed into
            // per-thread arrays nvcc 12.9 unrolled/peeled it and boxes i >= 1 of every first chunk never landed
            // (tools/tc_probe.py *_S4 cases)
#pragma unroll 1
            for (int i = 0; i < nv; i++) {
              const int mt = st * p.S + i;                         // bh == bn == 1: a tile is a row segment
              int tx_, ty_, on;
              mtile_coords(p, mt, tx_, ty_, on);
              const int ox = tx_ * p.bw - p.pad;
              const int oy = ty_ - p.pad + r;
              const uint32_t dst = base + as * a_slot_bytes + i * p.rb_bytes;
              if (c < p.C1)
                tma_load_4d(&tmA, dst, afull(as), c, ox, oy, on);
              else
                tma_load_4d(&tmA2, dst, afull(as), c - p.C1, ox, oy, on);
            }
          }
          __syncwarp();
          if (++as == p.a_slots) { as = 0; aph ^= 1; }
          for (int s = 0; s < p.kw; s++) {
            mbar_wait(bempty(bs), bph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(bfull(bs), b_slot_bytes);
              tma_load_3d(&tmB, b_base + bs * b_slot_bytes, bfull(bs), c, nt * p.ntile, r * p.kw + s);
            }
            __syncwarp();
            if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE elected lane runs the whole role =====================
    // The tensor pipe queues only a few MMAs, so every pause of the issuing thread (barrier polls, warp syncs,
    // descriptor arithmetic) longer than ~2 MMAs starves it (tools/mma_rate.cu).  Hence: a single lane, no per-tap
    // warp synchronisation, the wait for the NEXT weight slice placed behind the current slice's MMAs, descriptors
    // advanced by additions.
    const bool unrolled = (p.kw == 5 || p.kw == 3) && p.b_slots % p.kw == 0 && (p.S == 1 || p.S == 2 || p.S == 4);
    if (unrolled && elect_one()) {
      // (tail super tiles run all S sub-tiles; the surplus accumulators are never read)
#define RB_ROLE(KW_, NS_) rb_mma_role_unrolled<KW_, NS_>(p, base, b_base, ctrl, tmem_base, total_tiles, cchunks)
      if (p.kw == 5) {
        if (p.S == 1) RB_ROLE(5, 1); else if (p.S == 2) RB_ROLE(5, 2); else RB_ROLE(5, 4);
      } else {
        if (p.S == 1) RB_ROLE(3, 1); else if (p.S == 2) RB_ROLE(3, 2); else RB_ROLE(3, 4);
      }
#undef RB_ROLE
    } else if (!unrolled && elect_one()) {
      const uint32_t idesc = umma_idesc_16(p.ntile, p.bf16);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      bool have_b = false;            // bfull(bs) of the upcoming tap has already been waited for
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        const int st = t / p.n_ntiles;
        const int nv = min(p.S, p.n_mtiles - st * p.S);
        uint32_t first = 0;           // accumulate flag of the very first MMA of every sub-tile
        for (int r = 0; r < p.kh; r++) {
          for (int cc = 0; cc < cchunks; cc++) {
            mbar_wait(afull(as), aph);
            const uint32_t a_lo0 = umma_lo_of(base + as * a_slot_bytes);
            const uint32_t a_sub = (uint32_t)p.rb_bytes >> 4;
            for (int s = 0; s < p.kw; s++) {
              if (!have_b) mbar_wait(bfull(bs), bph);
              have_b = false;
              tc_fence_after();
              const uint32_t b_lo = umma_lo_of(b_base + bs * b_slot_bytes);
              const uint32_t a_lo = a_lo0 + (uint32_t)s * 8u;          // +128 bytes per tap (address field is >> 4)
              if (nv == 2) mma_tap<2>(d_tmem, a_lo, a_sub, b_lo, p.ntile, idesc, first);
              else if (nv == 4) mma_tap<4>(d_tmem, a_lo, a_sub, b_lo, p.ntile, idesc, first);
              else if (nv == 1) mma_tap<1>(d_tmem, a_lo, a_sub, b_lo, p.ntile, idesc, first);
              else mma_tap<3>(d_tmem, a_lo, a_sub, b_lo, p.ntile, idesc, first);
              first = 1;
              tc_commit(bempty(bs));
              const bool last_s = s == p.kw - 1;
              if (last_s) tc_commit(aempty(as));
              if (last_s && r == p.kh - 1 && cc == cchunks - 1) tc_commit(tfull_bar(acc));
              if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
              // poll for the next weight slice while the MMAs just issued execute (same A slot: no other wait needed)
              if (!last_s) {
                mbar_wait(bfull(bs), bph);
                have_b = true;
              }
            }
            if (++as == p.a_slots) { as = 0; aph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    epilogue_dispatch(p, tmem_base, tfull_bar(0), tempty_bar(0), bias_s, warp, lane, total_tiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// (A CTA-pair variant of the row-box kernel -- tcgen05 cta_group::2, M = 256 over a cluster of two CTAs, each supplying
// half of the weight rows -- was written and validated in round 1 and measured 3-8 % SLOWER than this kernel on every
// layer, N = 64 included (profiles/r1_tc_pair_experiment.txt); it was removed in round 2.)

// ---- split-K variant for the small layers ------------------------------------------------------------------------
// (hm_tc_conv_ws with a caller workspace.  Written from the round-1 launch list, where the 4x4..16x16 layers ran
// 4..128 CTAs for 20-90 us each because one CTA walks all K = taps*Cin/64 stages of its tile.  Measured on B200 in
// round 2: DCGAN step -0.1 ms, joint step -0.5 ms.)
// Work item = (M tile, N tile, K slice): the slice's partial accumulator is STORED to its own plane of an fp32
// workspace ws[slice][pixel][GEMM column]; tc_splitk_finish_kernel sums the planes in slice order -- a fixed order, so
// the result is deterministic (a first version reduced the slices with fp32 atomics: run-to-run differences of one
// fp16 ulp in a few activations, which the max-pools downstream amplify to percents of a gradient) -- and applies bias /
// activation / accumulate / depth-to-space exactly like epilogue_loop.  The workspace needs no initialisation.
struct TcSplitParams {
  TcParams p;        // S = 1, n_super = n_mtiles
  float* ws;         // [ksplit][B*Ho*Wo][ws_cols] fp32 scratch
  int ws_cols;       // n_ntiles * ntile
  int ksplit;        // K slices per tile (<= taps * Cin/64)
  long long plane;   // elements per slice plane = B*Ho*Wo*ws_cols
};

__device__ __forceinline__ void st_v4(float* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(a, b, c, d);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_conv_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                          const __grid_constant__ CUtensorMap tmB, const TcSplitParams q) {
  const TcParams& p = q.p;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = A_BYTES + p.ntile * 128;
  const uint32_t ctrl = base + p.stages * stage_bytes;        // 1024-aligned
  // control block: full[stages] | empty[stages] | tfull[2] | tempty[2] | tmem base address
  auto full_bar = [&](int s) { return ctrl + 8u * s; };
  auto empty_bar = [&](int s) { return ctrl + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return ctrl + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return ctrl + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = ctrl + 8u * (2 * p.stages + 4);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(gen_base + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_cols = p.ntile;                          // one M tile per work item, double-buffered accumulator
  const int tmem_cols = 2 * acc_cols <= 128 ? 128 : (2 * acc_cols <= 256 ? 256 : 512);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;

  const int ksteps_per_tap = p.Cin / KCH;
  const int ksteps = p.kh * p.kw * ksteps_per_tap;
  const int total_items = p.n_mtiles * p.n_ntiles * q.ksplit;     // item = (tile, K slice); tile = (M tile, N tile)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const int t = w / q.ksplit, sl = w - t * q.ksplit;
        const int mt = t / p.n_ntiles, nt = t - mt * p.n_ntiles;
        const int ks0 = (int)(((long long)sl * ksteps) / q.ksplit);
        const int ks1 = (int)(((long long)(sl + 1) * ksteps) / q.ksplit);
        const int bx = (mt % p.tiles_x) * p.bw * p.stride - p.pad;
        const int by = ((mt / p.tiles_x) % p.tiles_y) * p.bh * p.stride - p.pad;
        const int on = (mt / (p.tiles_x * p.tiles_y)) * p.bn;
#pragma unroll 1
        for (int ks = ks0; ks < ks1; ks++) {
          const int tap = ks / ksteps_per_tap, cc = ks - tap * ksteps_per_tap;
          const int r = tap / p.kw, s = tap - r * p.kw;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = base + stage * stage_bytes;
          mbar_expect_tx(full_bar(stage), A_BYTES + p.ntile * 128);
          const int c = cc * KCH;
          if (c < p.C1)
            tma_load_4d(&tmA, sa, full_bar(stage), c, bx + s, by + r, on);
          else
            tma_load_4d(&tmA2, sa, full_bar(stage), c - p.C1, bx + s, by + r, on);
          tma_load_3d(&tmB, sa + A_BYTES, full_bar(stage), c, nt * p.ntile, tap);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_16(p.ntile, p.bf16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x, it++) {
        const int t = w / q.ksplit, sl = w - t * q.ksplit;
        const int ks0 = (int)(((long long)sl * ksteps) / q.ksplit);
        const int ks1 = (int)(((long long)(sl + 1) * ksteps) / q.ksplit);
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        uint32_t first = 0;
        for (int ks = ks0; ks < ks1; ks++) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_lo = umma_lo_of(base + stage * stage_bytes);
          const uint32_t b_lo = a_lo + (A_BYTES >> 4);
          mma_tap<1>(d_tmem, a_lo, A_BYTES >> 4, b_lo, p.ntile, idesc, first);
          first = 1;
          tc_commit(empty_bar(stage));                        // frees the stage when these MMAs retire
          if (ks == ks1 - 1) tc_commit(tfull_bar(acc));       // partial accumulator complete
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5): partial sums -> this slice's plane of the workspace =============
    const int qd = warp & 3;                                  // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;
    const int px_per_img = p.bw * p.bh;
    const int in = row / px_per_img;
    const int rem = row - in * px_per_img;
    const int iy = rem / p.bw, ix = rem - iy * p.bw;
    int it = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x, it++) {
      const int t = w / q.ksplit, sl = w - t * q.ksplit;
      const int mt = t / p.n_ntiles, nt = t - mt * p.n_ntiles;
      const int acc = it & 1;
      mbar_wait(tfull_bar(acc), (it >> 1) & 1);
      tc_fence_after();
      const int tx = mt % p.tiles_x;
      const int ty = (mt / p.tiles_x) % p.tiles_y;
      const int tn = mt / (p.tiles_x * p.tiles_y);
      const int n = tn * p.bn + in, oy = ty * p.bh + iy, ox = tx * p.bw + ix;
      const bool valid = n < p.B && oy < p.Ho && ox < p.Wo;
      const size_t pix = (size_t)((size_t)n * p.Ho + oy) * p.Wo + ox;
      float* dst = q.ws + (size_t)sl * q.plane + pix * q.ws_cols + nt * p.ntile;
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + acc * acc_cols;
      for (int c0 = 0; c0 < p.ntile; c0 += 32) {
        uint32_t v[32];
        const bool full32 = p.ntile - c0 >= 32;
        if (full32) tmem_ld32(taddr + c0, v);
        else tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (!valid) continue;
        if (full32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) st_v4(dst + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 4) st_v4(dst + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// sum over slices of ws[slice][pixel][column] -> bias, (+ stored value), activation -> fp16 output (plain / two concat
// targets / depth-to-space), 8 columns per thread.
__global__ void __launch_bounds__(256) tc_splitk_finish_kernel(const TcSplitParams q) {
  const TcParams& p = q.p;
  const int groups = q.ws_cols >> 3;
  const long long npix = (long long)p.B * p.Ho * p.Wo;
  const long long total = npix * groups;
  const int creal = p.d2s ? 4 * p.cph : p.Cout;            // GEMM columns that exist (the N tile may be zero-padded)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / groups;
    const int col = (int)(i - pix * groups) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int sl = 0; sl < q.ksplit; sl++) {                  // fixed summation order: deterministic
      const float4* w4 = reinterpret_cast<const float4*>(q.ws + (size_t)sl * q.plane + pix * q.ws_cols + col);
      const float4 lo = w4[0], hi = w4[1];
      v[0] += lo.x; v[1] += lo.y; v[2] += lo.z; v[3] += lo.w;
      v[4] += hi.x; v[5] += hi.y; v[6] += hi.z; v[7] += hi.w;
    }
    const int ox = (int)(pix % p.Wo);
    const int oy = (int)((pix / p.Wo) % p.Ho);
    const int n = (int)(pix / ((long long)p.Wo * p.Ho));
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int c = col + j;
      if (c >= creal) continue;
      __half* dst;
      int cb;                                                // channel whose bias applies
      bool accum;
      if (p.d2s) {
        const int ph = c / p.cph, co = c - ph * p.cph;
        const size_t op = ((size_t)((size_t)n * 2 * p.Ho + 2 * oy + (ph >> 1)) * (2 * p.Wo) + 2 * ox + (ph & 1));
        dst = p.y + op * p.cph + co;
        cb = co;
        accum = (p.accumulate & 1) != 0;
      } else if (c < p.split) {
        dst = p.y ? p.y + (size_t)pix * p.split + c : nullptr;
        cb = c;
        accum = (p.accumulate & 1) != 0;
      } else {
        dst = p.y2 ? p.y2 + (size_t)pix * (p.Cout - p.split) + (c - p.split) : nullptr;
        cb = c;
        accum = (p.accumulate & 2) != 0;
      }
      if (!dst) continue;
      float a = v[j] + (p.bias ? p.bias[cb] : 0.f);
      if (accum) a += __half2float(*dst);
      switch (p.act) {
        case HM_ACT_LRELU: a = act_t<HM_ACT_LRELU>(a, p.slope); break;
        case HM_ACT_RELU: a = act_t<HM_ACT_RELU>(a, p.slope); break;
        case HM_ACT_SIGMOID: a = act_t<HM_ACT_SIGMOID>(a, p.slope); break;
        case HM_ACT_TANH: a = act_t<HM_ACT_TANH>(a, p.slope); break;
        default: break;
      }
      *dst = __float2half_rn(a);
    }
  }
}

// ---- host side ---------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static inline int pow2_floor(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

// NHWC fp16 activation tensor [B,H,W,C], box {64, bw, bh, bn}
static int encode_act(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int bw, int bh, int bn,
                      int stride = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // with element strides the box is given in traversed source elements: bw*stride of them yield bw loaded pixels
  cuuint32_t box[4] = {(cuuint32_t)KCH, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// weight pack [taps][Cout][Cin] fp16, box {64, ntile, 1}
static int encode_wgt(CUtensorMap* tm, const void* ptr, int taps, int Cout, int Cin, int ntile) {
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
  cuuint32_t box[3] = {(cuuint32_t)KCH, (cuuint32_t)ntile, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int pick_ntile(int Cout, int split) {
  if (Cout % 16 || split % 16) {
    // thin output: one zero-padded N tile (weight rows beyond Cout are TMA out-of-bounds = 0)
    if (split != Cout || Cout > 256) return 0;
    return (Cout + 15) / 16 * 16;
  }
  for (int n = 256; n >= 16; n -= 16)
    if (Cout % n == 0 && split % n == 0) return n;
  return 0;
}

}  // namespace hm

using namespace hm;

// nearest-2x upsampling feeding a 5x5 'same' stride-1 convolution: evaluated as four 3x3 convolutions on the
// low-resolution source (pack mode 8), never materialising the upsampled tensor
static bool is_up2conv(const HmConvDesc* d) {
  return d->up == HM_UP_NEAREST2 && d->kh == 5 && d->kw == 5 && d->pad == 2 && d->stride == 1 && !d->transposed &&
         d->C2 == 0 && d->Ho == 2 * d->H && d->Wo == 2 * d->W && d->split == d->Cout && !d->accumulate &&
         (d->Cout % 32 == 0 || d->Cout <= 4);
}

// input gradient of a 3x3 stride-2 pad-1 convolution (descriptor in hm_conv_gather's transposed form): evaluated as a
// 2x2-tap convolution of dy on ITS grid with N = (input-pixel phase, ci) and the depth-to-space epilogue; w_tc is pack
// mode 12 (taps the phase does not use are zero)
static bool is_dgrad_s2(const HmConvDesc* d) {
  return d->transposed == 1 && d->stride == 2 && d->kh == 3 && d->kw == 3 && d->pad == 1 && !d->up &&
         d->Ho == 2 * d->H && d->Wo == 2 * d->W;
}

// Deconv2DLayer with a 2x2 filter and stride 2 (reference architectures/p2p.py:23-24,272; no overlap): every input
// pixel q produces the 2x2 output block y[2q + (u,v)][co] = sum_ci x[q][ci] * W[ci][co][1-u][1-v], i.e. a 1x1
// convolution with N = (phase, co) columns (pack mode 17) followed by the depth-to-space scatter of the epilogue.
// Descriptor: transposed == 2, kh = kw = 2, stride = 2, H x W = INPUT grid, Ho = 2H, Wo = 2W.
static bool is_deconv_d2s(const HmConvDesc* d) {
  return d->transposed == 2 && d->kh == 2 && d->kw == 2 && d->stride == 2 && d->pad == 0 && !d->up &&
         d->Ho == 2 * d->H && d->Wo == 2 * d->W && d->oH == d->Ho && d->oW == d->Wo && d->os == 1 && !d->ou && !d->ov &&
         d->split == d->Cout && !d->accumulate;
}

// fp16 tensors, or bf16 hi/lo splits of fp32 tensors with float results (HM_BF16X3, hm_split_bf16x3)
static inline bool tc_dtype_ok(int dt) { return dt == HM_F16 || dt == HM_BF16X3; }

extern "C" int hm_tc_conv_supported(const HmConvDesc* d) {
  if (!d) return 0;
  if (tc_dtype_ok(d->dtype) && is_deconv_d2s(d))
    return d->C1 % KCH == 0 && d->C2 % KCH == 0 && d->C1 > 0 && (d->Cout % 32 == 0 || d->Cout <= 4) && 4 * d->Cout <= 2048;
  if (tc_dtype_ok(d->dtype) && is_up2conv(d))
    return d->C1 % KCH == 0 && d->C1 > 0 && d->os == 1 && !d->ou && !d->ov && d->oH == d->Ho && d->oW == d->Wo;
  if (tc_dtype_ok(d->dtype) && is_dgrad_s2(d))
    return d->C1 % KCH == 0 && d->C1 > 0 && d->C2 == 0 && d->os == 1 && !d->ou && !d->ov && d->oH == d->Ho &&
           d->oW == d->Wo && d->split == d->Cout && (d->Cout % 32 == 0 || d->Cout <= 4);
  if (!tc_dtype_ok(d->dtype) || d->transposed || d->up || (d->stride != 1 && d->stride != 2)) return 0;
  if (d->stride == 2) {
    if (d->Ho != (d->H + 2 * d->pad - d->kh) / 2 + 1 || d->Wo != (d->W + 2 * d->pad - d->kw) / 2 + 1) return 0;
    if (d->os != 1 || d->ou || d->ov || d->split <= 0 || d->split > d->Cout) return 0;
    if (d->C1 % KCH || d->C2 % KCH || d->C1 <= 0) return 0;
    if (d->Cout < 1 || pick_ntile(d->Cout, d->split) == 0) return 0;
    return d->oH == d->Ho && d->oW == d->Wo;
  }
  if (d->os != 1 || d->ou || d->ov || d->split <= 0 || d->split > d->Cout) return 0;
  // channels: multiples of 64 per source -- or ONE source whose channel count is any multiple of 8 (16-byte rows: a
  // DenseLayer's input, dcgan.py:16, latent_dim = 1000): the last 64-channel K slice is then partly out of bounds in
  // both tensor maps and zero-filled by the TMA unit
  if (d->C1 <= 0 || d->C2 % KCH || (d->C1 % KCH && (d->C2 != 0 || d->C1 % 8 || d->dtype != HM_F16))) return 0;
  if (d->Cout < 1 || pick_ntile(d->Cout, d->split) == 0) return 0;
  if (d->Cout > 8192 || (d->Cout > 2048 && d->Wo >= TILE_M)) return 0;
  if (d->Ho != d->H + 2 * d->pad - d->kh + 1 || d->Wo != d->W + 2 * d->pad - d->kw + 1) return 0;
  if (d->oH != d->Ho || d->oW != d->Wo) return 0;
  return 1;
}

static bool splitk_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HMGAN_TC_SPLITK");      // on unless HMGAN_TC_SPLITK=0
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// K slices per tile for the split-K variant (1 = do not split): only layers whose tiles fill less than half of the SMs
// and that walk at least 8 K stages per tile; at least 4 stages per slice.
static int splitk_factor(long long tiles, int ksteps) {
  const int sms = num_sms();
  if (tiles < 1 || tiles * 2 > sms || ksteps < 8) return 1;
  long long k = sms / tiles;
  if (k > ksteps / 4) k = ksteps / 4;
  return k < 2 ? 1 : (int)k;
}

// Bytes of fp32 scratch hm_tc_conv_ws may use for this problem: 0 when the split-K variant would not be chosen
// (switched off, unsupported shape, or enough tiles to fill the machine).  An upper bound that depends on d alone.
extern "C" long long hm_tc_conv_ws_bytes(const HmConvDesc* d) {
  if (!d || !splitk_enabled() || d->dtype != HM_F16 || !hm_tc_conv_supported(d)) return 0;
  const bool dc2 = is_deconv_d2s(d), dg2 = is_dgrad_s2(d);
  const bool phase = is_up2conv(d) || dg2 || dc2;
  const long long gh = phase ? d->H : d->Ho, gw = phase ? d->W : d->Wo;        // tile grid
  const int kh = dc2 ? 1 : (dg2 ? 2 : (phase ? 3 : d->kh)), kw = kh == d->kh && !phase ? d->kw : kh;
  if (gw >= TILE_M && kw > 1) return 0;                                         // row-box kernels take those layers
  const long long cols = ((phase ? 4LL * d->Cout : d->Cout) + 15) / 16 * 16;
  const long long npix = (long long)d->B * gh * gw;
  const int ntile = pick_ntile((int)(phase ? 4 * d->Cout : d->Cout), phase ? 4 * d->Cout : d->split);
  if (ntile <= 0) return 0;
  const long long tiles = (npix + TILE_M - 1) / TILE_M * ((cols + ntile - 1) / ntile);
  const int ksplit = splitk_factor(tiles, kh * kw * ((d->C1 + d->C2 + KCH - 1) / KCH));
  if (ksplit <= 1) return 0;
  return npix * ((cols + ntile - 1) / ntile * ntile) * 4 * ksplit;
}

static int tc_conv_impl(const HmConvDesc* d, const void* x1, const void* x2, const void* w_tc, const float* bias,
                        void* y, void* y2, void* ws, size_t ws_bytes, void* stream, uint8_t* pool_idx = nullptr);

// y[B,Ho,Wo,Cout] = act( corr(x1|x2, w_tc) + bias );  w_tc is the pack [kh*kw][Cout][C1+C2] (fp16, K-major).
extern "C" int hm_tc_conv(const HmConvDesc* d, const void* x1, const void* x2, const void* w_tc, const float* bias,
                          void* y, void* y2, void* stream) {
  return tc_conv_impl(d, x1, x2, w_tc, bias, y, y2, nullptr, 0, stream);
}

// The same with a caller-provided scratch workspace of ws_bytes >= hm_tc_conv_ws_bytes(d) bytes (no initialisation
// needed; one buffer serves every call on a stream): lets small layers split K over otherwise idle SMs.
extern "C" int hm_tc_conv_ws(const HmConvDesc* d, const void* x1, const void* x2, const void* w_tc, const float* bias,
                             void* y, void* y2, void* ws, long long ws_bytes, void* stream) {
  HM_CHECK_ARG(!ws || ((((uintptr_t)ws) & 15) == 0 && ws_bytes >= 0), "hm_tc_conv_ws: workspace must be 16-byte aligned");
  return tc_conv_impl(d, x1, x2, w_tc, bias, y, y2, ws, (size_t)ws_bytes, stream);
}

// largest N tile (multiple of 32, <= 128) of a pooled layer: two accumulator sets x S = 2 sub-tiles x ntile <= 512 columns
static int pool_ntile(int Cout) {
  for (int n = 128; n >= 32; n -= 32)
    if (Cout % n == 0) return n;
  return 0;
}

// conv + bias + activation + MaxPool2DLayer(2) in one pass (reference architectures/dcgan.py:42-47 for the layers whose
// rows are at least 128 pixels wide): plain stride-1 convolution, fp16, one output tensor, even Ho and Wo, Wo >= 128,
// Cout a multiple of 32, a monotonic activation.
extern "C" int hm_tc_conv_pool_supported(const HmConvDesc* d) {
  if (!d || d->dtype != HM_F16 || !hm_tc_conv_supported(d)) return 0;
  if (is_up2conv(d) || is_dgrad_s2(d) || is_deconv_d2s(d) || d->transposed || d->up || d->stride != 1) return 0;
  if (d->split != d->Cout || d->accumulate || d->Cout % 32 || pool_ntile(d->Cout) == 0) return 0;
  if (d->Ho % 2 || d->Wo % 2 || d->Wo < TILE_M || d->kw < 2 || d->kw > 9) return 0;
  if (d->act != HM_ACT_LINEAR && d->act != HM_ACT_LRELU && d->act != HM_ACT_RELU) return 0;
  if (d->act == HM_ACT_LRELU && d->slope < 0.f) return 0;
  return 1;
}

extern "C" int hm_tc_conv_pool(const HmConvDesc* d, const void* x1, const void* x2, const void* w_tc, const float* bias,
                               void* y_pooled, uint8_t* idx, void* stream) {
  HM_CHECK_ARG(d && x1 && w_tc && y_pooled && idx, "hm_tc_conv_pool: null argument");
  if (!hm_tc_conv_pool_supported(d)) {
    set_error("hm_tc_conv_pool: shape not supported (plain stride-1 fp16 convolution, even Ho/Wo, Wo >= 128, Cout %% 32 == 0)");
    return HM_ERR_UNSUPPORTED;
  }
  if (((uintptr_t)idx) & 15) {
    set_error("hm_tc_conv_pool: idx must be 16-byte aligned");
    return HM_ERR_ALIGN;
  }
  return tc_conv_impl(d, x1, x2, w_tc, bias, y_pooled, nullptr, nullptr, 0, stream, idx);
}

static int tc_conv_impl(const HmConvDesc* d, const void* x1, const void* x2, const void* w_tc, const float* bias,
                        void* y, void* y2, void* ws, size_t ws_bytes, void* stream, uint8_t* pool_idx) {
  HM_CHECK_ARG(d && x1 && w_tc && (y || y2), "hm_tc_conv: null argument");
  if (!hm_tc_conv_supported(d)) {
    set_error("hm_tc_conv: shape not supported by the tcgen05 path (need fp16, stride 1, C%%64==0, Cout%%16==0)");
    return HM_ERR_UNSUPPORTED;
  }
  HM_CHECK_ARG(d->C2 == 0 || x2, "hm_tc_conv: C2>0 but x2 is null");
  if (!encode_fn()) {
    set_error("hm_tc_conv: cuTensorMapEncodeTiled is not available from this driver");
    return HM_ERR_CUDA;
  }
  if ((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)w_tc) & 15) ||
      ((d->Cout % 16 == 0) && (((uintptr_t)y | (uintptr_t)y2) & 15)) ||
      ((is_up2conv(d) || is_dgrad_s2(d) || is_deconv_d2s(d)) && !y)) {
    set_error("hm_tc_conv: pointers must be 16-byte aligned");
    return HM_ERR_ALIGN;
  }
  TcParams p;
  const bool dg2 = is_dgrad_s2(d);
  const bool dc2 = is_deconv_d2s(d);
  const bool up2 = is_up2conv(d) || dg2 || dc2;                           // all use the phase / depth-to-space form
  p.d2s = up2 ? 1 : 0;
  p.cph = d->Cout;
  p.stride = (!up2 && d->stride == 2) ? 2 : 1;
  p.B = d->B; p.Ho = up2 ? d->H : d->Ho; p.Wo = up2 ? d->W : d->Wo;       // tile grid (low-res when phase-decomposed)
  const int cin_real = d->C1 + d->C2;
  p.Cin = (cin_real + KCH - 1) / KCH * KCH;           // K slices of 64 channels; a ragged tail is TMA zero fill
  p.C1 = d->C2 ? d->C1 : p.Cin; p.Cout = up2 ? 4 * d->Cout : d->Cout;
  p.kh = dc2 ? 1 : (dg2 ? 2 : (up2 ? 3 : d->kh)); p.kw = dc2 ? 1 : (dg2 ? 2 : (up2 ? 3 : d->kw));
  p.pad = (dg2 || dc2) ? 0 : (up2 ? 1 : d->pad);
  p.bw = pow2_floor(p.Wo < TILE_M ? p.Wo : TILE_M);
  p.bh = pow2_floor(p.Ho < TILE_M / p.bw ? p.Ho : TILE_M / p.bw);
  p.bn = TILE_M / (p.bw * p.bh);
  p.tiles_x = (p.Wo + p.bw - 1) / p.bw;
  p.tiles_y = (p.Ho + p.bh - 1) / p.bh;
  p.tiles_n = (d->B + p.bn - 1) / p.bn;
  p.n_mtiles = p.tiles_x * p.tiles_y * p.tiles_n;
  p.pool = pool_idx ? 1 : 0;
  p.idx = pool_idx;
  p.ntile = p.pool ? pool_ntile(p.Cout) : pick_ntile(p.Cout, up2 ? p.Cout : d->split);
  p.n_ntiles = (p.Cout + p.ntile - 1) / p.ntile;
  // the kernels keep the bias of every GEMM column in shared memory: 2048 columns for the row-box variant, up to
  // 8192 (a DenseLayer's width) for the plain kernel, which only small spatial extents reach
  const int ncols_pad = p.n_ntiles * p.ntile;
  const bool wide = ncols_pad > 2048;
  if (ncols_pad > 8192 || (wide && p.Wo >= TILE_M)) {
    set_error("hm_tc_conv: %d GEMM columns are not supported (2048; 8192 for rows shorter than 128 pixels)", ncols_pad);
    return HM_ERR_UNSUPPORTED;
  }
  const int bias_bytes = wide ? 4 * ncols_pad : 8192;
  p.vec_store = (p.Cout % 16 == 0 && d->split % 16 == 0) ? 1 : 0;
  int S = 256 / p.ntile;                                   // 2 (double buffer) * S * ntile <= 512 TMEM columns
  if (S < 1) S = 1;
  if (S > 4) S = 4;
  {
    static int smax = -1;                                  // diagnostic cap on the super-tile size
    if (smax < 0) {
      const char* e = getenv("HMGAN_TC_SMAX");
      smax = e ? atoi(e) : 4;
      if (smax < 1) smax = 1;
    }
    if (S > smax) S = smax;
  }
  while (S > 1 && ((p.n_mtiles + S - 1) / S) * p.n_ntiles < num_sms()) S >>= 1;   // keep every SM busy on small layers
  if (p.pool) S = 2;                                       // the two image rows of a pooling window
  p.S = S;
  p.n_super = (p.n_mtiles + S - 1) / S;
  const int stage_bytes = S * A_BYTES + p.ntile * 128;
  int stages = (227 * 1024 - 2048 - bias_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  p.stages = stages;
  p.act = d->act; p.slope = d->slope; p.bias = bias; p.y = (__half*)y; p.y2 = (__half*)y2;
  p.split = up2 ? p.Cout : d->split; p.accumulate = d->accumulate;
  p.bf16 = p.out32 = d->dtype == HM_BF16X3 ? 1 : 0;

  CUtensorMap tmA, tmA2, tmB;
  int rc = encode_act(&tmA, x1, d->B, d->H, d->W, d->C1, p.bw, p.bh, p.bn, p.stride);
  if (!rc) rc = d->C2 ? encode_act(&tmA2, x2, d->B, d->H, d->W, d->C2, p.bw, p.bh, p.bn, p.stride) : 0;
  if (!d->C2) tmA2 = tmA;
  if (!rc) rc = encode_wgt(&tmB, w_tc, p.kh * p.kw, p.Cout, cin_real, p.ntile);
  if (rc) {
    set_error("hm_tc_conv: cuTensorMapEncodeTiled failed (CUresult %d)", rc);
    return HM_ERR_CUDA;
  }
  // Row-box variant: tiles are row segments (bw == 128) and the filter is wider than one tap.
  static int rb_enabled = -1;
  if (rb_enabled < 0) {
    const char* e = getenv("HMGAN_TC_ROWBOX");
    rb_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if ((rb_enabled || p.pool) && p.stride == 1 && p.bw == TILE_M && p.bh == 1 && p.bn == 1 && p.kw > 1 && p.kw <= 9) {
    p.rb_bytes = (((TILE_M + p.kw - 1) * 128) + 1023) / 1024 * 1024;
    p.a_slots = (S * p.rb_bytes > 40 * 1024) ? 2 : 3;
    {
      const char* e = getenv("HMGAN_RB_ASLOTS");            // diagnostic override
      if (e && atoi(e) >= 1) p.a_slots = atoi(e);
    }
    int b_slots = (227 * 1024 - 10240 - p.a_slots * S * p.rb_bytes) / (p.ntile * 128);
    if (p.kw == 5 || p.kw == 3) {                           // ring of whole filter rows: enables the unrolled MMA role
      if (b_slots >= 2 * p.kw) b_slots = 2 * p.kw;
      else if (b_slots >= p.kw) b_slots = p.kw;
    }
    {
      static int bcap = -1;                                 // tuning knob: depth of the weight ring
      if (bcap < 0) {
        const char* e = getenv("HMGAN_RB_BCAP");
        bcap = (e && atoi(e) >= 2) ? atoi(e) : 10;
      }
      if (b_slots > bcap) b_slots = bcap;
    }
    {
      const char* e = getenv("HMGAN_RB_BSLOTS");            // diagnostic override
      if (e && atoi(e) >= 2 && atoi(e) < b_slots) b_slots = atoi(e);
    }
    if (b_slots >= 2) {
      p.b_slots = b_slots;
      CUtensorMap rA, rA2;
      rc = encode_act(&rA, x1, d->B, d->H, d->W, d->C1, TILE_M + p.kw - 1, 1, 1);
      if (!rc) rc = d->C2 ? encode_act(&rA2, x2, d->B, d->H, d->W, d->C2, TILE_M + p.kw - 1, 1, 1) : 0;
      if (!d->C2) rA2 = rA;
      if (rc) {
        set_error("hm_tc_conv: cuTensorMapEncodeTiled failed for the row box (CUresult %d)", rc);
        return HM_ERR_CUDA;
      }
      const size_t smem_rb = (size_t)p.a_slots * S * p.rb_bytes + (size_t)p.b_slots * p.ntile * 128 + 1024 + 1024 + 8192;
      static bool rb_attr = false;
      if (!rb_attr) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_rb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
          set_error("hm_tc_conv: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
          return HM_ERR_CUDA;
        }
        rb_attr = true;
      }
      int grid_rb = p.n_super * p.n_ntiles;
      if (grid_rb > num_sms()) grid_rb = num_sms();
      tc_conv_rb_kernel<<<grid_rb, TC_THREADS, smem_rb, (cudaStream_t)stream>>>(rA, rA2, tmB, p);
      HM_CHECK_LAUNCH("hm_tc_conv(row box)");
      return HM_OK;
    }
  }
  if (p.pool) {
    set_error("hm_tc_conv_pool: the row-box kernel did not take this shape");
    return HM_ERR_UNSUPPORTED;
  }
  if (ws && splitk_enabled() && !p.bf16) {
    const int ksteps = p.kh * p.kw * (p.Cin / KCH);
    const int ksplit = splitk_factor((long long)p.n_mtiles * p.n_ntiles, ksteps);
    const size_t plane = (size_t)d->B * p.Ho * p.Wo * (size_t)(p.n_ntiles * p.ntile);
    const size_t need = plane * 4 * (size_t)(ksplit > 1 ? ksplit : 1);
    if (ksplit > 1 && ws_bytes >= need) {
      TcSplitParams q;
      q.p = p;
      q.p.S = 1;
      q.p.n_super = p.n_mtiles;
      const int sk_stage = A_BYTES + p.ntile * 128;
      int sk_stages = (227 * 1024 - 10240) / sk_stage;
      if (sk_stages > 8) sk_stages = 8;
      q.p.stages = sk_stages;
      q.ws = (float*)ws;
      q.ws_cols = p.n_ntiles * p.ntile;
      q.ksplit = ksplit;
      q.plane = (long long)plane;
      static bool sk_attr = false;
      if (!sk_attr) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
          set_error("hm_tc_conv: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
          return HM_ERR_CUDA;
        }
        sk_attr = true;
      }
      const size_t sk_smem = (size_t)sk_stages * sk_stage + 1024 /*alignment slack*/ + 1024 /*control block*/;
      long long items = (long long)p.n_mtiles * p.n_ntiles * ksplit;
      int sk_grid = items > num_sms() ? num_sms() : (int)items;
      tc_conv_splitk_kernel<<<sk_grid, TC_THREADS, sk_smem < 120 * 1024 ? 120 * 1024 : sk_smem, (cudaStream_t)stream>>>(
          tmA, tmA2, tmB, q);
      HM_CHECK_LAUNCH("hm_tc_conv(split K)");
      const long long groups = (long long)d->B * p.Ho * p.Wo * (q.ws_cols / 8);
      long long fb = (groups + 255) / 256;
      if (fb > (long long)num_sms() * 8) fb = (long long)num_sms() * 8;
      tc_splitk_finish_kernel<<<(unsigned)fb, 256, 0, (cudaStream_t)stream>>>(q);
      HM_CHECK_LAUNCH("hm_tc_conv(split K finish)");
      return HM_OK;
    }
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*alignment slack*/ + 1024 + bias_bytes /*control block + bias*/;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("hm_tc_conv: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
      return HM_ERR_CUDA;
    }
    attr_set = true;
  }
  int grid = p.n_super * p.n_ntiles;
  if (grid > num_sms()) grid = num_sms();
  tc_conv_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmA2, tmB, p);
  HM_CHECK_LAUNCH("hm_tc_conv");
  return HM_OK;
}
