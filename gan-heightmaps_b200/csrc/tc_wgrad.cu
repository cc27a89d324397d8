// tcgen05 weight-gradient kernel for sm_100a (fast mode).
//
// Stands in for theano CorrMM_gradWeights behind lasagne Conv2DLayer (reference architectures/dcgan.py:22,42;
// architectures/p2p.py:20-21,208-209):
//     dWp[k = (tap, ci)][co] += sum over output pixels  x(pixel shifted by tap)[ci] * dy[pixel][co]
// GEMM view: D[M = 128 rows of the flattened (tap, ci) axis][N = Cout] += A[M][K] * B[N][K] with the REDUCTION
// dimension K = pixels.  In NHWC both operands are stored pixel-major ([pixel][channel]), i.e. they are
// "MN-major" for tcgen05: the TMA boxes {64 ch, bw, bh, bn} land as [128 pixels][64 ch] 128B-swizzled tiles
// (the same loads the forward kernel uses), and the shared-memory descriptors read them transposed
// (a_major = b_major = MN; LBO = distance between 64-channel blocks, SBO = distance between 8-pixel groups).
//   * M tile = two 64-row blocks of the (tap, ci) axis: two taps for a 64-channel input, two channel blocks of
//     one tap otherwise; each block is the TMA tile of x at that tap's shift (zero padding = TMA out-of-bounds fill).
//   * A CTA owns up to 512/Cout M tiles (all their fp32 accumulators live in TMEM at once) and a contiguous range of
//     pixel tiles (split-K across CTAs); per pixel tile it loads the dy tile once and streams the x blocks.
//   * Epilogue: tcgen05.ld -> fp32 atomic adds into the packed gradient [kh*kw*Cin][Cout] (the layout of
//     hm_conv_wgrad, so hm_unpack_conv_wgrad applies unchanged).
#include <cuda.h>
#include <stdlib.h>

#include "hm_common.cuh"

namespace hm {

constexpr int WG_THREADS = 192;
constexpr int BLK_BYTES = 128 * 64 * 2;     // one [128 pixel][64 channel] fp16 tile = 16 KB

struct WgParams {
  int B, Ho, Wo;
  int Cin, C1, Cout;
  int kh, kw, pad, stride;
  int bw, bh, bn, tiles_x, tiles_y, tiles_n, n_ptiles;
  int units;            // 64-row blocks of the (tap, ci) axis = kh*kw*Cin/64
  int n_mtiles;         // ceil(units/2)
  int acc;              // M tiles per CTA (accumulators resident in TMEM)
  int n_mgroups;        // ceil(n_mtiles/acc)
  int zsplit;           // CTAs along the pixel axis
  int a_slots, b_slots;
  float* dw;            // [units*64][pitch] fp32; this launch fills columns [n_off, n_off + Cout)
  int n_off, pitch;     // N tiling when the layer has more than 256 output channels
  int bf16;             // operands are bfloat16 (HM_BF16X3: batch-stacked hi/lo splits of fp32 tensors) instead of fp16
  int cpb;              // > 0: phase form of (nearest-2x -> 5x5): the N columns are (phase, co) with cpb = Cout_real/64
                        //      channel blocks per phase; dy is the HIGH-res gradient, read with TMA element strides of 2
};

__device__ __forceinline__ uint32_t wg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wg_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void wg_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wg_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 26)) __trap();     // a protocol bug fails the launch instead of hanging the GPU
  }
}
__device__ __forceinline__ void wg_tma_4d(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1, int c2,
                                          int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// one lane of the converged warp (see elect_one() in tc_conv.cu: avoids the per-instruction waterfall loops that
// `if (lane == 0)` regions compile to for tcgen05 / TMA instructions)
__device__ __forceinline__ bool wg_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// 16-byte vector reduction into global memory (sm_90+): one L2 transaction for four consecutive floats.  The epilogue
// of these kernels is nothing but fp32 reductions (64K per CTA); scalar atomicAdd made it ~20 % of the kernel.
__device__ __forceinline__ void wg_red4(float* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}
__device__ __forceinline__ void wg_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wg_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void wg_ld32(uint32_t taddr, uint32_t* v) {
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

// MN-major, 128-byte-swizzled operand: 64 channels (128 B) contiguous, 64-channel blocks LBO apart,
// pixel rows 128 B apart inside a swizzle atom, 8-pixel groups SBO = 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, A and B MN-major (bits 15,16), D = F32, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn(int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
    tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
                    const __grid_constant__ CUtensorMap tmDY, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (wg_smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t nblk = p.Cout / 64;                       // dy blocks per pixel tile
  const uint32_t b_bytes = nblk * BLK_BYTES;
  const uint32_t a_bytes = 2 * BLK_BYTES;
  const uint32_t a_base = base + p.b_slots * b_bytes;
  const uint32_t ctrl = a_base + p.a_slots * a_bytes;
  auto afull = [&](int s) { return ctrl + 8u * s; };
  auto aempty = [&](int s) { return ctrl + 8u * (p.a_slots + s); };
  auto bfull = [&](int s) { return ctrl + 8u * (2 * p.a_slots + s); };
  auto bempty = [&](int s) { return ctrl + 8u * (2 * p.a_slots + p.b_slots + s); };
  const uint32_t tfull = ctrl + 8u * (2 * p.a_slots + 2 * p.b_slots);
  const uint32_t tmem_slot = tfull + 8u;
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(smem_raw + (tmem_slot - wg_smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_slots; s++) {
      wg_mbar_init(afull(s), 1);
      wg_mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < p.b_slots; s++) {
      wg_mbar_init(bfull(s), 1);
      wg_mbar_init(bempty(s), 1);
    }
    wg_mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDY) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_p;

  // work item
  const int mg = blockIdx.x % p.n_mgroups;
  const int z = blockIdx.x / p.n_mgroups;
  const int mt0 = (int)(((long long)mg * p.n_mtiles) / p.n_mgroups);          // balanced partition of the M tiles
  const int mt1 = (int)(((long long)(mg + 1) * p.n_mtiles) / p.n_mgroups);
  const int pt0 = (int)(((long long)z * p.n_ptiles) / p.zsplit);
  const int pt1 = (int)(((long long)(z + 1) * p.n_ptiles) / p.zsplit);
  const int cblocks = p.Cin / 64;

  if (warp == 0) {
    if (wg_elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int pt = pt0; pt < pt1; pt++) {
        const int tx = pt % p.tiles_x;
        const int ty = (pt / p.tiles_x) % p.tiles_y;
        const int tn = pt / (p.tiles_x * p.tiles_y);
        const int ox0 = tx * p.bw, oy0 = ty * p.bh, n0 = tn * p.bn;
        wg_wait(bempty(bs), bph ^ 1);
        wg_expect_tx(bfull(bs), b_bytes);
        for (uint32_t j = 0; j < nblk; j++) {
          const int jb = p.n_off / 64 + (int)j;                // 64-column block of the GEMM's N axis
          if (p.cpb > 0) {                                     // (phase, co): phase (py,px) = pixels (2y+py, 2x+px) of dy
            const int ph = jb / p.cpb, cb = jb - ph * p.cpb;
            wg_tma_4d(&tmDY, base + bs * b_bytes + j * BLK_BYTES, bfull(bs), cb * 64, 2 * ox0 + (ph & 1), 2 * oy0 + (ph >> 1), n0);
          } else {
            wg_tma_4d(&tmDY, base + bs * b_bytes + j * BLK_BYTES, bfull(bs), jb * 64, ox0, oy0, n0);
          }
        }
        if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
        for (int mt = mt0; mt < mt1; mt++) {
          wg_wait(aempty(as), aph ^ 1);
          wg_expect_tx(afull(as), a_bytes);
          for (int h = 0; h < 2; h++) {
            int u = 2 * mt + h;
            if (u >= p.units) u = p.units - 1;               // odd tail: duplicate (rows ignored by the epilogue)
            const int tap = u / cblocks, cb = u - tap * cblocks;
            const int r = tap / p.kw, s = tap - r * p.kw;
            const int c = cb * 64;
            const uint32_t dst = a_base + as * a_bytes + h * BLK_BYTES;
            if (c < p.C1)
              wg_tma_4d(&tmX, dst, afull(as), c, ox0 * p.stride - p.pad + s, oy0 * p.stride - p.pad + r, n0);
            else
              wg_tma_4d(&tmX2, dst, afull(as), c - p.C1, ox0 * p.stride - p.pad + s, oy0 * p.stride - p.pad + r, n0);
          }
          if (++as == p.a_slots) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (wg_elect_one()) {
      // single issuing lane, 32-bit descriptor arithmetic, the NEXT step's barriers polled right behind the MMAs
      const uint32_t idesc = umma_idesc_f16_mn(p.Cout) | (p.bf16 ? ((1u << 7) | (1u << 10)) : 0u);
      constexpr uint32_t HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);    // SBO 1024 B, version 1, SWIZZLE_128B
      constexpr uint32_t LBO = (uint32_t)(BLK_BYTES >> 4) << 16;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      if (pt1 > pt0 && mt1 > mt0) {
        wg_wait(bfull(bs), bph);
        wg_wait(afull(as), aph);
      }
      for (int pt = pt0; pt < pt1; pt++) {
        const uint32_t b_lo = (((base + bs * b_bytes) & 0x3FFFF) >> 4) | LBO;
        const uint32_t accum0 = pt > pt0 ? 1u : 0u;
        for (int mt = mt0; mt < mt1; mt++) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_lo = (((a_base + as * a_bytes) & 0x3FFFF) >> 4) | LBO;
          const uint32_t d_tmem = tmem_base + (mt - mt0) * p.Cout;
#pragma unroll
          for (int kk = 0; kk < 8; kk++)                     // 16 pixels (two 8-row swizzle atoms) per MMA
            wg_mma(d_tmem, ((uint64_t)HI << 32) | (uint64_t)(a_lo + kk * 128), ((uint64_t)HI << 32) | (uint64_t)(b_lo + kk * 128),
                   idesc, accum0 | (uint32_t)kk);
          wg_commit(aempty(as));
          if (++as == p.a_slots) { as = 0; aph ^= 1; }
          const bool last_mt = mt == mt1 - 1;
          if (last_mt) {
            wg_commit(bempty(bs));
            if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
            if (pt + 1 < pt1) {
              wg_wait(bfull(bs), bph);
              wg_wait(afull(as), aph);
            }
          } else {
            wg_wait(afull(as), aph);
          }
        }
      }
      wg_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (pt1 > pt0) {
      wg_wait(tfull, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int mt = mt0; mt < mt1; mt++) {
        const int k = mt * 128 + row;                        // row of the packed gradient
        const bool valid = k < p.units * 64;
        float* dst = p.dw + (size_t)k * p.pitch + p.n_off;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (mt - mt0) * p.Cout;
        for (int c0 = 0; c0 < p.Cout; c0 += 32) {
          uint32_t v[32];
          wg_ld32(taddr + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) wg_red4(dst + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Row-box variant (pixel tiles that are 128 consecutive pixels of one image row): the x operand of the taps
// (r, 0..kw-1) of one filter row and one 64-channel block is ONE TMA box of 128+kw-1 pixel lines; a tap is the same
// box read from line offset s (the 128B swizzle is address-based, see tc_conv.cu).  A CTA loads, per pixel tile, only
// the distinct (row, channel-block) boxes its units need instead of one 16 KB tile per unit.
// ---------------------------------------------------------------------------------------------------------------
struct WgRbParams {
  WgParams w;
  int rb_bytes;        // bytes reserved per box (multiple of 1024)
  int max_boxes;       // boxes per A slot
};

__global__ void __launch_bounds__(WG_THREADS, 1)
    tc_wgrad_rb_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
                       const __grid_constant__ CUtensorMap tmDY, const WgRbParams q) {
  const WgParams& p = q.w;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (wg_smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t nblk = p.Cout / 64;
  const uint32_t b_bytes = nblk * BLK_BYTES;
  const uint32_t a_bytes = (uint32_t)q.max_boxes * q.rb_bytes;
  const uint32_t a_base = base + p.b_slots * b_bytes;
  const uint32_t ctrl = a_base + p.a_slots * a_bytes;
  auto afull = [&](int s) { return ctrl + 8u * s; };
  auto aempty = [&](int s) { return ctrl + 8u * (p.a_slots + s); };
  auto bfull = [&](int s) { return ctrl + 8u * (2 * p.a_slots + s); };
  auto bempty = [&](int s) { return ctrl + 8u * (2 * p.a_slots + p.b_slots + s); };
  const uint32_t tfull = ctrl + 8u * (2 * p.a_slots + 2 * p.b_slots);
  const uint32_t tmem_slot = tfull + 8u;
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(smem_raw + (tmem_slot - wg_smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_slots; s++) {
      wg_mbar_init(afull(s), 1);
      wg_mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < p.b_slots; s++) {
      wg_mbar_init(bfull(s), 1);
      wg_mbar_init(bempty(s), 1);
    }
    wg_mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDY) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_p;

  const int mg = blockIdx.x % p.n_mgroups;
  const int z = blockIdx.x / p.n_mgroups;
  const int mt0 = (int)(((long long)mg * p.n_mtiles) / p.n_mgroups);
  const int mt1 = (int)(((long long)(mg + 1) * p.n_mtiles) / p.n_mgroups);
  const int pt0 = (int)(((long long)z * p.n_ptiles) / p.zsplit);
  const int pt1 = (int)(((long long)(z + 1) * p.n_ptiles) / p.zsplit);
  const int cblocks = p.Cin / 64;
  // units [u_lo, u_hi) of this CTA; box id of a unit = (filter row, channel block); boxes are numbered in order of
  // first use, which is ascending because units ascend in (tap, cb) order
  const int u_lo = 2 * mt0;
  const int u_hi = min(2 * mt1, p.units);
  auto box_of = [&](int u) {                       // index into this CTA's box list
    const int tap = u / cblocks, cb = u - tap * cblocks;
    const int r = tap / p.kw;
    const int tap0 = u_lo / cblocks;
    const int r0 = tap0 / p.kw;
    return (r - r0) * cblocks + cb;                // rows r0.. each with cblocks boxes (some unused at the ends)
  };
  const int r_first = (u_lo / cblocks) / p.kw;
  const int r_last = ((u_hi - 1) / cblocks) / p.kw;
  const int n_boxes = (r_last - r_first + 1) * cblocks;
  const uint32_t box_bytes = (uint32_t)(128 + p.kw - 1) * 128u;

  if (warp == 0) {
    if (wg_elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int pt = pt0; pt < pt1; pt++) {
        const int ox0 = (pt % p.tiles_x) * 128;
        const int oy0 = (pt / p.tiles_x) % p.tiles_y;
        const int n0 = pt / (p.tiles_x * p.tiles_y);
        wg_wait(bempty(bs), bph ^ 1);
        wg_expect_tx(bfull(bs), b_bytes);
        for (uint32_t j = 0; j < nblk; j++) {
          const int jb = p.n_off / 64 + (int)j;                // 64-column block of the GEMM's N axis
          if (p.cpb > 0) {                                     // (phase, co): phase (py,px) = pixels (2y+py, 2x+px) of dy
            const int ph = jb / p.cpb, cb = jb - ph * p.cpb;
            wg_tma_4d(&tmDY, base + bs * b_bytes + j * BLK_BYTES, bfull(bs), cb * 64, 2 * ox0 + (ph & 1), 2 * oy0 + (ph >> 1), n0);
          } else {
            wg_tma_4d(&tmDY, base + bs * b_bytes + j * BLK_BYTES, bfull(bs), jb * 64, ox0, oy0, n0);
          }
        }
        if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
        wg_wait(aempty(as), aph ^ 1);
        wg_expect_tx(afull(as), n_boxes * box_bytes);
        for (int b = 0; b < n_boxes; b++) {
          const int r = r_first + b / cblocks, cb = b % cblocks;
          const int c = cb * 64;
          const uint32_t dst = a_base + as * a_bytes + b * q.rb_bytes;
          if (c < p.C1)
            wg_tma_4d(&tmX, dst, afull(as), c, ox0 - p.pad, oy0 - p.pad + r, n0);
          else
            wg_tma_4d(&tmX2, dst, afull(as), c - p.C1, ox0 - p.pad, oy0 - p.pad + r, n0);
        }
        if (++as == p.a_slots) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (wg_elect_one()) {
      // The tensor pipe queues only a few MMAs (tools/mma_rate.cu), so the issuing lane must not pause between groups:
      // the per-M-tile operand offsets (which need integer divisions) are computed ONCE, descriptors are 32-bit adds,
      // and the barriers of the NEXT pixel tile are polled right behind the MMAs of the current one.
      const uint32_t idesc = umma_idesc_f16_mn(p.Cout) | (p.bf16 ? ((1u << 7) | (1u << 10)) : 0u);
      const int nmt = mt1 - mt0;                                       // <= 8 accumulators (512 / Cout)
      uint32_t rel[8];                                                 // (offset in the A slot >> 4) | (LBO >> 4) << 16
#pragma unroll
      for (int i = 0; i < 8; i++) {
        rel[i] = 0;
        if (i < nmt) {
          const int mt = mt0 + i;
          int u0 = 2 * mt, u1 = 2 * mt + 1;
          if (u1 >= p.units) u1 = u0;                                  // odd tail: rows 64..127 are ignored
          const int s0 = (u0 / cblocks) % p.kw, s1 = (u1 / cblocks) % p.kw;
          const uint32_t o0 = box_of(u0) * q.rb_bytes + s0 * 128;
          const uint32_t o1 = box_of(u1) * q.rb_bytes + s1 * 128;
          rel[i] = (o0 >> 4) | ((((o1 - o0) >> 4) & 0x3FFFu) << 16);   // units ascend with the box order: o1 >= o0
        }
      }
      constexpr uint32_t HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);    // SBO 1024 B, version 1, SWIZZLE_128B
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      if (pt1 > pt0) {
        wg_wait(bfull(bs), bph);
        wg_wait(afull(as), aph);
      }
      for (int pt = pt0; pt < pt1; pt++) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b_lo = (((base + bs * b_bytes) & 0x3FFFF) >> 4) | ((uint32_t)(BLK_BYTES >> 4) << 16);
        const uint32_t a_lo = ((a_base + as * a_bytes) & 0x3FFFF) >> 4;
        const uint32_t accum0 = pt > pt0 ? 1u : 0u;
        // (accumulator outer, K slice inner.  The other order -- consecutive MMAs into different accumulators -- measured
        // 16-19 % SLOWER on the 256^2 / 128^2 layers, round 2.)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (i < nmt) {
            const uint32_t d_tmem = tmem_base + i * p.Cout;
#pragma unroll
            for (int kk = 0; kk < 8; kk++)                             // 16 pixels (two 8-row swizzle atoms) per MMA
              wg_mma(d_tmem, ((uint64_t)HI << 32) | (uint64_t)(a_lo + rel[i] + kk * 128),
                     ((uint64_t)HI << 32) | (uint64_t)(b_lo + kk * 128), idesc, accum0 | (uint32_t)kk);
          }
        }
        wg_commit(aempty(as));
        wg_commit(bempty(bs));
        if (++as == p.a_slots) { as = 0; aph ^= 1; }
        if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
        if (pt + 1 < pt1) {                                            // next pixel tile's operands
          wg_wait(bfull(bs), bph);
          wg_wait(afull(as), aph);
        }
      }
      wg_commit(tfull);
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    if (pt1 > pt0) {
      wg_wait(tfull, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int mt = mt0; mt < mt1; mt++) {
        const int k = mt * 128 + row;
        const bool valid = k < p.units * 64;
        float* dst = p.dw + (size_t)k * p.pitch + p.n_off;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (mt - mt0) * p.Cout;
        for (int c0 = 0; c0 < p.Cout; c0 += 32) {
          uint32_t v[32];
          wg_ld32(taddr + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) wg_red4(dst + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wg_encode_fn() {
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
static int wg_encode_act(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int bw, int bh, int bn,
                         int stride = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = wg_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}
static inline int wg_pow2_floor(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

}  // namespace hm

using namespace hm;

// weight gradient of (nearest-2x upsampling -> 5x5 'same' stride-1 convolution) in phase form: `d` is the layer's FORWARD
// descriptor (up = HM_UP_NEAREST2, H x W the low-res source grid, Ho x Wo = 2H x 2W); the result is the gradient of
// the four 3x3 phase filters, dw[(tap3, ci)][(phase, co)] (9*Cin rows, 4*Cout columns; hm_unpack_conv_wgrad mode 8
// folds it onto the 5x5 filter): 36 taps on H x W pixels instead of 25 taps on the materialised 2H x 2W tensor
static bool is_up2_wgrad(const HmConvDesc* d) {
  return d->up == HM_UP_NEAREST2 && d->kh == 5 && d->kw == 5 && d->pad == 2 && d->stride == 1 && !d->transposed &&
         d->C2 == 0 && d->Ho == 2 * d->H && d->Wo == 2 * d->W && d->oH == d->Ho && d->oW == d->Wo && d->os == 1 &&
         !d->ou && !d->ov && d->C1 > 0 && d->C1 % 64 == 0 && d->Cout > 0 && d->Cout % 64 == 0;
}

extern "C" int hm_tc_wgrad_supported(const HmConvDesc* d) {
  if (!d) return 0;
  if ((d->dtype == HM_F16 || d->dtype == HM_BF16X3) && is_up2_wgrad(d)) return 1;
  if ((d->dtype != HM_F16 && d->dtype != HM_BF16X3) || d->transposed || d->up || (d->stride != 1 && d->stride != 2)) return 0;
  if (d->os != 1 || d->ou || d->ov) return 0;
  if (d->C1 % 64 || d->C2 % 64 || d->C1 <= 0) return 0;
  if (d->Cout % 64 || d->Cout <= 0 || (d->Cout > 256 && d->Cout % 256)) return 0;
  if (d->Ho != (d->H + 2 * d->pad - d->kh) / d->stride + 1 || d->Wo != (d->W + 2 * d->pad - d->kw) / d->stride + 1)
    return 0;
  if (d->oH != d->Ho || d->oW != d->Wo) return 0;
  return 1;
}

static int tc_wgrad_tile(const HmConvDesc* d, const void* x1, const void* x2, const void* dy, float* dw, void* stream,
                         int n_off, int ntile);

extern "C" int hm_tc_wgrad(const HmConvDesc* d, const void* x1, const void* x2, const void* dy, float* dw,
                           void* stream) {
  HM_CHECK_ARG(d && x1 && dy && dw, "hm_tc_wgrad: null argument");
  const int ncols = hm_tc_wgrad_supported(d) && is_up2_wgrad(d) ? 4 * d->Cout : d->Cout;       // GEMM N
  if (hm_tc_wgrad_supported(d) && ncols > 256) {
    for (int n_off = 0; n_off < ncols; n_off += 256) {
      int rc = tc_wgrad_tile(d, x1, x2, dy, dw, stream, n_off, 256);
      if (rc) return rc;
    }
    return HM_OK;
  }
  return tc_wgrad_tile(d, x1, x2, dy, dw, stream, 0, ncols);
}

static int tc_wgrad_tile(const HmConvDesc* d, const void* x1, const void* x2, const void* dy, float* dw, void* stream,
                         int n_off, int ntile) {
  HM_CHECK_ARG(d && x1 && dy && dw, "hm_tc_wgrad: null argument");
  if (!hm_tc_wgrad_supported(d)) {
    set_error("hm_tc_wgrad: shape not supported by the tcgen05 path (need fp16, stride 1, C%%64==0, Cout%%64==0, <=256)");
    return HM_ERR_UNSUPPORTED;
  }
  HM_CHECK_ARG(d->C2 == 0 || x2, "hm_tc_wgrad: C2>0 but x2 is null");
  if (!wg_encode_fn()) {
    set_error("hm_tc_wgrad: cuTensorMapEncodeTiled is not available from this driver");
    return HM_ERR_CUDA;
  }
  if (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)dy) & 15) {
    set_error("hm_tc_wgrad: pointers must be 16-byte aligned");
    return HM_ERR_ALIGN;
  }
  const bool up2 = is_up2_wgrad(d);
  const int gH = up2 ? d->H : d->Ho, gW = up2 ? d->W : d->Wo;       // pixel grid of the reduction (low-res in phase form)
  const int kh = up2 ? 3 : d->kh, kw = up2 ? 3 : d->kw, pad = up2 ? 1 : d->pad;
  WgParams p;
  p.B = d->B; p.Ho = gH; p.Wo = gW;
  p.Cin = d->C1 + d->C2; p.C1 = d->C1; p.Cout = ntile;
  p.n_off = n_off; p.pitch = up2 ? 4 * d->Cout : d->Cout;
  p.bf16 = d->dtype == HM_BF16X3 ? 1 : 0;
  p.cpb = up2 ? d->Cout / 64 : 0;
  p.kh = kh; p.kw = kw; p.pad = pad; p.stride = d->stride;
  p.bw = wg_pow2_floor(gW < 128 ? gW : 128);
  p.bh = wg_pow2_floor(gH < 128 / p.bw ? gH : 128 / p.bw);
  p.bn = 128 / (p.bw * p.bh);
  p.tiles_x = (gW + p.bw - 1) / p.bw;
  p.tiles_y = (gH + p.bh - 1) / p.bh;
  p.tiles_n = (d->B + p.bn - 1) / p.bn;
  p.n_ptiles = p.tiles_x * p.tiles_y * p.tiles_n;
  p.units = kh * kw * (p.Cin / 64);
  p.n_mtiles = (p.units + 1) / 2;
  p.acc = 512 / p.Cout;
  p.n_mgroups = (p.n_mtiles + p.acc - 1) / p.acc;
  static int waves = -1;                          // tuning knob: work items per SM (split-K depth)
  if (waves < 0) {
    const char* e = getenv("HMGAN_WG_WAVES");
    waves = (e && atoi(e) >= 1) ? atoi(e) : 1;    // measured: 1 wave beats 2..4 on every layer (fewer partial sums to reduce)
  }
  int z = (waves * num_sms()) / p.n_mgroups;      // ~`waves` waves of work items
  if (z < 1) z = 1;
  if (z > p.n_ptiles) z = p.n_ptiles;
  p.zsplit = z;
  const int b_bytes = (p.Cout / 64) * BLK_BYTES, a_bytes = 2 * BLK_BYTES;
  p.b_slots = (3 * b_bytes + 4 * a_bytes <= 227 * 1024 - 4096) ? 3 : 2;
  int a_slots = (227 * 1024 - 4096 - p.b_slots * b_bytes) / a_bytes;
  if (a_slots > 6) a_slots = 6;
  if (a_slots < 2) {
    set_error("hm_tc_wgrad: not enough shared memory for Cout=%d", p.Cout);
    return HM_ERR_UNSUPPORTED;
  }
  p.a_slots = a_slots;
  p.dw = dw;
  CUtensorMap tmX, tmX2, tmDY;
  int rc = wg_encode_act(&tmX, x1, d->B, d->H, d->W, d->C1, p.bw, p.bh, p.bn, p.stride);
  if (!rc && d->C2) rc = wg_encode_act(&tmX2, x2, d->B, d->H, d->W, d->C2, p.bw, p.bh, p.bn, p.stride);
  if (!d->C2) tmX2 = tmX;
  if (!rc) rc = wg_encode_act(&tmDY, dy, d->B, d->Ho, d->Wo, d->Cout, p.bw, p.bh, p.bn, up2 ? 2 : 1);
  if (rc) {
    set_error("hm_tc_wgrad: cuTensorMapEncodeTiled failed (CUresult %d)", rc);
    return HM_ERR_CUDA;
  }
  // Row-box variant when the pixel tiles are row segments and the filter has several taps per row
  static int rb_enabled = -1;
  if (rb_enabled < 0) {
    const char* e = getenv("HMGAN_TC_ROWBOX");
    rb_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (rb_enabled && d->stride == 1 && p.bw == 128 && p.bh == 1 && p.bn == 1 && kw > 1 && kw <= 9) {
    WgRbParams q;
    q.w = p;
    q.rb_bytes = (((128 + kw - 1) * 128) + 1023) / 1024 * 1024;
    const int cblocks = p.Cin / 64;
    // worst case over m-groups of (filter rows spanned) * cblocks
    int max_boxes = 0;
    for (int mg = 0; mg < p.n_mgroups; mg++) {
      const int mt0 = (int)(((long long)mg * p.n_mtiles) / p.n_mgroups);
      const int mt1 = (int)(((long long)(mg + 1) * p.n_mtiles) / p.n_mgroups);
      const int u_lo = 2 * mt0, u_hi = (2 * mt1 < p.units ? 2 * mt1 : p.units);
      if (u_hi <= u_lo) continue;
      const int r0 = (u_lo / cblocks) / kw, r1 = ((u_hi - 1) / cblocks) / kw;
      const int nb = (r1 - r0 + 1) * cblocks;
      if (nb > max_boxes) max_boxes = nb;
    }
    q.max_boxes = max_boxes;
    const size_t a_rb = (size_t)max_boxes * q.rb_bytes;
    // ring depths: a dy tile (b) and the x boxes (a) of one pixel tile are consumed together, so both rings want >= 3
    // slots to cover the TMA latency (one pixel tile is only ~2000 MMA cycles); fall back to 2 where memory is short
    const size_t budget = 227 * 1024 - 4096;
    static int wg_bslots = -1;
    if (wg_bslots < 0) {
      const char* e = getenv("HMGAN_WG_BSLOTS");
      wg_bslots = (e && atoi(e) >= 2) ? atoi(e) : 3;
    }
    q.w.b_slots = wg_bslots;
    while (q.w.b_slots > 2 && (size_t)q.w.b_slots * b_bytes + 3 * a_rb > budget) q.w.b_slots--;
    int a_sl = (int)((budget - (size_t)q.w.b_slots * b_bytes) / a_rb);
    if (a_sl > 3) a_sl = 3;
    if (a_sl >= 2) {
      q.w.a_slots = a_sl;
      CUtensorMap rX, rX2;
      rc = wg_encode_act(&rX, x1, d->B, d->H, d->W, d->C1, 128 + kw - 1, 1, 1);
      if (!rc && d->C2) rc = wg_encode_act(&rX2, x2, d->B, d->H, d->W, d->C2, 128 + kw - 1, 1, 1);
      if (!d->C2) rX2 = rX;
      if (rc) {
        set_error("hm_tc_wgrad: cuTensorMapEncodeTiled failed for the row box (CUresult %d)", rc);
        return HM_ERR_CUDA;
      }
      const size_t smem_rb = (size_t)q.w.b_slots * b_bytes + (size_t)q.w.a_slots * a_rb + 1024 + 1024;
      static bool rb_attr = false;
      if (!rb_attr) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad_rb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
          set_error("hm_tc_wgrad: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
          return HM_ERR_CUDA;
        }
        rb_attr = true;
      }
      tc_wgrad_rb_kernel<<<p.n_mgroups * p.zsplit, WG_THREADS, smem_rb, (cudaStream_t)stream>>>(rX, rX2, tmDY, q);
      HM_CHECK_LAUNCH("hm_tc_wgrad(row box)");
      return HM_OK;
    }
  }
  const size_t smem = (size_t)p.b_slots * b_bytes + (size_t)p.a_slots * a_bytes + 1024 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("hm_tc_wgrad: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
      return HM_ERR_CUDA;
    }
    attr_set = true;
  }
  tc_wgrad_kernel<<<p.n_mgroups * p.zsplit, WG_THREADS, smem, (cudaStream_t)stream>>>(tmX, tmX2, tmDY, p);
  HM_CHECK_LAUNCH("hm_tc_wgrad");
  return HM_OK;
}
