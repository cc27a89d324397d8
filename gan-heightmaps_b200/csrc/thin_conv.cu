// "Thin" convolutions: layers whose GEMM view has a degenerate K or N and are therefore HBM/L1-bound, not
// tensor-core work: the first discriminator layer (1 -> 64 channels, reference
// architectures/dcgan.py:42 with Cin=1; p2p.py:145,285 likewise) and the last generator layer (64 -> 1,
// dcgan.py:32).  They are specialisations behind the hm_conv_gather / hm_conv_wgrad entry points (same
// contract, same packed layouts); the generic tiled kernels in simt_conv.cu remain the fallback.
//   * thin_in_conv:    Cin_total <= 4.  One thread = one output pixel x 8 output channels; the <=100 taps*Cin
//                      inputs come through L1 (neighbouring pixels share them), the weights from shared memory,
//                      the result leaves as one 16-byte store (coalesced 128 B per pixel for Cout = 64).
//   * thin_in_wgrad:   Cin_total <= 4.  dWp[(tap,ci)][co] = sum_pix x[pix+tap][ci] * dy[pix][co]; one thread =
//                      (tap, 8 output channels), a block streams a slab of pixels, fp32 register accumulators,
//                      one atomic per (block, weight).
//   * thin_out_wgrad:  Cout <= 4.   dWp[(tap,ci)][co] = sum_pix x[pix+tap][ci] * dy[pix][co]; one thread =
//                      (tap, 8 input channels): one 16-byte load of x per pixel, dy broadcast.
#include "hm_common.cuh"

namespace hm {

constexpr int THIN_MAXK = 100;   // taps * Cin handled by the thin-input kernels (5x5x4)

// ---- forward, Cin_total <= 4 ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) thin_in_conv_kernel(HmConvDesc d, const T* __restrict__ x1,
                                                           const T* __restrict__ x2, const T* __restrict__ w,
                                                           const float* __restrict__ bias, T* __restrict__ y) {
  extern __shared__ float ws[];                  // [K][Cout]
  const int Ct = d.C1 + d.C2, K = d.kh * d.kw * Ct;
  for (int i = threadIdx.x; i < K * d.Cout; i += blockDim.x) ws[i] = ldf(w + i);
  __syncthreads();
  const int groups = d.Cout >> 3;
  const long long total = (long long)d.B * d.Ho * d.Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long pix = i / groups;
    const int ox = (int)(pix % d.Wo);
    long long t2 = pix / d.Wo;
    const int oy = (int)(t2 % d.Ho);
    const int n = (int)(t2 / d.Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = bias ? bias[g * 8 + j] : 0.f;
    for (int r = 0; r < d.kh; r++) {
      const int iy = oy * d.stride - d.pad + r;
      if (iy < 0 || iy >= d.H) continue;
      for (int s = 0; s < d.kw; s++) {
        const int ix = ox * d.stride - d.pad + s;
        if (ix < 0 || ix >= d.W) continue;
        const size_t src = ((size_t)n * d.H + iy) * d.W + ix;
        for (int ci = 0; ci < Ct; ci++) {
          const float xv = ci < d.C1 ? ldf(x1 + src * d.C1 + ci) : ldf(x2 + src * d.C2 + (ci - d.C1));
          const float4* wr = reinterpret_cast<const float4*>(ws + ((r * d.kw + s) * Ct + ci) * d.Cout + g * 8);
          const float4 wa = wr[0], wb = wr[1];
          acc[0] = fmaf(xv, wa.x, acc[0]); acc[1] = fmaf(xv, wa.y, acc[1]);
          acc[2] = fmaf(xv, wa.z, acc[2]); acc[3] = fmaf(xv, wa.w, acc[3]);
          acc[4] = fmaf(xv, wb.x, acc[4]); acc[5] = fmaf(xv, wb.y, acc[5]);
          acc[6] = fmaf(xv, wb.z, acc[6]); acc[7] = fmaf(xv, wb.w, acc[7]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = act_fwd(acc[j], d.act, d.slope);
    store8(y + (size_t)pix * d.Cout + g * 8, acc);
  }
}

// ---- weight gradient, Cin_total <= 4 ------------------------------------------------------------------------
// thread = (tap, co8); block streams pixels [p0, p1)
template <typename T>
__global__ void __launch_bounds__(256) thin_in_wgrad_kernel(HmConvDesc d, const T* __restrict__ x1,
                                                            const T* __restrict__ x2, const T* __restrict__ dy,
                                                            float* dw, long long per_block) {
  const int Ct = d.C1 + d.C2, taps = d.kh * d.kw, groups = d.Cout >> 3;
  const int units = taps * groups;
  const int lanes = blockDim.x / units;            // pixel lanes per block (>= 1 by launch)
  const int u = threadIdx.x % units, pl = threadIdx.x / units;
  const long long M = (long long)d.B * d.Ho * d.Wo;
  const long long p0 = (long long)blockIdx.x * per_block, p1 = min(M, p0 + per_block);
  if (pl >= lanes) return;
  const int tap = u / groups, g = u - tap * groups;
  const int r = tap / d.kw, s = tap - r * d.kw;
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[c][j] = 0.f;
  for (long long pix = p0 + pl; pix < p1; pix += lanes) {
    const int ox = (int)(pix % d.Wo);
    long long t2 = pix / d.Wo;
    const int oy = (int)(t2 % d.Ho);
    const int n = (int)(t2 / d.Ho);
    const int iy = oy * d.stride - d.pad + r, ix = ox * d.stride - d.pad + s;
    if (iy < 0 || iy >= d.H || ix < 0 || ix >= d.W) continue;
    float g8[8];
    load8(dy + (size_t)pix * d.Cout + g * 8, g8);
    const size_t src = ((size_t)n * d.H + iy) * d.W + ix;
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
      if (ci < Ct) {
        const float xv = ci < d.C1 ? ldf(x1 + src * d.C1 + ci) : ldf(x2 + src * d.C2 + (ci - d.C1));
#pragma unroll
        for (int j = 0; j < 8; j++) acc[ci][j] = fmaf(xv, g8[j], acc[ci][j]);
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < 4; ci++)
    if (ci < Ct) {
#pragma unroll
      for (int j = 0; j < 8; j++) atomicAdd(dw + (size_t)(tap * Ct + ci) * d.Cout + g * 8 + j, acc[ci][j]);
    }
}


// ---- forward, Cin = 1, stride 1, 4 consecutive output pixels x 8 channels per thread ---------------------------
// Fully unrolled over the KHxKW taps; per filter row the thread loads KW+3 neighbouring inputs once and reuses
// them for its 4 pixels; weights come from shared memory as two float4 per tap.
template <typename T, int KH, int KW>
__global__ void __launch_bounds__(256) thin_in_conv_c1_kernel(HmConvDesc d, const T* __restrict__ x,
                                                              const T* __restrict__ w, const float* __restrict__ bias,
                                                              T* __restrict__ y) {
  extern __shared__ float ws[];                  // [KH*KW][Cout]
  for (int i = threadIdx.x; i < KH * KW * d.Cout; i += blockDim.x) ws[i] = ldf(w + i);
  __syncthreads();
  const int groups = d.Cout >> 3;
  const int wq = (d.Wo + 3) >> 2;                // 4-pixel strips per row
  const long long total = (long long)d.B * d.Ho * wq * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long t1 = i / groups;
    const int sx = (int)(t1 % wq);
    t1 /= wq;
    const int oy = (int)(t1 % d.Ho);
    const int n = (int)(t1 / d.Ho);
    const int ox0 = sx * 4;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[p][j] = bias ? bias[g * 8 + j] : 0.f;
#pragma unroll
    for (int r = 0; r < KH; r++) {
      const int iy = oy - d.pad + r;
      if (iy < 0 || iy >= d.H) continue;
      const T* row = x + ((size_t)n * d.H + iy) * d.W;
      float xv[KW + 3];
#pragma unroll
      for (int q = 0; q < KW + 3; q++) {
        const int ix = ox0 - d.pad + q;
        xv[q] = (ix >= 0 && ix < d.W) ? ldf(row + ix) : 0.f;
      }
#pragma unroll
      for (int s = 0; s < KW; s++) {
        const float4* wr = reinterpret_cast<const float4*>(ws + (r * KW + s) * d.Cout + g * 8);
        const float4 wa = wr[0], wb = wr[1];
#pragma unroll
        for (int p = 0; p < 4; p++) {
          const float v = xv[p + s];
          acc[p][0] = fmaf(v, wa.x, acc[p][0]); acc[p][1] = fmaf(v, wa.y, acc[p][1]);
          acc[p][2] = fmaf(v, wa.z, acc[p][2]); acc[p][3] = fmaf(v, wa.w, acc[p][3]);
          acc[p][4] = fmaf(v, wb.x, acc[p][4]); acc[p][5] = fmaf(v, wb.y, acc[p][5]);
          acc[p][6] = fmaf(v, wb.z, acc[p][6]); acc[p][7] = fmaf(v, wb.w, acc[p][7]);
        }
      }
    }
    T* dst = y + (((size_t)n * d.Ho + oy) * d.Wo + ox0) * d.Cout + g * 8;
#pragma unroll
    for (int p = 0; p < 4; p++) {
      if (ox0 + p < d.Wo) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[p][j] = act_fwd(acc[p][j], d.act, d.slope);
        store8(dst + (size_t)p * d.Cout, acc[p]);
      }
    }
  }
}

// ---- weight gradient, Cin_total <= 4, stride 1: thread = (filter row r, input channel ci, 8 output channels) ---
// holds KW x 8 fp32 accumulators; per pixel: one 16-byte load of dy, KW scalar loads of x.
template <typename T, int KW>
__global__ void __launch_bounds__(256) thin_in_wgrad_row_kernel(HmConvDesc d, const T* __restrict__ x1,
                                                                const T* __restrict__ x2, const T* __restrict__ dy,
                                                                float* dw, int rows_per_block) {
  const int Ct = d.C1 + d.C2, groups = d.Cout >> 3;
  const int units = d.kh * Ct * groups;
  const int lanes = blockDim.x / units;
  const int u = threadIdx.x % units, pl = threadIdx.x / units;
  if (pl >= lanes) return;
  const int g = u % groups;
  const int ci = (u / groups) % Ct;
  const int r = u / (groups * Ct);
  const T* src = ci < d.C1 ? x1 : x2;
  const int C = ci < d.C1 ? d.C1 : d.C2;
  const int cc = ci < d.C1 ? ci : ci - d.C1;
  float acc[KW][8];
#pragma unroll
  for (int s = 0; s < KW; s++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[s][j] = 0.f;
  const int total_rows = d.B * d.Ho;
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(total_rows, row0 + rows_per_block);
  for (int row = row0; row < row1; row++) {
    const int n = row / d.Ho, oy = row - n * d.Ho;
    const int iy = oy - d.pad + r;
    if (iy < 0 || iy >= d.H) continue;
    const T* xrow = src + ((size_t)n * d.H + iy) * d.W * C + cc;
    const T* grow = dy + (size_t)row * d.Wo * d.Cout + g * 8;
    for (int ox0 = pl * 4; ox0 < d.Wo; ox0 += lanes * 4) {     // 4 consecutive pixels per iteration
      float xv[KW + 3];
#pragma unroll
      for (int q = 0; q < KW + 3; q++) {
        const int ix = ox0 - d.pad + q;
        xv[q] = (ix >= 0 && ix < d.W) ? ldf(xrow + (size_t)ix * C) : 0.f;
      }
      float g8[4][8];
#pragma unroll
      for (int p = 0; p < 4; p++) {
        if (ox0 + p < d.Wo) {
          load8(grow + (size_t)(ox0 + p) * d.Cout, g8[p]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; j++) g8[p][j] = 0.f;
        }
      }
#pragma unroll
      for (int s = 0; s < KW; s++)
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
          for (int j = 0; j < 8; j++) acc[s][j] = fmaf(xv[p + s], g8[p][j], acc[s][j]);
    }
  }
#pragma unroll
  for (int s = 0; s < KW; s++)
#pragma unroll
    for (int j = 0; j < 8; j++)
      atomicAdd(dw + (size_t)((r * KW + s) * Ct + ci) * d.Cout + g * 8 + j, acc[s][j]);
}

// ---- weight gradient, Cout <= 4 -------------------------------------------------------------------------------
// thread = (tap, ci8); x may be read through the virtual nearest/bilinear 2x upsampling.
template <typename T>
__device__ __forceinline__ void gather8(const HmConvDesc& d, const T* __restrict__ src, int C, int n, int iy, int ix,
                                        int c0, float* v) {
  const size_t img = (size_t)n * d.H;
  if (d.up != HM_UP_BILINEAR2) {
    const int sh = d.up ? 1 : 0;
    load8(src + ((img + (iy >> sh)) * d.W + (ix >> sh)) * C + c0, v);
    return;
  }
  const int y0 = iy >> 1, x0 = ix >> 1;
  const int y1 = (iy & 1) ? min(y0 + 1, d.H - 1) : y0;
  const int x1 = (ix & 1) ? min(x0 + 1, d.W - 1) : x0;
  float a[8], b[8], e[8], f[8];
  load8(src + ((img + y0) * d.W + x0) * C + c0, a);
  load8(src + ((img + y0) * d.W + x1) * C + c0, b);
  load8(src + ((img + y1) * d.W + x0) * C + c0, e);
  load8(src + ((img + y1) * d.W + x1) * C + c0, f);
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = 0.25f * ((a[j] + b[j]) + (e[j] + f[j]));
}

template <typename T>
__global__ void __launch_bounds__(256) thin_out_wgrad_kernel(HmConvDesc d, const T* __restrict__ x1,
                                                             const T* __restrict__ x2, const T* __restrict__ dy,
                                                             float* dw, int per_block, int phases) {
  // blockIdx.y = output phase (oy&1, ox&1) when the gradient of a nearest-2x + 5x5 layer is taken as four 3x3
  // problems on the low-res source (hm_up2conv_wgrad_phases); 0 otherwise.
  if (phases) {
    d.ou = blockIdx.y >> 1;
    d.ov = blockIdx.y & 1;
  }
  const int Ct = d.C1 + d.C2, taps = d.kh * d.kw, cg = Ct >> 3;
  const int units = taps * cg;
  dw += (size_t)blockIdx.y * taps * Ct * d.Cout;
  const int M = d.B * d.Ho * d.Wo;
  const int p0 = blockIdx.x * per_block, p1 = min(M, p0 + per_block);
  const int sh = d.up ? 1 : 0;
  const int Hv = d.H << sh, Wv = d.W << sh;
  const int lanes = blockDim.x / units > 0 ? blockDim.x / units : 1;
  for (int u0 = 0; u0 < units; u0 += blockDim.x) {          // one pass unless taps*Cin/8 > 256
    const int u = u0 + (lanes > 1 ? threadIdx.x % units : threadIdx.x);
    const int pl = lanes > 1 ? threadIdx.x / units : 0;
    if (u >= units || pl >= lanes) continue;
    const int tap = u / cg, c0 = (u - tap * cg) * 8;
    const int r = tap / d.kw, s = tap - r * d.kw;
    const T* src = c0 < d.C1 ? x1 : x2;
    const int C = c0 < d.C1 ? d.C1 : d.C2;
    const int cc = c0 < d.C1 ? c0 : c0 - d.C1;
    float acc[4][8];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[c][j] = 0.f;
    for (int pix = p0 + pl; pix < p1; pix += lanes) {
      const int ox = pix % d.Wo;
      const int t2 = pix / d.Wo;
      const int oy = t2 % d.Ho;
      const int n = t2 / d.Ho;
      const int iy = oy * d.stride - d.pad + r, ix = ox * d.stride - d.pad + s;
      if (iy >= 0 && iy < Hv && ix >= 0 && ix < Wv) {
        float xv[8];
        gather8<T>(d, src, C, n, iy, ix, cc, xv);
        const size_t op = ((size_t)n * d.oH + (oy * d.os + d.ou)) * d.oW + (ox * d.os + d.ov);
#pragma unroll
        for (int co = 0; co < 4; co++)
          if (co < d.Cout) {
            const float g = ldf(dy + op * d.Cout + co);
#pragma unroll
            for (int j = 0; j < 8; j++) acc[co][j] = fmaf(xv[j], g, acc[co][j]);
          }
      }
    }
#pragma unroll
    for (int co = 0; co < 4; co++)
      if (co < d.Cout) {
#pragma unroll
        for (int j = 0; j < 8; j++) atomicAdd(dw + (size_t)(tap * Ct + c0 + j) * d.Cout + co, acc[co][j]);
      }
  }
}

// ---- re-layouts that turn the thin layers into tensor-core GEMMs ----------------------------------------------
// im2col of a ONE-channel image: Xc[p][t] = x[p + tap t - pad] (t < kh*kw), 0 for the padding taps up to 64.
// A 1->Cout kxk convolution is then the 1x1 convolution Xc[.,64] x Wt[Cout][64] (forward) and its weight
// gradient the GEMM Xc^T dy, both on the tcgen05 kernels.  One thread = one pixel x 8 taps (16-byte store).
__global__ void __launch_bounds__(256) im2col_c1_kernel(const __half* __restrict__ x, __half* __restrict__ xc, int B,
                                                        int H, int W, int kh, int kw, int pad) {
  // a thread always serves the same 8 taps (grid stride is a multiple of 8): decode them once
  const int g = threadIdx.x & 7;
  int dr[8], ds[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int t = g * 8 + j;
    dr[j] = t < kh * kw ? t / kw - pad : (1 << 20);          // out-of-range marker: never inside the image
    ds[j] = t < kh * kw ? t % kw - pad : 0;
  }
  const unsigned total = (unsigned)B * H * W;                  // pixels (< 2^31 for every supported size)
  const unsigned step = (gridDim.x * blockDim.x) >> 3;
  unsigned pix = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  for (; pix < total; pix += step) {
    const unsigned ox = pix % W;
    const unsigned t2 = pix / W;
    const unsigned oy = t2 % H;
    const __half* img = x + (size_t)(t2 - oy) * W;             // start of this image
    uint4 out;
    __half* o = reinterpret_cast<__half*>(&out);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int iy = (int)oy + dr[j], ix = (int)ox + ds[j];
      o[j] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? img[iy * W + ix] : __float2half(0.f);
    }
    *reinterpret_cast<uint4*>(xc + (size_t)pix * 64 + g * 8) = out;
  }
}

// General thin-source im2col (1..4 source channels, possibly split over two tensors = a ConcatLayer, stride 1 or 2):
// xc[B,Ho,Wo,64], xc[p][(r*kw+s)*C + c] = src[p*stride + (r,s) - pad][c], zero beyond kh*kw*C and outside the image.
// One thread = one output pixel x 8 columns (16-byte store); a thread always serves the same 8 columns.
__global__ void __launch_bounds__(256) im2col_thin_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2,
                                                          __half* __restrict__ xc, int B, int H, int W, int C1, int C2,
                                                          int kh, int kw, int stride, int pad, int Ho, int Wo) {
  const int g = threadIdx.x & 7;
  const int Ct = C1 + C2;
  int dr[8], ds[8], dc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = g * 8 + j;
    const int tap = k / Ct;
    dc[j] = k - tap * Ct;
    dr[j] = k < kh * kw * Ct ? tap / kw - pad : (1 << 20);     // out-of-range marker: never inside the image
    ds[j] = k < kh * kw * Ct ? tap % kw - pad : 0;
  }
  const unsigned total = (unsigned)B * Ho * Wo;
  const unsigned step = (gridDim.x * blockDim.x) >> 3;
  unsigned pix = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  for (; pix < total; pix += step) {
    const unsigned ox = pix % Wo;
    const unsigned t2 = pix / Wo;
    const unsigned oy = t2 % Ho;
    const unsigned b = t2 / Ho;
    const size_t img = (size_t)b * H * W;
    uint4 out;
    __half* o = reinterpret_cast<__half*>(&out);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int iy = (int)oy * stride + dr[j], ix = (int)ox * stride + ds[j];
      __half v = __float2half(0.f);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const size_t q = img + (size_t)iy * W + ix;
        v = dc[j] < C1 ? x1[q * C1 + dc[j]] : x2[q * C2 + (dc[j] - C1)];
      }
      o[j] = v;
    }
    *reinterpret_cast<uint4*>(xc + (size_t)pix * 64 + g * 8) = out;
  }
}

// space-to-depth of a thin gradient, zero-padded to 64 channels: out[q][ph*Co+co] = dy[2q + (ph>>1, ph&1)][co]
__global__ void s2d_pad64_kernel(const __half* __restrict__ dy, __half* __restrict__ out, int B, int h, int w, int Co) {
  const long long total = (long long)B * h * w * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    long long pix = i >> 3;
    const int qx = (int)(pix % w);
    long long t2 = pix / w;
    const int qy = (int)(t2 % h);
    const int n = (int)(t2 / h);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int c = g * 8 + j;
      const int ph = c / Co, co = c - ph * Co;
      v[j] = (c < 4 * Co)
                 ? __half2float(dy[(((size_t)n * 2 * h + 2 * qy + (ph >> 1)) * (2 * w) + 2 * qx + (ph & 1)) * Co + co])
                 : 0.f;
    }
    store8(out + (size_t)pix * 64 + g * 8, v);
  }
}

// ---- dispatch helpers called from simt_conv.cu ----------------------------------------------------------------
bool thin_in_conv_launch(const HmConvDesc* d, const void* x1, const void* x2, const void* w, const float* bias,
                         void* y, void* y2, cudaStream_t st) {
  const int Ct = d->C1 + d->C2;
  if (Ct > 4 || d->transposed || d->up || d->Cout % 8 || d->os != 1 || d->split != d->Cout || d->accumulate ||
      y2 != nullptr || y == nullptr || d->kh * d->kw * Ct > THIN_MAXK || d->kh * d->kw * Ct * d->Cout * 4 > 48 * 1024)
    return false;
  if (((uintptr_t)y & 15) != 0) return false;
  const long long total = (long long)d->B * d->Ho * d->Wo * (d->Cout / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)d->kh * d->kw * Ct * d->Cout * sizeof(float);
  if (Ct == 1 && d->stride == 1 && d->kh == d->kw && (d->kh == 5 || d->kh == 3)) {
    const long long tot4 = (long long)d->B * d->Ho * ((d->Wo + 3) / 4) * (d->Cout / 8);
    long long b4 = (tot4 + 255) / 256;
    if (b4 > cap) b4 = cap;
#define LAUNCH_C1(TT, KK)                                                                                  \
  thin_in_conv_c1_kernel<TT, KK, KK><<<(unsigned)b4, 256, smem, st>>>(*d, (const TT*)x1, (const TT*)w, bias, (TT*)y)
    if (d->dtype == HM_F32) {
      if (d->kh == 5) LAUNCH_C1(float, 5); else LAUNCH_C1(float, 3);
    } else {
      if (d->kh == 5) LAUNCH_C1(__half, 5); else LAUNCH_C1(__half, 3);
    }
#undef LAUNCH_C1
    return true;
  }
  if (d->dtype == HM_F32)
    thin_in_conv_kernel<float><<<(unsigned)blocks, 256, smem, st>>>(*d, (const float*)x1, (const float*)x2,
                                                                    (const float*)w, bias, (float*)y);
  else
    thin_in_conv_kernel<__half><<<(unsigned)blocks, 256, smem, st>>>(*d, (const __half*)x1, (const __half*)x2,
                                                                     (const __half*)w, bias, (__half*)y);
  return true;
}

void thin_out_wgrad_go(const HmConvDesc* d, const void* x1, const void* x2, const void* dy, float* dw,
                              cudaStream_t st, int phases) {
  const long long M = (long long)d->B * d->Ho * d->Wo;
  long long blocks = (long long)num_sms() * 4 / (phases ? 2 : 1);
  long long per = (M + blocks - 1) / blocks;
  if (per < 64) per = 64;
  blocks = (M + per - 1) / per;
  dim3 grid((unsigned)blocks, phases ? 4 : 1);
  if (d->dtype == HM_F32)
    thin_out_wgrad_kernel<float><<<grid, 256, 0, st>>>(*d, (const float*)x1, (const float*)x2, (const float*)dy, dw,
                                                       (int)per, phases);
  else
    thin_out_wgrad_kernel<__half><<<grid, 256, 0, st>>>(*d, (const __half*)x1, (const __half*)x2, (const __half*)dy,
                                                        dw, (int)per, phases);
}

bool thin_wgrad_launch(const HmConvDesc* d, const void* x1, const void* x2, const void* dy, float* dw,
                       cudaStream_t st) {
  const int Ct = d->C1 + d->C2;
  const long long M = (long long)d->B * d->Ho * d->Wo;
  if (d->transposed) return false;
  if (Ct <= 4 && !d->up && d->os == 1 && d->Cout % 8 == 0 && d->kh * d->kw * (d->Cout / 8) <= 256 &&
      ((uintptr_t)dy & 15) == 0) {
    const int runits = d->kh * Ct * (d->Cout / 8);
    if (d->stride == 1 && runits <= 256 && (d->kw == 5 || d->kw == 3)) {
      const int total_rows = d->B * d->Ho;
      int blocks = num_sms() * 4;
      int rpb = (total_rows + blocks - 1) / blocks;
      if (rpb < 1) rpb = 1;
      blocks = (total_rows + rpb - 1) / rpb;
#define LAUNCH_ROW(TT, KK)                                                                                   \
  thin_in_wgrad_row_kernel<TT, KK><<<blocks, 256, 0, st>>>(*d, (const TT*)x1, (const TT*)x2, (const TT*)dy, dw, rpb)
      if (d->dtype == HM_F32) {
        if (d->kw == 5) LAUNCH_ROW(float, 5); else LAUNCH_ROW(float, 3);
      } else {
        if (d->kw == 5) LAUNCH_ROW(__half, 5); else LAUNCH_ROW(__half, 3);
      }
#undef LAUNCH_ROW
      return true;
    }
    const int units = d->kh * d->kw * (d->Cout / 8);
    const int lanes = 256 / units;
    long long blocks = (long long)num_sms() * 4;
    long long per = (M + blocks - 1) / blocks;
    if (per < lanes) per = lanes;
    blocks = (M + per - 1) / per;
    if (d->dtype == HM_F32)
      thin_in_wgrad_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(*d, (const float*)x1, (const float*)x2,
                                                                    (const float*)dy, dw, per);
    else
      thin_in_wgrad_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(*d, (const __half*)x1, (const __half*)x2,
                                                                     (const __half*)dy, dw, per);
    return true;
  }
  if (d->Cout <= 4 && d->C1 % 8 == 0 && d->C2 % 8 == 0 && ((uintptr_t)x1 & 15) == 0 && ((uintptr_t)x2 & 15) == 0 &&
      M < (1LL << 31)) {
    thin_out_wgrad_go(d, x1, x2, dy, dw, st, 0);
    return true;
  }
  return false;
}

}  // namespace hm

extern "C" int hm_im2col_c1(const void* x, void* xc, int B, int H, int W, int kh, int kw, int pad, void* stream) {
  HM_CHECK_ARG(x && xc && B > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && kh * kw <= 64 && pad >= 0,
               "hm_im2col_c1: bad argument");
  HM_CHECK_ARG((((uintptr_t)xc) & 15) == 0, "hm_im2col_c1: output must be 16-byte aligned");
  const long long total = (long long)B * H * W * 8;
  long long blocks = (total + 255) / 256, cap = (long long)hm::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  hm::im2col_c1_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)xc, B, H, W, kh,
                                                                          kw, pad);
  HM_CHECK_LAUNCH("hm_im2col_c1");
  return HM_OK;
}

extern "C" int hm_im2col_thin(const void* x1, const void* x2, void* xc, int B, int H, int W, int C1, int C2, int kh,
                              int kw, int stride, int pad, int Ho, int Wo, void* stream) {
  HM_CHECK_ARG(x1 && xc && B > 0 && H > 0 && W > 0 && C1 > 0 && C2 >= 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0 &&
                   Ho > 0 && Wo > 0, "hm_im2col_thin: bad argument");
  HM_CHECK_ARG(kh * kw * (C1 + C2) <= 64, "hm_im2col_thin: kh*kw*(C1+C2) = %d exceeds 64 columns", kh * kw * (C1 + C2));
  HM_CHECK_ARG(C2 == 0 || x2, "hm_im2col_thin: C2 > 0 but x2 is null");
  HM_CHECK_ARG((((uintptr_t)xc) & 15) == 0, "hm_im2col_thin: output must be 16-byte aligned");
  const long long total = (long long)B * Ho * Wo * 8;
  long long blocks = (total + 255) / 256, cap = (long long)hm::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  hm::im2col_thin_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)x1, (const __half*)x2, (__half*)xc,
                                                                            B, H, W, C1, C2, kh, kw, stride, pad, Ho, Wo);
  HM_CHECK_LAUNCH("hm_im2col_thin");
  return HM_OK;
}

extern "C" int hm_s2d_pad64(const void* dy, void* out, int B, int h, int w, int Co, void* stream) {
  HM_CHECK_ARG(dy && out && B > 0 && h > 0 && w > 0 && Co > 0 && 4 * Co <= 64, "hm_s2d_pad64: bad argument");
  HM_CHECK_ARG((((uintptr_t)out) & 15) == 0, "hm_s2d_pad64: output must be 16-byte aligned");
  const long long total = (long long)B * h * w * 8;
  long long blocks = (total + 255) / 256, cap = (long long)hm::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  hm::s2d_pad64_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)dy, (__half*)out, B, h, w, Co);
  HM_CHECK_LAUNCH("hm_s2d_pad64");
  return HM_OK;
}

// Weight gradient of (nearest-2x upsampling -> 5x5 'same' convolution) with <= 4 output channels, as four 3x3
// problems on the LOW-RES source: dw_phases[phase][(dy*3+dx)*Cin + ci][co] += sum_q x[q + (dy-1,dx-1)][ci] *
// dy[2q + phase][co].  `d` is the layer's forward descriptor (up = HM_UP_NEAREST2, 5x5, pad 2); fold the result
// onto the 5x5 filter with hm_unpack_conv_wgrad(mode 9).  fp32, atomically accumulated, caller zeroes.
extern "C" int hm_up2conv_wgrad_phases(const HmConvDesc* d, const void* x, const void* dy, float* dw_phases,
                                       void* stream) {
  HM_CHECK_ARG(d && x && dy && dw_phases, "hm_up2conv_wgrad_phases: null argument");
  HM_CHECK_ARG(d->up == HM_UP_NEAREST2 && d->kh == 5 && d->kw == 5 && d->pad == 2 && d->stride == 1 && !d->transposed &&
                   d->C2 == 0 && d->Ho == 2 * d->H && d->Wo == 2 * d->W && d->oH == d->Ho && d->oW == d->Wo,
               "hm_up2conv_wgrad_phases: descriptor is not a nearest-2x + 5x5 'same' convolution");
  HM_CHECK_ARG(d->Cout >= 1 && d->Cout <= 4 && d->C1 % 8 == 0 && (((uintptr_t)x) & 15) == 0,
               "hm_up2conv_wgrad_phases: needs Cout <= 4, Cin %% 8 == 0 and a 16-byte aligned source");
  HmConvDesc q = *d;
  q.up = 0; q.kh = q.kw = 3; q.pad = 1;
  q.Ho = d->H; q.Wo = d->W;            // logical grid = low-res pixels; dy is read at (2q + phase)
  q.os = 2; q.ou = 0; q.ov = 0;
  hm::thin_out_wgrad_go(&q, x, nullptr, dy, dw_phases, (cudaStream_t)stream, 1);
  HM_CHECK_LAUNCH("hm_up2conv_wgrad_phases");
  return HM_OK;
}
