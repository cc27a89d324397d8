// Gather-GEMM convolutions on CUDA cores (any shape, fp32 accumulate).
//
// This is the parity-mode arithmetic (fp32 tensors) and the fast-mode fallback for
// the layers that are not tensor-core shaped (Cin=1, Cout in {1,3}, 2x2 kernels).
// Restates theano CorrMM / CorrMM_gradInputs / CorrMM_gradWeights as reached from
// lasagne Conv2DLayer / Deconv2DLayer / DenseLayer (reference architectures/dcgan.py:16,
// 22,32,42,50; architectures/p2p.py:20-24).
#include <stdlib.h>

#include "hm_common.cuh"

namespace hm {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct Site {
  int n, oy, ox;
};

// Value of the virtual (upsampled, concatenated) source at logical site/tap/channel.
template <typename T>
__device__ __forceinline__ float gather(const HmConvDesc& d, const T* __restrict__ x1,
                                        const T* __restrict__ x2, const Site& s, int k, int Ct) {
  int tap = k / Ct;
  int ci = k - tap * Ct;
  int r = tap / d.kw;
  int c = tap - r * d.kw;
  int iy, ix;
  if (!d.transposed) {
    iy = s.oy * d.stride - d.pad + r;
    ix = s.ox * d.stride - d.pad + c;
  } else {
    int ty = s.oy + d.pad - r, tx = s.ox + d.pad - c;
    if (ty < 0 || tx < 0) return 0.f;
    if (d.stride > 1) {
      if ((ty % d.stride) || (tx % d.stride)) return 0.f;
      ty /= d.stride;
      tx /= d.stride;
    }
    iy = ty;
    ix = tx;
  }
  const int sh = d.up ? 1 : 0;
  if (iy < 0 || ix < 0 || iy >= (d.H << sh) || ix >= (d.W << sh)) return 0.f;
  const T* src;
  int C;
  if (ci < d.C1) {
    src = x1;
    C = d.C1;
  } else {
    src = x2;
    C = d.C2;
    ci -= d.C1;
  }
  const size_t img = (size_t)s.n * d.H;
  if (d.up != HM_UP_BILINEAR2) {
    int py = iy >> sh, px = ix >> sh;
    return ldf(src + ((img + py) * d.W + px) * C + ci);
  }
  // theano bilinear_upsampling, ratio 2: y[2m]=x[m], y[2m+1]=(x[m]+x[min(m+1,n-1)])/2
  int y0 = iy >> 1, x0 = ix >> 1;
  int y1 = (iy & 1) ? min(y0 + 1, d.H - 1) : y0;
  int x1i = (ix & 1) ? min(x0 + 1, d.W - 1) : x0;
  float a = ldf(src + ((img + y0) * d.W + x0) * C + ci);
  float b = ldf(src + ((img + y0) * d.W + x1i) * C + ci);
  float e = ldf(src + ((img + y1) * d.W + x0) * C + ci);
  float f = ldf(src + ((img + y1) * d.W + x1i) * C + ci);
  return 0.25f * ((a + b) + (e + f));
}

__device__ __forceinline__ Site decode_site(const HmConvDesc& d, long long m) {
  Site s;
  int hw = d.Ho * d.Wo;
  s.n = (int)(m / hw);
  int rem = (int)(m - (long long)s.n * hw);
  s.oy = rem / d.Wo;
  s.ox = rem - s.oy * d.Wo;
  return s;
}

template <typename T>
__global__ void __launch_bounds__(NT) conv_gather_kernel(HmConvDesc d, const T* __restrict__ x1,
                                                         const T* __restrict__ x2,
                                                         const T* __restrict__ w,
                                                         const float* __restrict__ bias, T* y, T* y2) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int Ct = d.C1 + d.C2;
  const int K = d.kh * d.kw * Ct;
  const long long M = (long long)d.B * d.Ho * d.Wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load role: row a_m, 4 consecutive k
  const int a_m = tid >> 2, a_k = (tid & 3) * 4;
  const long long am = m0 + a_m;
  const bool a_ok = am < M;
  Site as = decode_site(d, a_ok ? am : 0);
  // B-load role: k row b_k, 4 consecutive co
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int k = k0 + a_k + j;
      As[a_k + j][a_m] = (a_ok && k < K) ? gather<T>(d, x1, x2, as, k, Ct) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int k = k0 + b_k, co = n0 + b_n + j;
      Bs[b_k][b_n + j] = (k < K && co < d.Cout) ? ldf(w + (size_t)k * d.Cout + co) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; i++) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    Site s = decode_site(d, m);
    size_t pix = ((size_t)s.n * d.oH + (s.oy * d.os + d.ou)) * d.oW + (s.ox * d.os + d.ov);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int co = n0 + tx * 4 + j;
      if (co >= d.Cout) continue;
      float v = acc[i][j] + (bias ? bias[co] : 0.f);
      v = act_fwd(v, d.act, d.slope);
      T* dst;
      int accum;
      if (co < d.split) {
        dst = y + pix * d.split + co;
        accum = d.accumulate & 1;
      } else {
        dst = y2 + pix * (d.Cout - d.split) + (co - d.split);
        accum = d.accumulate & 2;
      }
      if (co < d.split ? (y == nullptr) : (y2 == nullptr)) continue;
      if (accum) v += ldf(dst);
      stf(dst, v);
    }
  }
}

// dWp[k][co] += sum_sites gather(site,k) * dy[site][co]
template <typename T>
__global__ void __launch_bounds__(NT) conv_wgrad_kernel(HmConvDesc d, const T* __restrict__ x1,
                                                        const T* __restrict__ x2,
                                                        const T* __restrict__ dy, float* dw,
                                                        long long sites_per_z) {
  __shared__ float As[BK][BM + 4];  // [site][k]
  __shared__ float Bs[BK][BN + 4];  // [site][co]
  const int tid = threadIdx.x;
  const int Ct = d.C1 + d.C2;
  const int K = d.kh * d.kw * Ct;
  const long long M = (long long)d.B * d.Ho * d.Wo;
  const int k0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long s_begin = (long long)blockIdx.z * sites_per_z;
  const long long s_end = min(M, s_begin + sites_per_z);

  const int l_s = tid >> 4, l_c = (tid & 15) * 4;  // both tiles: site row, 4 consecutive cols
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (long long s0 = s_begin; s0 < s_end; s0 += BK) {
    long long m = s0 + l_s;
    bool ok = m < s_end;
    Site st = decode_site(d, ok ? m : 0);
    size_t pix = ((size_t)st.n * d.oH + (st.oy * d.os + d.ou)) * d.oW + (st.ox * d.os + d.ov);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int k = k0 + l_c + j;
      As[l_s][l_c + j] = (ok && k < K) ? gather<T>(d, x1, x2, st, k, Ct) : 0.f;
      int co = n0 + l_c + j;
      Bs[l_s][l_c + j] = (ok && co < d.Cout) ? ldf(dy + pix * d.Cout + co) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int k = k0 + ty * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int co = n0 + tx * 4 + j;
      if (co < d.Cout) atomicAdd(dw + (size_t)k * d.Cout + co, acc[i][j]);
    }
  }
}

// ---- weight packing --------------------------------------------------------
template <typename T>
__device__ __forceinline__ void pack_elem(const float* __restrict__ w, T* __restrict__ wp, int mode, int cout,
                                          int cin, int kh, int kw, int u, int v, long long i) {
  float val;
  if (mode == 0) {  // Wp[(r*kw+s)*cin+ci][co] = W[co][ci][kh-1-r][kw-1-s]
    int co = (int)(i % cout);
    long long k = i / cout;
    int ci = (int)(k % cin);
    int tap = (int)(k / cin);
    int r = tap / kw, s = tap % kw;
    val = w[(((size_t)co * cin + ci) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)];
  } else if (mode == 1) {  // Wp[(r*kw+s)*cout+co][ci] = W[co][ci][kh-1-r][kw-1-s]
    int ci = (int)(i % cin);
    long long k = i / cin;
    int co = (int)(k % cout);
    int tap = (int)(k / cout);
    int r = tap / kw, s = tap % kw;
    val = w[(((size_t)co * cin + ci) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)];
  } else if (mode == 2) {  // deconv W (cin,cout,kh,kw), one phase: Wp[ci][co] = W[ci][co][kh-1-u][kw-1-v]
    int co = (int)(i % cout);
    int ci = (int)(i / cout);
    val = w[(((size_t)ci * cout + co) * kh + (kh - 1 - u)) * kw + (kw - 1 - v)];
  } else if (mode == 3) {  // deconv W: Wp[(u*kw+v)*cout+co][ci] = W[ci][co][kh-1-u][kw-1-v]
    int ci = (int)(i % cin);
    long long k = i / cin;
    int co = (int)(k % cout);
    int tap = (int)(k / cout);
    int uu = tap / kw, vv = tap % kw;
    val = w[(((size_t)ci * cout + co) * kh + (kh - 1 - uu)) * kw + (kw - 1 - vv)];
  } else if (mode == 5) {  // tcgen05 forward pack, K-major: Wt[(r*kw+s)][co][ci] = W[co][ci][kh-1-r][kw-1-s]
    int ci = (int)(i % cin);
    long long k = i / cin;
    int co = (int)(k % cout);
    int tap = (int)(k / cout);
    int r = tap / kw, s = tap % kw;
    val = w[(((size_t)co * cin + ci) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)];
  } else if (mode == 6) {  // tcgen05 input-gradient pack (a forward correlation of dy): Wt[(r*kw+s)][ci][co] = W[co][ci][r][s]
    int co = (int)(i % cout);
    long long k = i / cout;
    int ci = (int)(k % cin);
    int tap = (int)(k / cin);
    int r = tap / kw, s = tap % kw;
    val = w[(((size_t)co * cin + ci) * kh + r) * kw + s];
  } else if (mode == 7) {  // input gradient of a stride-1 conv as a FORWARD gather of dy (pad' = k-1-pad):
                           // Wp[(r*kw+s)*cout+co][ci] = W[co][ci][r][s]
    int ci = (int)(i % cin);
    long long k = i / cin;
    int co = (int)(k % cout);
    int tap = (int)(k / cout);
    int r = tap / kw, s = tap % kw;
    val = w[(((size_t)co * cin + ci) * kh + r) * kw + s];
  } else if (mode == 8) {
    // nearest-2x upsampling + 5x5 'same' convolution == four 3x3 'same' convolutions on the LOW-RES tensor, one per
    // output phase (oy&1, ox&1):  Wt[(dy*3+dx)][phase*cout+co][ci] = sum_{r in R(py,dy), s in R(px,dx)} W[co][ci][4-r][4-s]
    // with R(p,d) = { r in 0..4 : floor((p + r - 2) / 2) == d - 1 }   (K-major tcgen05 pack)
    int ci = (int)(i % cin);
    long long k = i / cin;
    int nn = (int)(k % (4 * cout));
    int tap = (int)(k / (4 * cout));
    int ph = nn / cout, co = nn - ph * cout;
    int py = ph >> 1, px = ph & 1, dy_ = tap / 3 - 1, dx_ = tap % 3 - 1;
    val = 0.f;
    for (int r = 0; r < 5; r++) {
      if (((py + r - 2 + 4) >> 1) - 2 != dy_) continue;
      for (int s = 0; s < 5; s++) {
        if (((px + s - 2 + 4) >> 1) - 2 != dx_) continue;
        val += w[(((size_t)co * cin + ci) * 5 + (4 - r)) * 5 + (4 - s)];
      }
    }
  } else if (mode == 12) {
    // input gradient of a 3x3 stride-2 pad-1 convolution as a 2x2-tap convolution of dy with N = (phase, ci):
    // Wt[(dy*2+dx)][phase*cin+ci][co] = W[co][ci][2-r][2-s], r = r(py,dy), s = r(px,dx) with r(0,0)=1, r(1,0)=2, r(1,1)=0;
    // (py,dy) = (0,1) has no tap -> 0.   (K-major tcgen05 pack, K = co)
    int co = (int)(i % cout);
    long long k = i / cout;
    int nn = (int)(k % (4 * cin));
    int tap = (int)(k / (4 * cin));
    int ph = nn / cin, ci = nn - ph * cin;
    int py = ph >> 1, px = ph & 1, dy_ = tap >> 1, dx_ = tap & 1;
    int r = py == 0 ? (dy_ == 0 ? 1 : -1) : (dy_ == 0 ? 2 : 0);
    int s = px == 0 ? (dx_ == 0 ? 1 : -1) : (dx_ == 0 ? 2 : 0);
    val = (r >= 0 && s >= 0) ? w[(((size_t)co * cin + ci) * 3 + (2 - r)) * 3 + (2 - s)] : 0.f;
  } else if (mode == 11) {
    // one-channel-input convolution as a 1x1 convolution over the im2col tensor (hm_im2col_c1):
    // Wt[co][t] = W[co][0][kh-1-r][kw-1-s] for tap t = r*kw+s < kh*kw, 0 up to 64   (K-major, K = 64)
    int t = (int)(i % 64);
    int co = (int)(i / 64);
    int r = t / kw, s = t % kw;
    val = t < kh * kw ? w[((size_t)co * kh + (kh - 1 - r)) * kw + (kw - 1 - s)] : 0.f;
  } else if (mode == 14) {
    // hm_c1s2_conv, plain form: input gradient of (nearest-2x -> 5x5 'same' conv, Cout == 1) as a 6x6 stride-2 gather of
    // the one-channel dy:  Wk[ci][u*6+v] = sum_{r in R(u&1, 2-(u>>1)), s in R(v&1, 2-(v>>1))} W[0][ci][4-r][4-s],
    // R(p,d) as in mode 8;  columns 36..63 are zero.   (K-major, K = 64)
    int k = (int)(i % 64);
    int ci = (int)(i / 64);
    val = 0.f;
    if (k < 36) {
      int uu = k / 6, vv = k % 6;
      int py = uu & 1, px = vv & 1, dy_ = 2 - (uu >> 1), dx_ = 2 - (vv >> 1);     // phase, 3x3 tap (0..2)
      for (int r = 0; r < 5; r++) {
        if (((py + r - 2 + 4) >> 1) - 2 + 1 != dy_) continue;
        for (int q = 0; q < 5; q++) {
          if (((px + q - 2 + 4) >> 1) - 2 + 1 != dx_) continue;
          val += w[((size_t)ci * 5 + (4 - r)) * 5 + (4 - q)];
        }
      }
    }
  } else if (mode == 21) {
    // DenseLayer W (in, out) as a 1x1 tensor-core convolution: Wt[co][ci] = W[ci][co]   (K-major, K = ci)
    int ci = (int)(i % cin);
    int co = (int)(i / cin);
    val = w[(size_t)ci * cout + co];
  } else if (mode == 20) {
    // input gradient of (nearest-2x -> 5x5 'same' conv) as ONE 6x6 stride-2 pad-2 convolution of dy on the low-res grid:
    //   dx[q][ci] = sum_{u,v<6} sum_co dy[2q-2+(u,v)][co] * Wt[(u*6+v)][ci][co]
    // with  Wt[(u*6+v)][ci][co] = sum_{r in R(u&1, 2-(u>>1)), s in R(v&1, 2-(v>>1))} W[co][ci][4-r][4-s]   (R as in mode 8:
    // the adjoint of the four 3x3 phase filters; mode 14 is its Cout == 1 case).   (K-major tcgen05 pack, K = co)
    int co = (int)(i % cout);
    long long k = i / cout;
    int ci = (int)(k % cin);
    int tap = (int)(k / cin);
    int uu = tap / 6, vv = tap % 6;
    int py = uu & 1, px = vv & 1, dy_ = 2 - (uu >> 1), dx_ = 2 - (vv >> 1);
    val = 0.f;
    for (int r = 0; r < 5; r++) {
      if (((py + r - 2 + 4) >> 1) - 2 + 1 != dy_) continue;
      for (int q = 0; q < 5; q++) {
        if (((px + q - 2 + 4) >> 1) - 2 + 1 != dx_) continue;
        val += w[(((size_t)co * cin + ci) * 5 + (4 - r)) * 5 + (4 - q)];
      }
    }
  } else if (mode == 15) {
    // hm_c1s2_conv, pooled form: conv5x5 'same' (Cin == 1) evaluated at the four positions d = (dy,dx) of every 2x2
    // pooling window from the window's 6x6 patch:  Wk[d*cout+co][u*6+v] = W[co][0][4-(u-dy)][4-(v-dx)] where the
    // tap (u-dy, v-dx) lies inside the 5x5 filter, else 0;  columns 36..63 are zero.   (K-major, K = 64)
    int k = (int)(i % 64);
    int nn = (int)(i / 64);
    int dd = nn / cout, co = nn - dd * cout;
    val = 0.f;
    if (k < 36) {
      int r = k / 6 - (dd >> 1), q = k % 6 - (dd & 1);
      if (r >= 0 && r < 5 && q >= 0 && q < 5) val = w[((size_t)co * 5 + (4 - r)) * 5 + (4 - q)];
    }
  } else if (mode == 16) {
    // hm_c1s2_bwd input-gradient operand: mode 15 transposed per window position, Wk2[d*64+k][co] = Wk15[d*cout+co][k]
    int co = (int)(i % cout);
    int row = (int)(i / cout);
    int dd = row / 64, k = row % 64;
    val = 0.f;
    if (k < 36) {
      int r = k / 6 - (dd >> 1), q = k % 6 - (dd & 1);
      if (r >= 0 && r < 5 && q >= 0 && q < 5) val = w[((size_t)co * 5 + (4 - r)) * 5 + (4 - q)];
    }
  } else if (mode == 19) {
    // thin-source convolution as a 1x1 convolution over hm_im2col_thin's tensor:
    // Wt[co][(r*kw+s)*cin + ci] = W[co][ci][kh-1-r][kw-1-s] for (r*kw+s)*cin + ci < kh*kw*cin, 0 up to 64  (K-major)
    int k = (int)(i % 64);
    int co = (int)(i / 64);
    val = 0.f;
    if (k < kh * kw * cin) {
      int tap = k / cin, ci = k - tap * cin;
      int r = tap / kw, q = tap % kw;
      val = w[(((size_t)co * cin + ci) * kh + (kh - 1 - r)) * kw + (kw - 1 - q)];
    }
  } else if (mode == 17) {
    // Deconv2DLayer W (cin,cout,2,2), stride 2, all four output phases at once (hm_tc_conv, transposed == 2):
    // Wt[(u*2+v)*cout + co][ci] = W[ci][co][1-u][1-v]   (K-major, K = ci)
    int ci = (int)(i % cin);
    int nn = (int)(i / cin);
    int ph = nn / cout, co = nn - ph * cout;
    val = w[(((size_t)ci * cout + co) * 2 + (1 - (ph >> 1))) * 2 + (1 - (ph & 1))];
  } else if (mode == 18) {
    // its input gradient as a 1x1 convolution of s2d(dy) (hm_s2d_pad64: 64 channels, (phase, co) in the first 4*cout):
    // Wt[ci][k] = W[ci][co][1-u][1-v] for k = (u*2+v)*cout + co < 4*cout, else 0   (K-major, K = 64)
    int k = (int)(i % 64);
    int ci = (int)(i / 64);
    val = 0.f;
    if (k < 4 * cout) {
      int ph = k / cout, co = k - ph * cout;
      val = w[(((size_t)ci * cout + co) * 2 + (1 - (ph >> 1))) * 2 + (1 - (ph & 1))];
    }
  } else if (mode == 22) {
    // mode 18 without the padding: Wt[ci][(u*2+v)*cout + co] = W[ci][co][1-u][1-v]   (K-major, K = 4*cout): the input
    // gradient of a 2x2 deconvolution of a 1x1 input as a 1x1 convolution over dy viewed as [B,1,1,(u,v,co)]
    int k = (int)(i % (4 * cout));
    int ci = (int)(i / (4 * cout));
    int ph = k / cout, co = k - ph * cout;
    val = w[(((size_t)ci * cout + co) * 2 + (1 - (ph >> 1))) * 2 + (1 - (ph & 1))];
  } else {
    val = w[i];
  }
  stf(wp + i, val);
}

// (A tiled shared-memory variant of the mode 5 / 6 packs and of the mode-0 un-pack was written and measured in round 2:
// it coalesces the strided side of the permutation, but inside the captured step the ~50 packs are launch-bound and
// partly hidden on the side stream -- 12.55 vs 12.67 ms DCGAN, 18.19 vs 18.26 ms joint, 11.20 vs 11.24 ms pix2pix, i.e. no
// gain -- so it was removed again.)
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wp, int mode, int cout,
                                   int cin, int kh, int kw, int u, int v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pack_elem<T>(w, wp, mode, cout, cin, kh, kw, u, v, i);
}

// every pack of a network in ONE launch: blockIdx.y selects the job (a table in device memory), blockIdx.x strides over
// its elements.  A training step re-packs ~35 small weight tensors after the update; as separate launches they cost more
// in launch gaps than in work.
template <typename T>
__global__ void pack_weight_multi_kernel(const HmPackJob* __restrict__ jobs) {
  const HmPackJob j = jobs[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += (long long)gridDim.x * blockDim.x)
    pack_elem<T>(j.w, (T*)j.wp, j.mode, j.cout, j.cin, j.kh, j.kw, j.u, j.v, i);
}

// packed fp32 gradient -> master layout gradient
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int mode,
                                    int cout, int cin, int kh, int kw, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 0) {  // i indexes master W[co][ci][a][b]
    int b = (int)(i % kw);
    long long t = i / kw;
    int a = (int)(t % kh);
    t /= kh;
    int ci = (int)(t % cin);
    int co = (int)(t / cin);
    int r = kh - 1 - a, s = kw - 1 - b;
    dw[i] = dwp[((size_t)(r * kw + s) * cin + ci) * cout + co];
  } else if (mode == 2) {  // master deconv W[ci][co][a][b]; packed as kh*kw phase blocks [tap][ci][co]
    int b = (int)(i % kw);
    long long t = i / kw;
    int a = (int)(t % kh);
    t /= kh;
    int co = (int)(t % cout);
    int ci = (int)(t / cout);
    int u = kh - 1 - a, v = kw - 1 - b;
    dw[i] = dwp[((size_t)(u * kw + v) * cin + ci) * cout + co];
  } else if (mode == 8 || mode == 9 || mode == 10) {
    // fold the gradients of the four 3x3 phase filters back onto the 5x5 master filter (adjoint of pack mode 8).
    // mode 8: dwp[(tap3, ci)][(phase, co)]   (tcgen05 weight-gradient layout)
    // mode 9: dwp[phase][(tap3, ci)][co]     (one thin weight-gradient launch per phase)
    // mode 10: dwp[(tap3, ci)][64], column = phase*cout+co  (tcgen05 weight gradient against hm_s2d_pad64)
    int b = (int)(i % 5);
    long long t = i / 5;
    int a = (int)(t % 5);
    t /= 5;
    int ci = (int)(t % cin);
    int co = (int)(t / cin);
    int r = 4 - a, s = 4 - b;
    float acc = 0.f;
    for (int py = 0; py < 2; py++)
      for (int px = 0; px < 2; px++) {
        int dy_ = ((py + r - 2 + 4) >> 1) - 2 + 1, dx_ = ((px + s - 2 + 4) >> 1) - 2 + 1;   // 0..2
        int tap = dy_ * 3 + dx_, ph = py * 2 + px;
        acc += mode == 8    ? dwp[((size_t)(tap * cin + ci) * 4 + ph) * cout + co]
               : mode == 9 ? dwp[((size_t)ph * 9 * cin + tap * cin + ci) * cout + co]
                           : dwp[(size_t)(tap * cin + ci) * 64 + ph * cout + co];
      }
    dw[i] = acc;
  } else if (mode == 14) {
    // adjoint of pack mode 14 (Cout == 1): master W[0][ci][a][b] from dwp[ci][64] = the gradient of Wk[ci][u*6+v]
    // (hm_c1s2_wgrad); tap (r,s) = (4-a, 4-b) enters the 6x6 entry of either phase once
    int b = (int)(i % 5);
    long long t = i / 5;
    int a = (int)(t % 5);
    int ci = (int)(t / 5);
    int r = 4 - a, s = 4 - b;
    float acc = 0.f;
    for (int py = 0; py < 2; py++)
      for (int px = 0; px < 2; px++) {
        int dy_ = ((py + r - 2 + 4) >> 1) - 2 + 1, dx_ = ((px + s - 2 + 4) >> 1) - 2 + 1;   // 0..2
        int uu = 2 * (2 - dy_) + py, vv = 2 * (2 - dx_) + px;
        acc += dwp[(size_t)ci * 64 + uu * 6 + vv];
      }
    dw[i] = acc;
  } else if (mode == 17) {
    // master deconv W[ci][co][a][b] (2x2) from dwp[ci][ld], column (u*2+v)*cout + co, (u,v) = (1-a, 1-b): the tensor-core
    // weight gradient of x against hm_s2d_pad64(dy) (ld = 64) or, for the wide 1x1-input form, against dy (ld = 4*cout)
    int b = (int)(i % 2);
    long long t = i / 2;
    int a = (int)(t % 2);
    t /= 2;
    int co = (int)(t % cout);
    int ci = (int)(t / cout);
    const int ld = 4 * cout <= 64 ? 64 : 4 * cout;
    dw[i] = dwp[(size_t)ci * ld + ((1 - a) * 2 + (1 - b)) * cout + co];
  } else {
    dw[i] = dwp[i];
  }
}

}  // namespace hm

using namespace hm;

static int check_desc(const HmConvDesc* d, const char* who) {
  HM_CHECK_ARG(d != nullptr, "%s: null descriptor", who);
  HM_CHECK_ARG(d->dtype == HM_F32 || d->dtype == HM_F16, "%s: bad dtype %d", who, d->dtype);
  HM_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->C1 > 0 && d->C2 >= 0, "%s: bad source shape", who);
  HM_CHECK_ARG(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->pad >= 0, "%s: bad kernel geometry", who);
  HM_CHECK_ARG(d->Ho > 0 && d->Wo > 0 && d->Cout > 0, "%s: bad output grid", who);
  HM_CHECK_ARG(d->os >= 1 && d->ou >= 0 && d->ov >= 0 && (d->Ho - 1) * d->os + d->ou < d->oH &&
                   (d->Wo - 1) * d->os + d->ov < d->oW,
               "%s: output scatter exceeds the physical tensor", who);
  HM_CHECK_ARG(d->split > 0 && d->split <= d->Cout, "%s: bad channel split", who);
  HM_CHECK_ARG(!(d->transposed && d->up), "%s: virtual upsampling only on the forward gather", who);
  return HM_OK;
}

extern "C" int hm_conv_gather(const HmConvDesc* d, const void* x1, const void* x2, const void* w,
                              const float* bias, void* y, void* y2, void* stream) {
  int rc = check_desc(d, "hm_conv_gather");
  if (rc) return rc;
  HM_CHECK_ARG(x1 && w && (y || y2), "hm_conv_gather: null tensor");
  HM_CHECK_ARG(d->C2 == 0 || x2, "hm_conv_gather: C2>0 but x2 is null");
  long long M = (long long)d->B * d->Ho * d->Wo;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d->Cout + BN - 1) / BN));
  cudaStream_t st = (cudaStream_t)stream;
  if (thin_in_conv_launch(d, x1, x2, w, bias, y, y2, st)) {
    HM_CHECK_LAUNCH("hm_conv_gather(thin input)");
    return HM_OK;
  }
  if (d->dtype == HM_F32)
    conv_gather_kernel<float><<<grid, NT, 0, st>>>(*d, (const float*)x1, (const float*)x2,
                                                   (const float*)w, bias, (float*)y, (float*)y2);
  else
    conv_gather_kernel<__half><<<grid, NT, 0, st>>>(*d, (const __half*)x1, (const __half*)x2,
                                                    (const __half*)w, bias, (__half*)y, (__half*)y2);
  HM_CHECK_LAUNCH("hm_conv_gather");
  return HM_OK;
}

extern "C" int hm_conv_wgrad(const HmConvDesc* d, const void* x1, const void* x2, const void* dy,
                             float* dw, void* stream) {
  int rc = check_desc(d, "hm_conv_wgrad");
  if (rc) return rc;
  HM_CHECK_ARG(x1 && dy && dw, "hm_conv_wgrad: null tensor");
  HM_CHECK_ARG(d->C2 == 0 || x2, "hm_conv_wgrad: C2>0 but x2 is null");
  HM_CHECK_ARG(!d->transposed, "hm_conv_wgrad: descriptor must be a forward gather");
  if (thin_wgrad_launch(d, x1, x2, dy, dw, (cudaStream_t)stream)) {
    HM_CHECK_LAUNCH("hm_conv_wgrad(thin)");
    return HM_OK;
  }
  int K = d->kh * d->kw * (d->C1 + d->C2);
  long long M = (long long)d->B * d->Ho * d->Wo;
  unsigned gx = (K + BM - 1) / BM, gy = (d->Cout + BN - 1) / BN;
  long long want = (4LL * num_sms() + gx * gy - 1) / (gx * gy);
  long long maxz = (M + 4 * BK - 1) / (4 * BK);
  long long gz = want < 1 ? 1 : (want > maxz ? maxz : want);
  if (gz > 65535) gz = 65535;
  long long per = (M + gz - 1) / gz;
  per = (per + BK - 1) / BK * BK;
  gz = (M + per - 1) / per;
  dim3 grid(gx, gy, (unsigned)gz);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == HM_F32)
    conv_wgrad_kernel<float><<<grid, NT, 0, st>>>(*d, (const float*)x1, (const float*)x2,
                                                  (const float*)dy, dw, per);
  else
    conv_wgrad_kernel<__half><<<grid, NT, 0, st>>>(*d, (const __half*)x1, (const __half*)x2,
                                                   (const __half*)dy, dw, per);
  HM_CHECK_LAUNCH("hm_conv_wgrad");
  return HM_OK;
}

static long long pack_count(int mode, int cout, int cin, int kh, int kw) {
  long long n = (mode == 2) ? (long long)cin * cout : (long long)cout * cin * kh * kw;
  if (mode == 8) n = 36LL * cout * cin;
  if (mode == 11) n = 64LL * cout;
  if (mode == 12) n = 16LL * cout * cin;
  if (mode == 14) n = 64LL * cin;
  if (mode == 15 || mode == 16) n = 256LL * cout;
  if (mode == 17 || mode == 22) n = 4LL * cout * cin;
  if (mode == 18) n = 64LL * cin;
  if (mode == 19) n = 64LL * cout;
  if (mode == 20) n = 36LL * cout * cin;
  return n;
}

extern "C" long long hm_pack_conv_weight_count(int mode, int cout, int cin, int kh, int kw) {
  return pack_count(mode, cout, cin, kh, kw);
}

extern "C" int hm_pack_conv_weight_multi(const HmPackJob* jobs_dev, int n_jobs, long long max_n, int dst_dtype,
                                         void* stream) {
  HM_CHECK_ARG(jobs_dev && n_jobs > 0 && n_jobs <= 65535 && max_n > 0, "hm_pack_conv_weight_multi: bad argument");
  HM_CHECK_ARG(dst_dtype == HM_F32 || dst_dtype == HM_F16, "hm_pack_conv_weight_multi: bad dtype %d", dst_dtype);
  long long bx = (max_n + 255) / 256;
  if (bx > 64) bx = 64;                                   // grid-stride inside a job
  dim3 grid((unsigned)bx, (unsigned)n_jobs);
  cudaStream_t st = (cudaStream_t)stream;
  if (dst_dtype == HM_F32)
    pack_weight_multi_kernel<float><<<grid, 256, 0, st>>>(jobs_dev);
  else
    pack_weight_multi_kernel<__half><<<grid, 256, 0, st>>>(jobs_dev);
  HM_CHECK_LAUNCH("hm_pack_conv_weight_multi");
  return HM_OK;
}

extern "C" int hm_pack_conv_weight(const float* w, void* wp, int mode, int cout, int cin, int kh, int kw,
                                   int u, int v, int dst_dtype, void* stream) {
  HM_CHECK_ARG(w && wp, "hm_pack_conv_weight: null pointer");
  HM_CHECK_ARG((mode >= 0 && mode <= 8) || mode == 11 || mode == 12 || (mode >= 14 && mode <= 22),
               "hm_pack_conv_weight: bad mode %d", mode);
  HM_CHECK_ARG(mode != 14 || (cout == 1 && kh == 5 && kw == 5), "hm_pack_conv_weight: mode 14 needs Cout == 1 and a 5x5 filter");
  HM_CHECK_ARG((mode != 15 && mode != 16) || (cin == 1 && kh == 5 && kw == 5),
               "hm_pack_conv_weight: modes 15/16 need Cin == 1 and a 5x5 filter");
  HM_CHECK_ARG(mode != 16 || cout == 64, "hm_pack_conv_weight: mode 16 needs Cout == 64");
  HM_CHECK_ARG((mode != 17 && mode != 18 && mode != 22) || (kh == 2 && kw == 2),
               "hm_pack_conv_weight: modes 17/18/22 are defined for 2x2 filters");
  HM_CHECK_ARG(mode != 18 || 4 * cout <= 64, "hm_pack_conv_weight: mode 18 needs 4*Cout <= 64");
  HM_CHECK_ARG(mode != 19 || kh * kw * cin <= 64, "hm_pack_conv_weight: mode 19 needs kh*kw*Cin <= 64");
  HM_CHECK_ARG(mode != 12 || (kh == 3 && kw == 3), "hm_pack_conv_weight: mode 12 is defined for 3x3 filters");
  HM_CHECK_ARG(mode != 11 || (cin == 1 && kh * kw <= 64), "hm_pack_conv_weight: mode 11 needs Cin == 1 and <= 64 taps");
  HM_CHECK_ARG((mode != 8 && mode != 20) || (kh == 5 && kw == 5), "hm_pack_conv_weight: modes 8 and 20 are defined for 5x5 filters");
  const long long n = pack_count(mode, cout, cin, kh, kw);
  unsigned blocks = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dst_dtype == HM_F32)
    pack_weight_kernel<float><<<blocks, 256, 0, st>>>(w, (float*)wp, mode, cout, cin, kh, kw, u, v, n);
  else
    pack_weight_kernel<__half><<<blocks, 256, 0, st>>>(w, (__half*)wp, mode, cout, cin, kh, kw, u, v, n);
  HM_CHECK_LAUNCH("hm_pack_conv_weight");
  return HM_OK;
}

extern "C" int hm_unpack_conv_wgrad(const float* dwp, float* dw, int mode, int cout, int cin, int kh,
                                    int kw, void* stream) {
  HM_CHECK_ARG(dwp && dw, "hm_unpack_conv_wgrad: null pointer");
  HM_CHECK_ARG(mode == 0 || mode == 2 || mode == 4 || ((mode == 8 || mode == 9 || mode == 10) && kh == 5 && kw == 5) ||
                   (mode == 14 && kh == 5 && kw == 5 && cout == 1) ||
                   (mode == 17 && kh == 2 && kw == 2),
               "hm_unpack_conv_wgrad: bad mode %d", mode);
  long long n = (long long)cout * cin * kh * kw;
  unpack_wgrad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dwp, dw, mode, cout,
                                                                                     cin, kh, kw, n);
  HM_CHECK_LAUNCH("hm_unpack_conv_wgrad");
  return HM_OK;
}
