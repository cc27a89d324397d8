// HBM-bound kernels of the training step: BatchNorm statistics/apply/backward,
// pooling, resampling adjoints, activation backward, losses, optimisers, layout.
// All are coalesced over the innermost (channel) dimension of NHWC tensors, use
// warp-shuffle reductions, and run grid-stride with grids sized from the SM count.
#include <stdarg.h>

#include "hm_common.cuh"

namespace hm {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// the *_v8u / *_v8h variants below (rows in flight / hoisted per-channel constants) are the default: measured on B200
// -0.25 ms per DCGAN step with bit-identical results; HMGAN_EW_HOIST=0 selects the plain kernels
static inline bool ew_hoist() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HMGAN_EW_HOIST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static inline unsigned ew_grid(long long n, int threads = 256, int per_sm = 8) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ---------------------------------------------------------------------------
// Column reductions over x[M,C]: each block owns a slab of rows; thread t handles
// channel (t % CT) ... and row lane (t / CT); partial sums go to double atomics.
// ---------------------------------------------------------------------------
constexpr int RED_THREADS = 256;

template <typename T, int MODE>  // MODE 0: sum,sumsq of x ; 1: sum only (col_sum)
__global__ void __launch_bounds__(RED_THREADS) col_reduce_kernel(const T* __restrict__ x, long long M, int C,
                                                                double* sums, float* fsum) {
  __shared__ float sh0[RED_THREADS], sh1[RED_THREADS];
  const int ct = C < RED_THREADS ? C : RED_THREADS;  // channels covered per pass
  const int lanes = RED_THREADS / ct;                 // row lanes
  const int tc = threadIdx.x % ct, tr = threadIdx.x / ct;
  const bool active = tr < lanes;
  for (int c0 = 0; c0 < C; c0 += ct) {
    int c = c0 + tc;
    float s0 = 0.f, s1 = 0.f;
    if (active && c < C) {
      for (long long m = (long long)blockIdx.x * lanes + tr; m < M; m += (long long)gridDim.x * lanes) {
        float v = ldf(x + (size_t)m * C + c);
        s0 += v;
        if (MODE == 0) s1 += v * v;
      }
    }
    sh0[threadIdx.x] = s0;
    sh1[threadIdx.x] = s1;
    __syncthreads();
    if (tr == 0 && c < C) {
      for (int l = 1; l < lanes; l++) {
        s0 += sh0[l * ct + tc];
        s1 += sh1[l * ct + tc];
      }
      if (MODE == 0) {
        atomicAdd(sums + c, (double)s0);
        atomicAdd(sums + C + c, (double)s1);
      } else {
        atomicAdd(fsum + c, s0);
      }
    }
    __syncthreads();
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long M, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* running_mean, float* running_inv_std, float eps, float alpha,
                                   int update_running, float* mean, float* inv_std, float* scale,
                                   float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float m, s;
  if (sums) {
    double mu = sums[c] / (double)M;
    double var = sums[C + c] / (double)M - mu * mu;
    if (var < 0) var = 0;
    m = (float)mu;
    s = (float)(1.0 / sqrt(var + (double)eps));
    if (update_running) {
      running_mean[c] = (1.f - alpha) * running_mean[c] + alpha * m;
      running_inv_std[c] = (1.f - alpha) * running_inv_std[c] + alpha * s;
    }
  } else {
    m = running_mean[c];
    s = running_inv_std[c];
  }
  if (mean) mean[c] = m;
  if (inv_std) inv_std[c] = s;
  float sc = gamma[c] * s;
  scale[c] = sc;
  shift[c] = beta[c] - m * sc;
}

template <typename T>
__global__ void bn_apply_act_kernel(const T* __restrict__ x, T* __restrict__ a, long long n, int C,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    int act, float slope) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float v = ldf(x + i) * scale[c] + shift[c];
    stf(a + i, act_fwd(v, act, slope));
  }
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
    bn_bwd_reduce_kernel(const T* __restrict__ da, const T* __restrict__ a, const T* __restrict__ x, long long M,
                         int C, const float* __restrict__ mean, const float* __restrict__ inv_std, int act,
                         float slope, double* red) {
  __shared__ float sh0[RED_THREADS], sh1[RED_THREADS];
  const int ct = C < RED_THREADS ? C : RED_THREADS;
  const int lanes = RED_THREADS / ct;
  const int tc = threadIdx.x % ct, tr = threadIdx.x / ct;
  const bool active = tr < lanes;
  for (int c0 = 0; c0 < C; c0 += ct) {
    int c = c0 + tc;
    float s0 = 0.f, s1 = 0.f;
    if (active && c < C) {
      float mu = mean[c], is = inv_std[c];
      for (long long m = (long long)blockIdx.x * lanes + tr; m < M; m += (long long)gridDim.x * lanes) {
        size_t i = (size_t)m * C + c;
        float g = ldf(da + i) * act_grad_from_out(ldf(a + i), act, slope);
        float xh = (ldf(x + i) - mu) * is;
        s0 += g;
        s1 += g * xh;
      }
    }
    sh0[threadIdx.x] = s0;
    sh1[threadIdx.x] = s1;
    __syncthreads();
    if (tr == 0 && c < C) {
      for (int l = 1; l < lanes; l++) {
        s0 += sh0[l * ct + tc];
        s1 += sh1[l * ct + tc];
      }
      atomicAdd(red + c, (double)s0);
      atomicAdd(red + C + c, (double)s1);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ da, const T* __restrict__ a, const T* __restrict__ x,
                                    T* __restrict__ dx, long long M, int C, const float* __restrict__ mean,
                                    const float* __restrict__ inv_std, const float* __restrict__ gamma, int act,
                                    float slope, const double* __restrict__ red, float* dgamma, float* dbeta) {
  const long long n = M * C;
  const float invM = 1.f / (float)M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float r0 = (float)red[c], r1 = (float)red[C + c];
    float is = inv_std[c];
    float g = ldf(da + i) * act_grad_from_out(ldf(a + i), act, slope);
    float xh = (ldf(x + i) - mean[c]) * is;
    stf(dx + i, gamma[c] * is * (g - r0 * invM - xh * r1 * invM));
    if (i < C) {
      if (dgamma) dgamma[c] = r1;
      if (dbeta) dbeta[c] = r0;
    }
  }
}

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* dx, long long n, int act,
                               float slope, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = ldf(dy + i) * act_grad_from_out(ldf(y + i), act, slope);
    if (accumulate) v += ldf(dx + i);
    stf(dx + i, v);
  }
}

template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ p, uint8_t* __restrict__ idx, int B,
                                    int H, int W, int C) {
  const int Hp = H >> 1, Wp = W >> 1;
  const long long n = (long long)B * Hp * Wp * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int px = (int)(t % Wp);
    t /= Wp;
    int py = (int)(t % Hp);
    int b = (int)(t / Hp);
    const T* base = x + (((size_t)b * H + 2 * py) * W + 2 * px) * C + c;
    float v0 = ldf(base), v1 = ldf(base + C), v2 = ldf(base + (size_t)W * C), v3 = ldf(base + (size_t)W * C + C);
    float m = v0;
    int k = 0;
    if (v1 > m) { m = v1; k = 1; }
    if (v2 > m) { m = v2; k = 2; }
    if (v3 > m) { m = v3; k = 3; }
    stf(p + i, m);
    idx[i] = (uint8_t)k;
  }
}

template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ dp, const T* __restrict__ p,
                                    const uint8_t* __restrict__ idx, T* __restrict__ dx, int B, int H, int W,
                                    int C, int act, float slope) {
  const int Hp = H >> 1, Wp = W >> 1;
  const long long n = (long long)B * Hp * Wp * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int px = (int)(t % Wp);
    t /= Wp;
    int py = (int)(t % Hp);
    int b = (int)(t / Hp);
    float g = ldf(dp + i) * act_grad_from_out(ldf(p + i), act, slope);
    int k = idx[i];
    T* base = dx + (((size_t)b * H + 2 * py) * W + 2 * px) * C + c;
    stf(base, k == 0 ? g : 0.f);
    stf(base + C, k == 1 ? g : 0.f);
    stf(base + (size_t)W * C, k == 2 ? g : 0.f);
    stf(base + (size_t)W * C + C, k == 3 ? g : 0.f);
  }
}

// dx[b,y,x,c] (+)= sum over the up-res sites that read x[b,y,x,c]
template <typename T>
__global__ void upsample2_bwd_kernel(const T* __restrict__ dy, T* dx, int B, int H, int W, int C, int mode,
                                     int accumulate) {
  const long long n = (long long)B * H * W * C;
  const int H2 = 2 * H, W2 = 2 * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int x = (int)(t % W);
    t /= W;
    int y = (int)(t % H);
    int b = (int)(t / H);
    const T* img = dy + (size_t)b * H2 * W2 * C + c;
    float acc = 0.f;
    if (mode == HM_UP_NEAREST2) {
      for (int a = 0; a < 2; a++)
        for (int e = 0; e < 2; e++) acc += ldf(img + ((size_t)(2 * y + a) * W2 + (2 * x + e)) * C);
    } else {
      // 1-D adjoint weights: up[2m] <- x[m] (1); up[2m+1] <- x[m] (.5), x[min(m+1,n-1)] (.5)
      // so x[m] receives: up[2m]*1 + up[2m+1]*.5 + up[2m-1]*.5 (m>0) + up[2n-1]*.5 extra when m==n-1
      int ys[3], xs[3];
      float wy[3], wx[3];
      int ny = 0, nx = 0;
      ys[ny] = 2 * y; wy[ny++] = 1.f;
      ys[ny] = 2 * y + 1; wy[ny++] = (y == H - 1) ? 1.f : .5f;
      if (y > 0) { ys[ny] = 2 * y - 1; wy[ny++] = .5f; }
      xs[nx] = 2 * x; wx[nx++] = 1.f;
      xs[nx] = 2 * x + 1; wx[nx++] = (x == W - 1) ? 1.f : .5f;
      if (x > 0) { xs[nx] = 2 * x - 1; wx[nx++] = .5f; }
      for (int a = 0; a < ny; a++)
        for (int e = 0; e < nx; e++) acc += wy[a] * wx[e] * ldf(img + ((size_t)ys[a] * W2 + xs[e]) * C);
    }
    if (accumulate) acc += ldf(dx + i);
    stf(dx + i, acc);
  }
}

template <typename T>
__global__ void upsample2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C,
                                     int mode) {
  const int H2 = 2 * H, W2 = 2 * W;
  const long long n = (long long)B * H2 * W2 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int ox = (int)(t % W2);
    t /= W2;
    int oy = (int)(t % H2);
    int b = (int)(t / H2);
    const T* img = x + (size_t)b * H * W * C + c;
    int y0 = oy >> 1, x0 = ox >> 1;
    float v;
    if (mode == HM_UP_NEAREST2) {
      v = ldf(img + ((size_t)y0 * W + x0) * C);
    } else {
      int y1 = (oy & 1) ? min(y0 + 1, H - 1) : y0;
      int x1 = (ox & 1) ? min(x0 + 1, W - 1) : x0;
      v = 0.25f * ((ldf(img + ((size_t)y0 * W + x0) * C) + ldf(img + ((size_t)y0 * W + x1) * C)) +
                   (ldf(img + ((size_t)y1 * W + x0) * C) + ldf(img + ((size_t)y1 * W + x1) * C)));
    }
    stf(y + i, v);
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int B, int C, int H,
                                    int W) {
  const long long n = (long long)B * C * H * W;
  const long long hw = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {  // i indexes dst (NHWC)
    int c = (int)(i % C);
    long long t = i / C;
    long long p = t % hw;
    long long b = t / hw;
    stf(dst + i, src[(b * C + c) * hw + p]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int B, int C, int H,
                                    int W) {
  const long long n = (long long)B * C * H * W;
  const long long hw = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {  // i indexes dst (NCHW)
    long long p = i % hw;
    long long t = i / hw;
    int c = (int)(t % C);
    long long b = t / C;
    dst[i] = ldf(src + (b * hw + p) * C + c);
  }
}

// ReshapeLayer((-1,C,H,W)) of a flat NCHW-ordered feature vector <-> NHWC buffer (same dtype).
template <typename T>
__global__ void permute_kernel(const T* __restrict__ src, T* __restrict__ dst, int B, int C, int H, int W,
                               int inverse) {
  const long long n = (long long)B * C * H * W;
  const long long hw = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {  // i indexes the NHWC side
    int c = (int)(i % C);
    long long t = i / C;
    long long p = t % hw;
    long long b = t / hw;
    long long j = (b * C + c) * hw + p;           // NCHW side
    if (inverse) dst[j] = src[i];
    else dst[i] = src[j];
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    stf(d + i, ldf(s + i));
}


// ===========================================================================
// 8-wide (16-byte) variants for channel counts that are multiples of 8: every
// thread owns 8 consecutive channels of one pixel, so a warp moves 512 B per
// access and the per-element index arithmetic is paid once per 8 elements.
// ===========================================================================
template <typename T, int MODE>  // MODE 0: sum,sumsq ; 1: sum only
__global__ void __launch_bounds__(256) col_reduce_v8_kernel(const T* __restrict__ x, long long M, int C,
                                                            double* sums, float* fsum) {
  __shared__ float sh0[256 * 8], sh1[MODE == 0 ? 256 * 8 : 8];
  const int cg = C >> 3;                       // channel groups
  const int ct = cg < 256 ? cg : 256;          // groups covered per pass
  const int lanes = 256 / ct;
  const int tc = threadIdx.x % ct, tr = threadIdx.x / ct;
  const bool active = tr < lanes;
  for (int g0 = 0; g0 < cg; g0 += ct) {
    const int g = g0 + tc;
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s0[j] = s1[j] = 0.f;
    if (active && g < cg) {
      for (long long m = (long long)blockIdx.x * lanes + tr; m < M; m += (long long)gridDim.x * lanes) {
        float v[8];
        load8(x + (size_t)m * C + g * 8, v);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s0[j] += v[j];
          if (MODE == 0) s1[j] += v[j] * v[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      sh0[threadIdx.x * 8 + j] = s0[j];
      if (MODE == 0) sh1[threadIdx.x * 8 + j] = s1[j];
    }
    __syncthreads();
    if (tr == 0 && g < cg) {
      for (int l = 1; l < lanes; l++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s0[j] += sh0[(l * ct + tc) * 8 + j];
          if (MODE == 0) s1[j] += sh1[(l * ct + tc) * 8 + j];
        }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (MODE == 0) {
          atomicAdd(sums + g * 8 + j, (double)s0[j]);
          atomicAdd(sums + C + g * 8 + j, (double)s1[j]);
        } else {
          atomicAdd(fsum + g * 8 + j, s0[j]);
        }
      }
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void bn_apply_act_v8_kernel(const T* __restrict__ x, T* __restrict__ a, long long n8, int C,
                                       const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                       float slope) {
  const int cg = C >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cg) * 8;
    float v[8];
    load8(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = act_fwd(v[j] * scale[c0 + j] + shift[c0 + j], act, slope);
    store8(a + i * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    bn_bwd_reduce_v8_kernel(const T* __restrict__ da, const T* __restrict__ a, const T* __restrict__ x, long long M,
                            int C, const float* __restrict__ mean, const float* __restrict__ inv_std, int act,
                            float slope, double* red) {
  __shared__ float sh0[256 * 8], sh1[256 * 8];
  const int cg = C >> 3;
  const int ct = cg < 256 ? cg : 256;
  const int lanes = 256 / ct;
  const int tc = threadIdx.x % ct, tr = threadIdx.x / ct;
  const bool active = tr < lanes;
  for (int g0 = 0; g0 < cg; g0 += ct) {
    const int g = g0 + tc;
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s0[j] = s1[j] = 0.f;
    if (active && g < cg) {
      float mu[8], is[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        mu[j] = mean[g * 8 + j];
        is[j] = inv_std[g * 8 + j];
      }
      for (long long m = (long long)blockIdx.x * lanes + tr; m < M; m += (long long)gridDim.x * lanes) {
        const size_t o = (size_t)m * C + g * 8;
        float gv[8], av[8], xv[8];
        load8(da + o, gv);
        load8(a + o, av);
        load8(x + o, xv);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float gg = gv[j] * act_grad_from_out(av[j], act, slope);
          s0[j] += gg;
          s1[j] += gg * (xv[j] - mu[j]) * is[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      sh0[threadIdx.x * 8 + j] = s0[j];
      sh1[threadIdx.x * 8 + j] = s1[j];
    }
    __syncthreads();
    if (tr == 0 && g < cg) {
      for (int l = 1; l < lanes; l++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s0[j] += sh0[(l * ct + tc) * 8 + j];
          s1[j] += sh1[(l * ct + tc) * 8 + j];
        }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        atomicAdd(red + g * 8 + j, (double)s0[j]);
        atomicAdd(red + C + g * 8 + j, (double)s1[j]);
      }
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void bn_bwd_apply_v8_kernel(const T* __restrict__ da, const T* __restrict__ a, const T* __restrict__ x,
                                       T* __restrict__ dx, long long M, int C, const float* __restrict__ mean,
                                       const float* __restrict__ inv_std, const float* __restrict__ gamma, int act,
                                       float slope, const double* __restrict__ red, float* dgamma, float* dbeta) {
  const int cg = C >> 3;
  const long long n8 = M * cg;
  const float invM = 1.f / (float)M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cg) * 8;
    float gv[8], av[8], xv[8], o[8];
    load8(da + i * 8, gv);
    load8(a + i * 8, av);
    load8(x + i * 8, xv);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int c = c0 + j;
      const float r0 = (float)red[c], r1 = (float)red[C + c], is = inv_std[c];
      const float gg = gv[j] * act_grad_from_out(av[j], act, slope);
      const float xh = (xv[j] - mean[c]) * is;
      o[j] = gamma[c] * is * (gg - r0 * invM - xh * r1 * invM);
      if (i < cg) {
        if (dgamma) dgamma[c] = r1;
        if (dbeta) dbeta[c] = r0;
      }
    }
    store8(dx + i * 8, o);
  }
}

// ---------------------------------------------------------------------------
// Variants with more memory-level parallelism / less per-iteration work (the default; written after the round-1
// profiles showed the BatchNorm passes at ~45 % of the HBM peak).
//   * reductions: every thread keeps UNR rows in flight (independent 16-byte loads issued before any use);
//   * apply: when the grid stride is a multiple of the channel-group count, a thread sees the SAME 8 channels in every
//     iteration, so the per-channel constants are loaded once instead of 40 scalar loads per 64 bytes of traffic.
// ---------------------------------------------------------------------------
constexpr int EW_UNR = 4;

template <typename T, int MODE>  // MODE 0: sum,sumsq ; 1: sum only
__global__ void __launch_bounds__(256) col_reduce_v8u_kernel(const T* __restrict__ x, long long M, int C,
                                                             double* sums, float* fsum) {
  __shared__ float sh0[256 * 8], sh1[MODE == 0 ? 256 * 8 : 8];
  const int cg = C >> 3;
  const int ct = cg < 256 ? cg : 256;
  const int lanes = 256 / ct;
  const int tc = threadIdx.x % ct, tr = threadIdx.x / ct;
  const bool active = tr < lanes;
  const long long step = (long long)gridDim.x * lanes;
  for (int g0 = 0; g0 < cg; g0 += ct) {
    const int g = g0 + tc;
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s0[j] = s1[j] = 0.f;
    if (active && g < cg) {
      const T* px = x + g * 8;
      long long m = (long long)blockIdx.x * lanes + tr;
      for (; m + (EW_UNR - 1) * step < M; m += EW_UNR * step) {
        float v[EW_UNR][8];
#pragma unroll
        for (int u = 0; u < EW_UNR; u++) load8(px + (size_t)(m + u * step) * C, v[u]);
#pragma unroll
        for (int u = 0; u < EW_UNR; u++)
#pragma unroll
          for (int j = 0; j < 8; j++) {
            s0[j] += v[u][j];
            if (MODE == 0) s1[j] += v[u][j] * v[u][j];
          }
      }
      for (; m < M; m += step) {
        float v[8];
        load8(px + (size_t)m * C, v);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s0[j] += v[j];
          if (MODE == 0) s1[j] += v[j] * v[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      sh0[threadIdx.x * 8 + j] = s0[j];
      if (MODE == 0) sh1[threadIdx.x * 8 + j] = s1[j];
    }
    __syncthreads();
    if (tr == 0 && g < cg) {
      for (int l = 1; l < lanes; l++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s0[j] += sh0[(l * ct + tc) * 8 + j];
          if (MODE == 0) s1[j] += sh1[(l * ct + tc) * 8 + j];
        }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (MODE == 0) {
          atomicAdd(sums + g * 8 + j, (double)s0[j]);
          atomicAdd(sums + C + g * 8 + j, (double)s1[j]);
        } else {
          atomicAdd(fsum + g * 8 + j, s0[j]);
        }
      }
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    bn_bwd_reduce_v8u_kernel(const T* __restrict__ da, const T* __restrict__ a, const T* __restrict__ x, long long M,
                             int C, const float* __restrict__ mean, const float* __restrict__ inv_std, int act,
                             float slope, double* red) {
  __shared__ float sh0[256 * 8], sh1[256 * 8];
  const int cg = C >> 3;
  const int ct = cg < 256 ? cg : 256;
  const int lanes = 256 / ct;
  const int tc = threadIdx.x % ct, tr = threadIdx.x / ct;
  const bool active = tr < lanes;
  const long long step = (long long)gridDim.x * lanes;
  for (int g0 = 0; g0 < cg; g0 += ct) {
    const int g = g0 + tc;
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s0[j] = s1[j] = 0.f;
    if (active && g < cg) {
      float mu[8], is[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        mu[j] = mean[g * 8 + j];
        is[j] = inv_std[g * 8 + j];
      }
      long long m = (long long)blockIdx.x * lanes + tr;
      for (; m + step < M; m += 2 * step) {                 // two rows in flight: six independent 16-byte loads
        const size_t o0 = (size_t)m * C + g * 8, o1 = (size_t)(m + step) * C + g * 8;
        float g0v[8], a0v[8], x0v[8], g1v[8], a1v[8], x1v[8];
        load8(da + o0, g0v);
        load8(a + o0, a0v);
        load8(x + o0, x0v);
        load8(da + o1, g1v);
        load8(a + o1, a1v);
        load8(x + o1, x1v);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float gg0 = g0v[j] * act_grad_from_out(a0v[j], act, slope);
          const float gg1 = g1v[j] * act_grad_from_out(a1v[j], act, slope);
          s0[j] += gg0;
          s1[j] += gg0 * (x0v[j] - mu[j]) * is[j];
          s0[j] += gg1;
          s1[j] += gg1 * (x1v[j] - mu[j]) * is[j];
        }
      }
      for (; m < M; m += step) {
        const size_t o = (size_t)m * C + g * 8;
        float gv[8], av[8], xv[8];
        load8(da + o, gv);
        load8(a + o, av);
        load8(x + o, xv);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float gg = gv[j] * act_grad_from_out(av[j], act, slope);
          s0[j] += gg;
          s1[j] += gg * (xv[j] - mu[j]) * is[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      sh0[threadIdx.x * 8 + j] = s0[j];
      sh1[threadIdx.x * 8 + j] = s1[j];
    }
    __syncthreads();
    if (tr == 0 && g < cg) {
      for (int l = 1; l < lanes; l++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s0[j] += sh0[(l * ct + tc) * 8 + j];
          s1[j] += sh1[(l * ct + tc) * 8 + j];
        }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        atomicAdd(red + g * 8 + j, (double)s0[j]);
        atomicAdd(red + C + g * 8 + j, (double)s1[j]);
      }
    }
    __syncthreads();
  }
}

// requires (gridDim.x * blockDim.x) % (C/8) == 0 (checked by the host wrapper): the channel group of a thread is then
// the same in every grid-stride iteration
template <typename T>
__global__ void __launch_bounds__(256)
    bn_bwd_apply_v8h_kernel(const T* __restrict__ da, const T* __restrict__ a, const T* __restrict__ x,
                            T* __restrict__ dx, long long M, int C, const float* __restrict__ mean,
                            const float* __restrict__ inv_std, const float* __restrict__ gamma, int act, float slope,
                            const double* __restrict__ red, float* dgamma, float* dbeta) {
  const int cg = C >> 3;
  const long long n8 = M * cg;
  const float invM = 1.f / (float)M;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= n8) return;
  const int c0 = (int)(i0 % cg) * 8;
  float mu[8], is[8], gi[8], r0m[8], r1m[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = c0 + j;
    const float r0 = (float)red[c], r1 = (float)red[C + c];
    mu[j] = mean[c];
    is[j] = inv_std[c];
    gi[j] = gamma[c] * is[j];
    r0m[j] = r0 * invM;
    r1m[j] = r1 * invM;
    if (i0 < cg) {
      if (dgamma) dgamma[c] = r1;
      if (dbeta) dbeta[c] = r0;
    }
  }
  for (long long i = i0; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float gv[8], av[8], xv[8], o[8];
    load8(da + i * 8, gv);
    load8(a + i * 8, av);
    load8(x + i * 8, xv);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float gg = gv[j] * act_grad_from_out(av[j], act, slope);
      const float xh = (xv[j] - mu[j]) * is[j];
      o[j] = gi[j] * (gg - r0m[j] - xh * r1m[j]);
    }
    store8(dx + i * 8, o);
  }
}

// ---------------------------------------------------------------------------
// BatchNorm backward WITHOUT reading the layer's input x: xhat is recovered from the OUTPUT a = act(gamma*xhat + beta),
//     xhat = (act^-1(a) - beta) / gamma,        act^-1(a) = a >= 0 ? a : a / slope   (linear: a)
// for invertible activations (linear, leaky rectify with slope > 0) and channels with |gamma| >= 2^-10; a thread whose 8
// channels include a smaller gamma reads x for them as before.  One tensor less per pass: reduce 3 -> 2 reads,
// apply 4 -> 3 accesses (fast mode only; the float32 modes keep the x-based form, which is the oracle's arithmetic).
// ---------------------------------------------------------------------------
struct BnInv8 {
  float mu[8], is[8], ig[8], bt[8];
  bool all_ok;
  unsigned ok;
};
__device__ __forceinline__ void bn_inv8_load(BnInv8& k, const float* mean, const float* inv_std, const float* gamma,
                                             const float* beta, int c0) {
  k.ok = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float g = gamma[c0 + j];
    k.mu[j] = mean[c0 + j];
    k.is[j] = inv_std[c0 + j];
    k.bt[j] = beta[c0 + j];
    const bool ok = fabsf(g) >= 9.765625e-4f;
    k.ig[j] = ok ? 1.f / g : 0.f;
    k.ok |= ok ? (1u << j) : 0u;
  }
  k.all_ok = k.ok == 0xffu;
}
__device__ __forceinline__ float bn_xhat_from_a(const BnInv8& k, int j, float a, float x, float inv_slope) {
  if ((k.ok >> j) & 1u) return ((a >= 0.f ? a : a * inv_slope) - k.bt[j]) * k.ig[j];
  return (x - k.mu[j]) * k.is[j];
}

template <typename T>
__global__ void __launch_bounds__(256)
    bn_bwd_reduce_a_kernel(const T* __restrict__ da, const T* __restrict__ a, const T* __restrict__ x, long long M, int C,
                           const float* __restrict__ mean, const float* __restrict__ inv_std,
                           const float* __restrict__ gamma, const float* __restrict__ beta, int act, float slope,
                           double* red) {
  __shared__ float sh0[256 * 8], sh1[256 * 8];
  const int cg = C >> 3;
  const int ct = cg < 256 ? cg : 256;
  const int lanes = 256 / ct;
  const int tc = threadIdx.x % ct, tr = threadIdx.x / ct;
  const bool active = tr < lanes;
  const long long step = (long long)gridDim.x * lanes;
  const float inv_slope = act == HM_ACT_LRELU ? 1.f / slope : 1.f;
  for (int g0 = 0; g0 < cg; g0 += ct) {
    const int g = g0 + tc;
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s0[j] = s1[j] = 0.f;
    if (active && g < cg) {
      BnInv8 k;
      bn_inv8_load(k, mean, inv_std, gamma, beta, g * 8);
      long long m = (long long)blockIdx.x * lanes + tr;
      for (; m + step < M; m += 2 * step) {                 // two rows in flight: four independent 16-byte loads
        const size_t o0 = (size_t)m * C + g * 8, o1 = (size_t)(m + step) * C + g * 8;
        float g0v[8], a0v[8], x0v[8], g1v[8], a1v[8], x1v[8];
        load8(da + o0, g0v);
        load8(a + o0, a0v);
        load8(da + o1, g1v);
        load8(a + o1, a1v);
        if (!k.all_ok) {
          load8(x + o0, x0v);
          load8(x + o1, x1v);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float gg0 = g0v[j] * act_grad_from_out(a0v[j], act, slope);
          const float gg1 = g1v[j] * act_grad_from_out(a1v[j], act, slope);
          s0[j] += gg0 + gg1;
          s1[j] += gg0 * bn_xhat_from_a(k, j, a0v[j], k.all_ok ? 0.f : x0v[j], inv_slope) +
                   gg1 * bn_xhat_from_a(k, j, a1v[j], k.all_ok ? 0.f : x1v[j], inv_slope);
        }
      }
      for (; m < M; m += step) {
        const size_t o = (size_t)m * C + g * 8;
        float gv[8], av[8], xv[8];
        load8(da + o, gv);
        load8(a + o, av);
        if (!k.all_ok) load8(x + o, xv);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float gg = gv[j] * act_grad_from_out(av[j], act, slope);
          s0[j] += gg;
          s1[j] += gg * bn_xhat_from_a(k, j, av[j], k.all_ok ? 0.f : xv[j], inv_slope);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      sh0[threadIdx.x * 8 + j] = s0[j];
      sh1[threadIdx.x * 8 + j] = s1[j];
    }
    __syncthreads();
    if (tr == 0 && g < cg) {
      for (int l = 1; l < lanes; l++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s0[j] += sh0[(l * ct + tc) * 8 + j];
          s1[j] += sh1[(l * ct + tc) * 8 + j];
        }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        atomicAdd(red + g * 8 + j, (double)s0[j]);
        atomicAdd(red + C + g * 8 + j, (double)s1[j]);
      }
    }
    __syncthreads();
  }
}

// requires (gridDim.x * blockDim.x) % (C/8) == 0, as bn_bwd_apply_v8h_kernel
template <typename T>
__global__ void __launch_bounds__(256)
    bn_bwd_apply_a_kernel(const T* __restrict__ da, const T* __restrict__ a, const T* __restrict__ x, T* __restrict__ dx,
                          long long M, int C, const float* __restrict__ mean, const float* __restrict__ inv_std,
                          const float* __restrict__ gamma, const float* __restrict__ beta, int act, float slope,
                          const double* __restrict__ red, float* dgamma, float* dbeta) {
  const int cg = C >> 3;
  const long long n8 = M * cg;
  const float invM = 1.f / (float)M;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= n8) return;
  const int c0 = (int)(i0 % cg) * 8;
  const float inv_slope = act == HM_ACT_LRELU ? 1.f / slope : 1.f;
  BnInv8 k;
  bn_inv8_load(k, mean, inv_std, gamma, beta, c0);
  float gi[8], r0m[8], r1m[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = c0 + j;
    const float r0 = (float)red[c], r1 = (float)red[C + c];
    gi[j] = gamma[c] * k.is[j];
    r0m[j] = r0 * invM;
    r1m[j] = r1 * invM;
    if (i0 < cg) {
      if (dgamma) dgamma[c] = r1;
      if (dbeta) dbeta[c] = r0;
    }
  }
  for (long long i = i0; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float gv[8], av[8], xv[8], o[8];
    load8(da + i * 8, gv);
    load8(a + i * 8, av);
    if (!k.all_ok) load8(x + i * 8, xv);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float gg = gv[j] * act_grad_from_out(av[j], act, slope);
      const float xh = bn_xhat_from_a(k, j, av[j], k.all_ok ? 0.f : xv[j], inv_slope);
      o[j] = gi[j] * (gg - r0m[j] - xh * r1m[j]);
    }
    store8(dx + i * 8, o);
  }
}

// same idea for the forward apply: scale/shift of a thread's 8 channels are loop invariants
template <typename T>
__global__ void __launch_bounds__(256)
    bn_apply_act_v8h_kernel(const T* __restrict__ x, T* __restrict__ a, long long n8, int C,
                            const float* __restrict__ scale, const float* __restrict__ shift, int act, float slope) {
  const int cg = C >> 3;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= n8) return;
  const int c0 = (int)(i0 % cg) * 8;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = scale[c0 + j];
    sf[j] = shift[c0 + j];
  }
  for (long long i = i0; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    load8(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = act_fwd(v[j] * sc[j] + sf[j], act, slope);
    store8(a + i * 8, v);
  }
}

template <typename T>
__global__ void act_bwd_v8_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* dx, long long n8, int act,
                                  float slope, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float g[8], yv[8], o[8];
    load8(dy + i * 8, g);
    load8(y + i * 8, yv);
    if (accumulate) load8(dx + i * 8, o);
#pragma unroll
    for (int j = 0; j < 8; j++) o[j] = g[j] * act_grad_from_out(yv[j], act, slope) + (accumulate ? o[j] : 0.f);
    store8(dx + i * 8, o);
  }
}

template <typename T>
__global__ void maxpool2_fwd_v8_kernel(const T* __restrict__ x, T* __restrict__ p, uint8_t* __restrict__ idx, int B,
                                       int H, int W, int C) {
  const int Hp = H >> 1, Wp = W >> 1, cg = C >> 3;
  const long long n8 = (long long)B * Hp * Wp * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long t = i / cg;
    const int px = (int)(t % Wp);
    t /= Wp;
    const int py = (int)(t % Hp);
    const int b = (int)(t / Hp);
    const T* base = x + (((size_t)b * H + 2 * py) * W + 2 * px) * C + g * 8;
    float v0[8], v1[8], v2[8], v3[8], m[8];
    load8(base, v0);
    load8(base + C, v1);
    load8(base + (size_t)W * C, v2);
    load8(base + (size_t)W * C + C, v3);
    uint32_t klo = 0, khi = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float mm = v0[j];
      uint32_t k = 0;
      if (v1[j] > mm) { mm = v1[j]; k = 1; }
      if (v2[j] > mm) { mm = v2[j]; k = 2; }
      if (v3[j] > mm) { mm = v3[j]; k = 3; }
      m[j] = mm;
      if (j < 4) klo |= k << (8 * j); else khi |= k << (8 * (j - 4));
    }
    store8(p + i * 8, m);
    *reinterpret_cast<uint2*>(idx + i * 8) = make_uint2(klo, khi);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    maxpool2_bwd_v8_kernel(const T* __restrict__ dp, const T* __restrict__ p, const uint8_t* __restrict__ idx,
                           T* __restrict__ dx, int B, int H, int W, int C, int act, float slope, float* db,
                           T* __restrict__ dxs, const float* __restrict__ scale) {
  // dxs/scale (optional): second copy dxs = scale[image] * dx, and db then sums the SCALED gradient (per-sample
  // weighting of the weight/bias gradients, hm_maxpool2_bwd_scaled).
  // db (optional): per-channel sum of the scattered gradient = bias gradient of the convolution that fed the pool.
  // Needs 256 % (C/8) == 0 so that a thread keeps its channel group over the grid-stride loop.
  __shared__ float shb[256 * 8];
  float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int Hp = H >> 1, Wp = W >> 1, cg = C >> 3;
  const long long n8 = (long long)B * Hp * Wp * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long t = i / cg;
    const int px = (int)(t % Wp);
    t /= Wp;
    const int py = (int)(t % Hp);
    const int b = (int)(t / Hp);
    float gv[8], pv[8];
    load8(dp + i * 8, gv);
    load8(p + i * 8, pv);
    const uint2 kk = *reinterpret_cast<const uint2*>(idx + i * 8);
    const float sc = scale ? scale[b] : 1.f;
    float o0[8], o1[8], o2[8], o3[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float gg = gv[j] * act_grad_from_out(pv[j], act, slope);
      const uint32_t k = ((j < 4 ? kk.x : kk.y) >> (8 * (j & 3))) & 0xff;
      bs[j] += gg * sc;
      o0[j] = k == 0 ? gg : 0.f;
      o1[j] = k == 1 ? gg : 0.f;
      o2[j] = k == 2 ? gg : 0.f;
      o3[j] = k == 3 ? gg : 0.f;
    }
    const size_t off = (((size_t)b * H + 2 * py) * W + 2 * px) * C + g * 8;
    if (dx) {                                   // (null: only the weighted copy / the bias gradient are wanted)
      T* base = dx + off;
      store8(base, o0);
      store8(base + C, o1);
      store8(base + (size_t)W * C, o2);
      store8(base + (size_t)W * C + C, o3);
    }
    if (dxs) {
#pragma unroll
      for (int j = 0; j < 8; j++) { o0[j] *= sc; o1[j] *= sc; o2[j] *= sc; o3[j] *= sc; }
      T* bs2 = dxs + off;
      store8(bs2, o0);
      store8(bs2 + C, o1);
      store8(bs2 + (size_t)W * C, o2);
      store8(bs2 + (size_t)W * C + C, o3);
    }
  }
  if (db) {
#pragma unroll
    for (int j = 0; j < 8; j++) shb[threadIdx.x * 8 + j] = bs[j];
    __syncthreads();
    if ((int)threadIdx.x < cg) {
      for (int l = threadIdx.x + cg; l < 256; l += cg)
#pragma unroll
        for (int j = 0; j < 8; j++) bs[j] += shb[l * 8 + j];
#pragma unroll
      for (int j = 0; j < 8; j++) atomicAdd(db + threadIdx.x * 8 + j, bs[j]);
    }
  }
}

template <typename T>
__global__ void upsample2_fwd_v8_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C,
                                        int mode) {
  const int H2 = 2 * H, W2 = 2 * W, cg = C >> 3;
  const long long n8 = (long long)B * H2 * W2 * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long t = i / cg;
    const int ox = (int)(t % W2);
    t /= W2;
    const int oy = (int)(t % H2);
    const int b = (int)(t / H2);
    const T* img = x + (size_t)b * H * W * C + g * 8;
    const int y0 = oy >> 1, x0 = ox >> 1;
    float v[8];
    if (mode == HM_UP_NEAREST2) {
      load8(img + ((size_t)y0 * W + x0) * C, v);
    } else {
      const int y1 = (oy & 1) ? min(y0 + 1, H - 1) : y0;
      const int x1 = (ox & 1) ? min(x0 + 1, W - 1) : x0;
      float a[8], bb[8], e[8], f[8];
      load8(img + ((size_t)y0 * W + x0) * C, a);
      load8(img + ((size_t)y0 * W + x1) * C, bb);
      load8(img + ((size_t)y1 * W + x0) * C, e);
      load8(img + ((size_t)y1 * W + x1) * C, f);
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = 0.25f * ((a[j] + bb[j]) + (e[j] + f[j]));
    }
    store8(y + i * 8, v);
  }
}

template <typename T>
__global__ void upsample2_bwd_v8_kernel(const T* __restrict__ dy, T* dx, int B, int H, int W, int C, int mode,
                                        int accumulate) {
  const int cg = C >> 3;
  const long long n8 = (long long)B * H * W * cg;
  const int H2 = 2 * H, W2 = 2 * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long t = i / cg;
    const int x = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    const T* img = dy + (size_t)b * H2 * W2 * C + g * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    if (mode == HM_UP_NEAREST2) {
      for (int a = 0; a < 2; a++)
        for (int e = 0; e < 2; e++) {
          float v[8];
          load8(img + ((size_t)(2 * y + a) * W2 + (2 * x + e)) * C, v);
#pragma unroll
          for (int j = 0; j < 8; j++) acc[j] += v[j];
        }
    } else {
      int ys[3], xs[3];
      float wy[3], wx[3];
      int ny = 0, nx = 0;
      ys[ny] = 2 * y; wy[ny++] = 1.f;
      ys[ny] = 2 * y + 1; wy[ny++] = (y == H - 1) ? 1.f : .5f;
      if (y > 0) { ys[ny] = 2 * y - 1; wy[ny++] = .5f; }
      xs[nx] = 2 * x; wx[nx++] = 1.f;
      xs[nx] = 2 * x + 1; wx[nx++] = (x == W - 1) ? 1.f : .5f;
      if (x > 0) { xs[nx] = 2 * x - 1; wx[nx++] = .5f; }
      for (int a = 0; a < ny; a++)
        for (int e = 0; e < nx; e++) {
          float v[8];
          load8(img + ((size_t)ys[a] * W2 + xs[e]) * C, v);
          const float wgt = wy[a] * wx[e];
#pragma unroll
          for (int j = 0; j < 8; j++) acc[j] += wgt * v[j];
        }
    }
    if (accumulate) {
      float o[8];
      load8(dx + i * 8, o);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] += o[j];
    }
    store8(dx + i * 8, acc);
  }
}

template <typename T>
__global__ void copy_v8_kernel(const T* __restrict__ s, T* __restrict__ d, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x)
#pragma unroll
    for (int q = 0; q < (int)(sizeof(T) / 2); q++)     // 8 elements = 1 (fp16) or 2 (fp32) 16-byte words
      reinterpret_cast<uint4*>(d)[i * (sizeof(T) / 2) + q] = reinterpret_cast<const uint4*>(s)[i * (sizeof(T) / 2) + q];
}

static inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// dst[m][0..nc) (+)= src[m][c0..c0+nc)  (channel slice of an NHWC tensor; concat gradients)
template <typename T>
__global__ void slice_channels_kernel(const T* __restrict__ src, T* dst, long long M, int C, int c0, int nc,
                                      int accumulate) {
  const long long n = M * nc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nc;
    const int c = (int)(i - m * nc);
    float v = ldf(src + m * C + c0 + c);
    if (accumulate) v += ldf(dst + i);
    stf(dst + i, v);
  }
}

// ---- losses ---------------------------------------------------------------
template <typename T>
__global__ void adv_loss_kernel(const T* __restrict__ h, T* dh, long long R, int G, int out_act, float target,
                                int lsgan, int relu_head, float weight, float gscale, int accumulate,
                                float* loss) {
  float local = 0.f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R;
       r += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < G; g++) s += ldf(h + r * G + g);
    float out = act_fwd(s / (float)G, out_act, 0.f);
    float l, dl;
    if (lsgan) {
      float e = out - target;
      l = e * e;
      dl = 2.f * e;
    } else {  // binary_crossentropy(out, target)
      l = -(target * logf(out) + (1.f - target) * logf(1.f - out));
      dl = -(target / out) + (1.f - target) / (1.f - out);
    }
    local += l;
    if (dh) {
      dl *= act_grad_from_out(out, out_act, 0.f);
      float gbase = gscale * weight * dl / ((float)R * (float)G);
      for (int g = 0; g < G; g++) {
        float v = gbase;
        if (relu_head && !(ldf(h + r * G + g) > 0.f)) v = 0.f;
        if (accumulate) v += ldf(dh + r * G + g);
        stf(dh + r * G + g, v);
      }
    }
  }
  local = warp_sum(local);
  __shared__ float sh[32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = local;
  __syncthreads();
  if (w == 0) {
    float v = l < (blockDim.x >> 5) ? sh[l] : 0.f;
    v = warp_sum(v);
    if (l == 0) atomicAdd(loss, weight * v / (float)R);
  }
}

// Fake half of a scalar-output discriminator, both of its losses at once (see hm_adv_loss_pair in hmgan.h).
template <typename T>
__global__ void adv_loss_pair_kernel(const T* __restrict__ h, T* dh, T* dhw, float* sw, float* sg, long long R, int G,
                                     int out_act, int lsgan, int relu_head, float gscale, float* loss_disc,
                                     float* loss_gen) {
  float l0 = 0.f, l1 = 0.f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R;
       r += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < G; g++) s += ldf(h + r * G + g);
    const float out = act_fwd(s / (float)G, out_act, 0.f);
    float la, lb, da, db_;
    if (lsgan) {
      la = out * out;                  da = 2.f * out;                 // target 0 (discriminator loss on a fake)
      lb = (out - 1.f) * (out - 1.f);  db_ = 2.f * (out - 1.f);        // target 1 (generator loss)
    } else {
      la = -logf(1.f - out);           da = 1.f / (1.f - out);
      lb = -logf(out);                 db_ = -1.f / out;
    }
    l0 += la;
    l1 += lb;
    const float k = gscale * act_grad_from_out(out, out_act, 0.f) / ((float)R * (float)G);
    const float a = da * k, b = db_ * k;
    const float c = fabsf(a) >= fabsf(b) ? a : b;
    sw[r] = c != 0.f ? a / c : 0.f;
    sg[r] = c != 0.f ? b / c : 0.f;
    for (int g = 0; g < G; g++) {
      const bool on = !relu_head || ldf(h + r * G + g) > 0.f;
      stf(dh + r * G + g, on ? c : 0.f);
      stf(dhw + r * G + g, on ? a : 0.f);
    }
  }
  l0 = warp_sum(l0);
  l1 = warp_sum(l1);
  __shared__ float sh0[32], sh1[32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh0[w] = l0; sh1[w] = l1; }
  __syncthreads();
  if (w == 0) {
    float v0 = l < (blockDim.x >> 5) ? sh0[l] : 0.f, v1 = l < (blockDim.x >> 5) ? sh1[l] : 0.f;
    v0 = warp_sum(v0);
    v1 = warp_sum(v1);
    if (l == 0) {
      atomicAdd(loss_disc, v0 / (float)R);
      atomicAdd(loss_gen, v1 / (float)R);
    }
  }
}

template <typename T>
__global__ void scale_rows_kernel(const T* __restrict__ src, const float* __restrict__ scale, T* __restrict__ dst,
                                  long long R, long long L) {
  const long long n = R * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    stf(dst + i, ldf(src + i) * scale[i / L]);
}

template <typename T>
__global__ void recon_loss_kernel(const T* __restrict__ p, const T* __restrict__ y, T* dp, long long n, int l2,
                                  float weight, float gscale, int accumulate, float* loss) {
  float local = 0.f;
  const float gn = gscale * weight / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float e = ldf(p + i) - ldf(y + i);
    float g;
    if (l2) {
      local += e * e;
      g = 2.f * e * gn;
    } else {
      local += fabsf(e);
      g = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * gn;
    }
    if (dp) {
      if (accumulate) g += ldf(dp + i);
      stf(dp + i, g);
    }
  }
  local = warp_sum(local);
  __shared__ float sh[32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = local;
  __syncthreads();
  if (w == 0) {
    float v = l < (blockDim.x >> 5) ? sh[l] : 0.f;
    v = warp_sum(v);
    if (l == 0) atomicAdd(loss, weight * v / (float)n);
  }
}

// ---- optimisers -----------------------------------------------------------
__global__ void rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ acc,
                               long long n, const float* __restrict__ lr_p, float rho, float eps, float gscale) {
  const float lr = *lr_p;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float a = rho * acc[i] + (1.f - rho) * gi * gi;
    acc[i] = a;
    p[i] = p[i] - lr * gi / sqrtf(a + eps);
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, const float* __restrict__ lr_p, float b1, float b2,
                            float eps, float corr, float gscale) {
  const float a_t = *lr_p * corr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - a_t * mi / (sqrtf(vi) + eps);
  }
}

// adam with the step count in DEVICE memory (so that a captured CUDA graph advances it on every replay)
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, const float* __restrict__ lr_p, float b1, float b2,
                                float eps, const int* __restrict__ t_p, float gscale) {
  const float t = (float)*t_p;
  const float a_t = *lr_p * (sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t)));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - a_t * mi / (sqrtf(vi) + eps);
  }
}
__global__ void inc_i32_kernel(int* t) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *t += 1;
}

}  // namespace hm

using namespace hm;

#define DISPATCH_T(dtype, ...)                 \
  if ((dtype) == HM_F32) {                     \
    using T = float;                           \
    __VA_ARGS__;                               \
  } else {                                     \
    using T = __half;                          \
    __VA_ARGS__;                               \
  }

#define CHECK_DTYPE(dtype, who) \
  HM_CHECK_ARG((dtype) == HM_F32 || (dtype) == HM_F16, "%s: bad dtype %d", who, (int)(dtype))

extern "C" int hm_version(void) { return 100; }
extern "C" const char* hm_last_error_string(void) { return g_err; }
extern "C" int hm_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int hm_bn_stats(const void* x, int dtype, long long M, int C, double* sums, void* stream) {
  CHECK_DTYPE(dtype, "hm_bn_stats");
  HM_CHECK_ARG(x && sums && M > 0 && C > 0, "hm_bn_stats: bad argument");
  if (C % 8 == 0 && al16(x)) {
    int cg = C / 8, l8 = 256 / (cg < 256 ? cg : 256);
    unsigned g8 = ew_grid((M + l8 - 1) / l8, 1, 4);
    if (ew_hoist()) {
      DISPATCH_T(dtype, (col_reduce_v8u_kernel<T, 0><<<g8, 256, 0, (cudaStream_t)stream>>>((const T*)x, M, C, sums,
                                                                                          nullptr)));
    } else {
      DISPATCH_T(dtype, (col_reduce_v8_kernel<T, 0><<<g8, 256, 0, (cudaStream_t)stream>>>((const T*)x, M, C, sums,
                                                                                         nullptr)));
    }
    HM_CHECK_LAUNCH("hm_bn_stats");
    return HM_OK;
  }
  int lanes = RED_THREADS / (C < RED_THREADS ? C : RED_THREADS);
  unsigned grid = ew_grid((M + lanes - 1) / lanes, 1, 4);
  DISPATCH_T(dtype, (col_reduce_kernel<T, 0><<<grid, RED_THREADS, 0, (cudaStream_t)stream>>>((const T*)x, M, C,
                                                                                             sums, nullptr)));
  HM_CHECK_LAUNCH("hm_bn_stats");
  return HM_OK;
}

extern "C" int hm_col_sum(const void* dy, int dtype, long long M, int C, float* db, void* stream) {
  CHECK_DTYPE(dtype, "hm_col_sum");
  HM_CHECK_ARG(dy && db && M > 0 && C > 0, "hm_col_sum: bad argument");
  if (C % 8 == 0 && al16(dy)) {
    int cg = C / 8, l8 = 256 / (cg < 256 ? cg : 256);
    unsigned g8 = ew_grid((M + l8 - 1) / l8, 1, 4);
    if (ew_hoist()) {
      DISPATCH_T(dtype, (col_reduce_v8u_kernel<T, 1><<<g8, 256, 0, (cudaStream_t)stream>>>((const T*)dy, M, C, nullptr,
                                                                                          db)));
    } else {
      DISPATCH_T(dtype, (col_reduce_v8_kernel<T, 1><<<g8, 256, 0, (cudaStream_t)stream>>>((const T*)dy, M, C, nullptr,
                                                                                         db)));
    }
    HM_CHECK_LAUNCH("hm_col_sum");
    return HM_OK;
  }
  int lanes = RED_THREADS / (C < RED_THREADS ? C : RED_THREADS);
  unsigned grid = ew_grid((M + lanes - 1) / lanes, 1, 4);
  DISPATCH_T(dtype, (col_reduce_kernel<T, 1><<<grid, RED_THREADS, 0, (cudaStream_t)stream>>>((const T*)dy, M, C,
                                                                                             nullptr, db)));
  HM_CHECK_LAUNCH("hm_col_sum");
  return HM_OK;
}

extern "C" int hm_bn_finalize(const double* sums, long long M, int C, const float* gamma, const float* beta,
                              float* running_mean, float* running_inv_std, float eps, float alpha,
                              int update_running, float* mean, float* inv_std, float* scale, float* shift,
                              void* stream) {
  HM_CHECK_ARG(gamma && beta && scale && shift && C > 0, "hm_bn_finalize: bad argument");
  HM_CHECK_ARG(sums || (running_mean && running_inv_std), "hm_bn_finalize: no statistics given");
  HM_CHECK_ARG(!update_running || (running_mean && running_inv_std), "hm_bn_finalize: no running buffers");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      sums, M, C, gamma, beta, running_mean, running_inv_std, eps, alpha, update_running, mean, inv_std, scale,
      shift);
  HM_CHECK_LAUNCH("hm_bn_finalize");
  return HM_OK;
}

extern "C" int hm_bn_apply_act(const void* x, void* a, int dtype, long long M, int C, const float* scale,
                               const float* shift, int act, float slope, void* stream) {
  CHECK_DTYPE(dtype, "hm_bn_apply_act");
  HM_CHECK_ARG(x && a && scale && shift && M > 0 && C > 0, "hm_bn_apply_act: bad argument");
  long long n = M * C;
  if (C % 8 == 0 && al16(x) && al16(a)) {
    const unsigned ga = ew_grid(n / 8);
    if (ew_hoist() && ((long long)ga * 256) % (C / 8) == 0) {
      DISPATCH_T(dtype, (bn_apply_act_v8h_kernel<T><<<ga, 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, (T*)a, n / 8, C, scale, shift, act, slope)));
    } else {
      DISPATCH_T(dtype, (bn_apply_act_v8_kernel<T><<<ga, 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, (T*)a, n / 8, C, scale, shift, act, slope)));
    }
    HM_CHECK_LAUNCH("hm_bn_apply_act");
    return HM_OK;
  }
  DISPATCH_T(dtype, (bn_apply_act_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, (T*)a, n, C, scale, shift, act, slope)));
  HM_CHECK_LAUNCH("hm_bn_apply_act");
  return HM_OK;
}

extern "C" int hm_bn_bwd_reduce(const void* da, const void* a, const void* x, int dtype, long long M, int C,
                                const float* mean, const float* inv_std, int act, float slope, double* red,
                                void* stream) {
  CHECK_DTYPE(dtype, "hm_bn_bwd_reduce");
  HM_CHECK_ARG(da && a && x && mean && inv_std && red && M > 0 && C > 0, "hm_bn_bwd_reduce: bad argument");
  if (C % 8 == 0 && al16(da) && al16(a) && al16(x)) {
    int cg = C / 8, l8 = 256 / (cg < 256 ? cg : 256);
    unsigned g8 = ew_grid((M + l8 - 1) / l8, 1, 4);
    if (ew_hoist()) {
      DISPATCH_T(dtype, (bn_bwd_reduce_v8u_kernel<T><<<g8, 256, 0, (cudaStream_t)stream>>>(
                            (const T*)da, (const T*)a, (const T*)x, M, C, mean, inv_std, act, slope, red)));
    } else {
      DISPATCH_T(dtype, (bn_bwd_reduce_v8_kernel<T><<<g8, 256, 0, (cudaStream_t)stream>>>(
                            (const T*)da, (const T*)a, (const T*)x, M, C, mean, inv_std, act, slope, red)));
    }
    HM_CHECK_LAUNCH("hm_bn_bwd_reduce");
    return HM_OK;
  }
  int lanes = RED_THREADS / (C < RED_THREADS ? C : RED_THREADS);
  unsigned grid = ew_grid((M + lanes - 1) / lanes, 1, 4);
  DISPATCH_T(dtype, (bn_bwd_reduce_kernel<T><<<grid, RED_THREADS, 0, (cudaStream_t)stream>>>(
                        (const T*)da, (const T*)a, (const T*)x, M, C, mean, inv_std, act, slope, red)));
  HM_CHECK_LAUNCH("hm_bn_bwd_reduce");
  return HM_OK;
}

extern "C" int hm_bn_bwd_apply(const void* da, const void* a, const void* x, void* dx, int dtype, long long M,
                               int C, const float* mean, const float* inv_std, const float* gamma, int act,
                               float slope, const double* red, float* dgamma, float* dbeta, void* stream) {
  CHECK_DTYPE(dtype, "hm_bn_bwd_apply");
  HM_CHECK_ARG(da && a && x && dx && mean && inv_std && gamma && red && M > 0 && C > 0,
               "hm_bn_bwd_apply: bad argument");
  long long n = M * C;
  if (C % 8 == 0 && al16(da) && al16(a) && al16(x) && al16(dx)) {
    const unsigned ga = ew_grid(n / 8);
    if (ew_hoist() && ((long long)ga * 256) % (C / 8) == 0) {
      DISPATCH_T(dtype, (bn_bwd_apply_v8h_kernel<T><<<ga, 256, 0, (cudaStream_t)stream>>>(
                            (const T*)da, (const T*)a, (const T*)x, (T*)dx, M, C, mean, inv_std, gamma, act, slope,
                            red, dgamma, dbeta)));
    } else {
      DISPATCH_T(dtype, (bn_bwd_apply_v8_kernel<T><<<ga, 256, 0, (cudaStream_t)stream>>>(
                            (const T*)da, (const T*)a, (const T*)x, (T*)dx, M, C, mean, inv_std, gamma, act, slope,
                            red, dgamma, dbeta)));
    }
    HM_CHECK_LAUNCH("hm_bn_bwd_apply");
    return HM_OK;
  }
  DISPATCH_T(dtype, (bn_bwd_apply_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)da, (const T*)a, (const T*)x, (T*)dx, M, C, mean, inv_std, gamma, act, slope,
                        red, dgamma, dbeta)));
  HM_CHECK_LAUNCH("hm_bn_bwd_apply");
  return HM_OK;
}

// xhat from the layer's OUTPUT where that is possible (see bn_bwd_reduce_a_kernel): same results up to rounding, one
// tensor read less.  Falls back to hm_bn_bwd_reduce / hm_bn_bwd_apply for shapes or activations it does not cover.
static inline bool bn_from_a_ok(int act, float slope, int C) {
  return C % 8 == 0 && (act == HM_ACT_LINEAR || (act == HM_ACT_LRELU && slope > 1e-3f));
}

extern "C" int hm_bn_bwd_reduce_a(const void* da, const void* a, const void* x, int dtype, long long M, int C,
                                  const float* mean, const float* inv_std, const float* gamma, const float* beta, int act,
                                  float slope, double* red, void* stream) {
  CHECK_DTYPE(dtype, "hm_bn_bwd_reduce_a");
  HM_CHECK_ARG(da && a && x && mean && inv_std && gamma && beta && red && M > 0 && C > 0, "hm_bn_bwd_reduce_a: bad argument");
  if (!bn_from_a_ok(act, slope, C) || !al16(da) || !al16(a) || !al16(x))
    return hm_bn_bwd_reduce(da, a, x, dtype, M, C, mean, inv_std, act, slope, red, stream);
  int cg = C / 8, l8 = 256 / (cg < 256 ? cg : 256);
  unsigned g8 = ew_grid((M + l8 - 1) / l8, 1, 4);
  DISPATCH_T(dtype, (bn_bwd_reduce_a_kernel<T><<<g8, 256, 0, (cudaStream_t)stream>>>(
                        (const T*)da, (const T*)a, (const T*)x, M, C, mean, inv_std, gamma, beta, act, slope, red)));
  HM_CHECK_LAUNCH("hm_bn_bwd_reduce_a");
  return HM_OK;
}

extern "C" int hm_bn_bwd_apply_a(const void* da, const void* a, const void* x, void* dx, int dtype, long long M, int C,
                                 const float* mean, const float* inv_std, const float* gamma, const float* beta, int act,
                                 float slope, const double* red, float* dgamma, float* dbeta, void* stream) {
  CHECK_DTYPE(dtype, "hm_bn_bwd_apply_a");
  HM_CHECK_ARG(da && a && x && dx && mean && inv_std && gamma && beta && red && M > 0 && C > 0,
               "hm_bn_bwd_apply_a: bad argument");
  const long long n = M * C;
  const unsigned ga = ew_grid(n / 8 > 0 ? n / 8 : 1);
  if (!bn_from_a_ok(act, slope, C) || !al16(da) || !al16(a) || !al16(x) || !al16(dx) || ((long long)ga * 256) % (C / 8) != 0)
    return hm_bn_bwd_apply(da, a, x, dx, dtype, M, C, mean, inv_std, gamma, act, slope, red, dgamma, dbeta, stream);
  DISPATCH_T(dtype, (bn_bwd_apply_a_kernel<T><<<ga, 256, 0, (cudaStream_t)stream>>>(
                        (const T*)da, (const T*)a, (const T*)x, (T*)dx, M, C, mean, inv_std, gamma, beta, act, slope, red,
                        dgamma, dbeta)));
  HM_CHECK_LAUNCH("hm_bn_bwd_apply_a");
  return HM_OK;
}

extern "C" int hm_act_bwd(const void* dy, const void* y, void* dx, int dtype, long long n, int act, float slope,
                          int accumulate, void* stream) {
  CHECK_DTYPE(dtype, "hm_act_bwd");
  HM_CHECK_ARG(dy && y && dx && n > 0, "hm_act_bwd: bad argument");
  if (n % 8 == 0 && al16(dy) && al16(y) && al16(dx)) {
    DISPATCH_T(dtype, (act_bwd_v8_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>(
                          (const T*)dy, (const T*)y, (T*)dx, n / 8, act, slope, accumulate)));
    HM_CHECK_LAUNCH("hm_act_bwd");
    return HM_OK;
  }
  DISPATCH_T(dtype, (act_bwd_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)dy, (const T*)y, (T*)dx, n, act, slope, accumulate)));
  HM_CHECK_LAUNCH("hm_act_bwd");
  return HM_OK;
}

extern "C" int hm_maxpool2_fwd(const void* x, void* p, uint8_t* idx, int dtype, int B, int H, int W, int C,
                               void* stream) {
  CHECK_DTYPE(dtype, "hm_maxpool2_fwd");
  HM_CHECK_ARG(x && p && idx && B > 0 && H >= 2 && W >= 2 && C > 0, "hm_maxpool2_fwd: bad argument");
  long long n = (long long)B * (H / 2) * (W / 2) * C;
  if (C % 8 == 0 && al16(x) && al16(p) && (((uintptr_t)idx) & 7) == 0) {
    DISPATCH_T(dtype, (maxpool2_fwd_v8_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)p,
                                                                                                   idx, B, H, W, C)));
    HM_CHECK_LAUNCH("hm_maxpool2_fwd");
    return HM_OK;
  }
  DISPATCH_T(dtype, (maxpool2_fwd_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)p,
                                                                                          idx, B, H, W, C)));
  HM_CHECK_LAUNCH("hm_maxpool2_fwd");
  return HM_OK;
}

extern "C" int hm_maxpool2_bwd(const void* dp, const void* p, const uint8_t* idx, void* dx, int dtype, int B,
                               int H, int W, int C, int act, float slope, float* db, void* stream) {
  CHECK_DTYPE(dtype, "hm_maxpool2_bwd");
  HM_CHECK_ARG(dp && p && idx && dx && B > 0 && H >= 2 && W >= 2 && C > 0, "hm_maxpool2_bwd: bad argument");
  long long n = (long long)B * (H / 2) * (W / 2) * C;
  if (C % 8 == 0 && al16(dp) && al16(p) && al16(dx) && (((uintptr_t)idx) & 7) == 0) {
    const bool fuse_db = db && (256 % (C / 8) == 0);
    DISPATCH_T(dtype, (maxpool2_bwd_v8_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>(
                          (const T*)dp, (const T*)p, idx, (T*)dx, B, H, W, C, act, slope, fuse_db ? db : nullptr,
                          (T*)nullptr, nullptr)));
    HM_CHECK_LAUNCH("hm_maxpool2_bwd");
    if (db && !fuse_db) return hm_col_sum(dx, dtype, (long long)B * H * W, C, db, stream);
    return HM_OK;
  }
  DISPATCH_T(dtype, (maxpool2_bwd_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)dp, (const T*)p, idx, (T*)dx, B, H, W, C, act, slope)));
  HM_CHECK_LAUNCH("hm_maxpool2_bwd");
  if (db) return hm_col_sum(dx, dtype, (long long)B * H * W, C, db, stream);
  return HM_OK;
}

extern "C" int hm_upsample2_bwd(const void* dy, void* dx, int dtype, int B, int H, int W, int C, int mode,
                                int accumulate, void* stream) {
  CHECK_DTYPE(dtype, "hm_upsample2_bwd");
  HM_CHECK_ARG(dy && dx && B > 0 && H > 0 && W > 0 && C > 0, "hm_upsample2_bwd: bad argument");
  HM_CHECK_ARG(mode == HM_UP_NEAREST2 || mode == HM_UP_BILINEAR2, "hm_upsample2_bwd: bad mode %d", mode);
  long long n = (long long)B * H * W * C;
  if (C % 8 == 0 && al16(dy) && al16(dx)) {
    DISPATCH_T(dtype, (upsample2_bwd_v8_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>(
                          (const T*)dy, (T*)dx, B, H, W, C, mode, accumulate)));
    HM_CHECK_LAUNCH("hm_upsample2_bwd");
    return HM_OK;
  }
  DISPATCH_T(dtype, (upsample2_bwd_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)dy, (T*)dx, B, H, W, C, mode, accumulate)));
  HM_CHECK_LAUNCH("hm_upsample2_bwd");
  return HM_OK;
}

extern "C" int hm_upsample2_fwd(const void* x, void* y, int dtype, int B, int H, int W, int C, int mode,
                                void* stream) {
  CHECK_DTYPE(dtype, "hm_upsample2_fwd");
  HM_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0, "hm_upsample2_fwd: bad argument");
  HM_CHECK_ARG(mode == HM_UP_NEAREST2 || mode == HM_UP_BILINEAR2, "hm_upsample2_fwd: bad mode %d", mode);
  long long n = (long long)B * H * W * C * 4;
  if (C % 8 == 0 && al16(x) && al16(y)) {
    DISPATCH_T(dtype, (upsample2_fwd_v8_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y,
                                                                                                    B, H, W, C, mode)));
    HM_CHECK_LAUNCH("hm_upsample2_fwd");
    return HM_OK;
  }
  DISPATCH_T(dtype, (upsample2_fwd_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, B,
                                                                                           H, W, C, mode)));
  HM_CHECK_LAUNCH("hm_upsample2_fwd");
  return HM_OK;
}

extern "C" int hm_nchw_to_nhwc(const float* src, void* dst, int dtype, int B, int C, int H, int W,
                               void* stream) {
  CHECK_DTYPE(dtype, "hm_nchw_to_nhwc");
  HM_CHECK_ARG(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "hm_nchw_to_nhwc: bad argument");
  long long n = (long long)B * C * H * W;
  DISPATCH_T(dtype, (nchw_to_nhwc_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(src, (T*)dst, B, C, H,
                                                                                          W)));
  HM_CHECK_LAUNCH("hm_nchw_to_nhwc");
  return HM_OK;
}

extern "C" int hm_nhwc_to_nchw(const void* src, float* dst, int dtype, int B, int C, int H, int W,
                               void* stream) {
  CHECK_DTYPE(dtype, "hm_nhwc_to_nchw");
  HM_CHECK_ARG(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "hm_nhwc_to_nchw: bad argument");
  long long n = (long long)B * C * H * W;
  DISPATCH_T(dtype, (nhwc_to_nchw_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const T*)src, dst, B,
                                                                                          C, H, W)));
  HM_CHECK_LAUNCH("hm_nhwc_to_nchw");
  return HM_OK;
}

extern "C" int hm_permute(const void* src, void* dst, int dtype, int B, int C, int H, int W, int inverse,
                          void* stream) {
  CHECK_DTYPE(dtype, "hm_permute");
  HM_CHECK_ARG(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "hm_permute: bad argument");
  long long n = (long long)B * C * H * W;
  DISPATCH_T(dtype, (permute_kernel<T><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const T*)src, (T*)dst, B, C,
                                                                                     H, W, inverse)));
  HM_CHECK_LAUNCH("hm_permute");
  return HM_OK;
}

extern "C" int hm_slice_channels(const void* src, void* dst, int dtype, long long M, int C, int c0, int nc,
                                 int accumulate, void* stream) {
  CHECK_DTYPE(dtype, "hm_slice_channels");
  HM_CHECK_ARG(src && dst && M > 0 && C > 0 && c0 >= 0 && nc > 0 && c0 + nc <= C, "hm_slice_channels: bad argument");
  DISPATCH_T(dtype, (slice_channels_kernel<T><<<ew_grid(M * nc), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)src, (T*)dst, M, C, c0, nc, accumulate)));
  HM_CHECK_LAUNCH("hm_slice_channels");
  return HM_OK;
}

extern "C" int hm_cast(const void* src, int sd, void* dst, int dd, long long n, void* stream) {
  CHECK_DTYPE(sd, "hm_cast");
  CHECK_DTYPE(dd, "hm_cast");
  HM_CHECK_ARG(src && dst && n > 0, "hm_cast: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned g = ew_grid(n);
  if (sd == dd && n % 8 == 0 && al16(src) && al16(dst)) {
    if (sd == HM_F32) copy_v8_kernel<float><<<ew_grid(n / 8), 256, 0, st>>>((const float*)src, (float*)dst, n / 8);
    else copy_v8_kernel<__half><<<ew_grid(n / 8), 256, 0, st>>>((const __half*)src, (__half*)dst, n / 8);
    HM_CHECK_LAUNCH("hm_cast");
    return HM_OK;
  }
  if (sd == HM_F32 && dd == HM_F16) cast_kernel<float, __half><<<g, 256, 0, st>>>((const float*)src, (__half*)dst, n);
  else if (sd == HM_F16 && dd == HM_F32) cast_kernel<__half, float><<<g, 256, 0, st>>>((const __half*)src, (float*)dst, n);
  else if (sd == HM_F32) cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else cast_kernel<__half, __half><<<g, 256, 0, st>>>((const __half*)src, (__half*)dst, n);
  HM_CHECK_LAUNCH("hm_cast");
  return HM_OK;
}

// uint8 NHWC image data (the reference's on-disk layout) -> compute-dtype NHWC with the iterator's normalisation done in
// float32 as numpy does it (reference util.py:33-35): x/255 for grayscale images, (x-127.5)/127.5 otherwise.  IEEE
// subtraction and division, so the fp32 result equals the host iterator's bit for bit.
__device__ __forceinline__ float u8_norm(unsigned v, int tanh_range) {
  float f = (float)v;
  return tanh_range ? __fdiv_rn(__fsub_rn(f, 127.5f), 127.5f) : __fdiv_rn(f, 255.0f);
}

template <typename T>
__global__ void __launch_bounds__(256) u8_norm_v16_kernel(const uint4* __restrict__ s, T* __restrict__ d, long long n16,
                                                          int tanh_range) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(s + i);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    float f[16];
#pragma unroll
    for (int k = 0; k < 16; k++) f[k] = u8_norm((w[k >> 2] >> (8 * (k & 3))) & 0xffu, tanh_range);
    store8(d + i * 16, f);
    store8(d + i * 16 + 8, f + 8);
  }
}

template <typename T>
__global__ void u8_norm_kernel(const uint8_t* __restrict__ s, T* __restrict__ d, long long n, int tanh_range) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    stf(d + i, u8_norm(s[i], tanh_range));
}

extern "C" int hm_u8_normalize(const uint8_t* src, void* dst, int dtype, long long n, int tanh_range, void* stream) {
  CHECK_DTYPE(dtype, "hm_u8_normalize");
  HM_CHECK_ARG(src && dst && n > 0, "hm_u8_normalize: bad argument");
  HM_CHECK_ARG(tanh_range == 0 || tanh_range == 1, "hm_u8_normalize: bad range selector %d", tanh_range);
  cudaStream_t st = (cudaStream_t)stream;
  if (n % 16 == 0 && al16(src) && al16(dst)) {
    DISPATCH_T(dtype, (u8_norm_v16_kernel<T><<<ew_grid(n / 16), 256, 0, st>>>((const uint4*)src, (T*)dst, n / 16,
                                                                            tanh_range)));
  } else {
    DISPATCH_T(dtype, (u8_norm_kernel<T><<<ew_grid(n), 256, 0, st>>>(src, (T*)dst, n, tanh_range)));
  }
  HM_CHECK_LAUNCH("hm_u8_normalize");
  return HM_OK;
}

extern "C" int hm_adv_loss(const void* h, void* dh, int dtype, long long R, int G, int out_act, float target,
                           int lsgan, int relu_head, float weight, float gscale, int accumulate, float* loss,
                           void* stream) {
  CHECK_DTYPE(dtype, "hm_adv_loss");
  HM_CHECK_ARG(h && loss && R > 0 && G > 0, "hm_adv_loss: bad argument");
  DISPATCH_T(dtype, (adv_loss_kernel<T><<<ew_grid(R, 256, 2), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)h, (T*)dh, R, G, out_act, target, lsgan, relu_head, weight, gscale, accumulate,
                        loss)));
  HM_CHECK_LAUNCH("hm_adv_loss");
  return HM_OK;
}

extern "C" int hm_maxpool2_bwd_scaled(const void* dp, const void* p, const uint8_t* idx, void* dx, void* dxs,
                                      const float* scale, int dtype, int B, int H, int W, int C, int act, float slope,
                                      float* db, void* stream) {
  CHECK_DTYPE(dtype, "hm_maxpool2_bwd_scaled");
  HM_CHECK_ARG(dp && p && idx && dxs && scale && B > 0 && H >= 2 && W >= 2 && C > 0,
               "hm_maxpool2_bwd_scaled: bad argument");
  if (!(C % 8 == 0 && 256 % (C / 8) == 0 && al16(dp) && al16(p) && (!dx || al16(dx)) && al16(dxs) && (((uintptr_t)idx) & 7) == 0)) {
    set_error("hm_maxpool2_bwd_scaled: needs C %% 8 == 0, 256 %% (C/8) == 0 and 16-byte aligned tensors (C = %d)", C);
    return HM_ERR_UNSUPPORTED;
  }
  long long n = (long long)B * (H / 2) * (W / 2) * C;
  DISPATCH_T(dtype, (maxpool2_bwd_v8_kernel<T><<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)dp, (const T*)p, idx, (T*)dx, B, H, W, C, act, slope, db, (T*)dxs, scale)));
  HM_CHECK_LAUNCH("hm_maxpool2_bwd_scaled");
  return HM_OK;
}

extern "C" int hm_adv_loss_pair(const void* h, void* dh, void* dhw, float* sw, float* sg, int dtype, long long R, int G,
                                int out_act, int lsgan, int relu_head, float gscale, float* loss_disc, float* loss_gen,
                                void* stream) {
  CHECK_DTYPE(dtype, "hm_adv_loss_pair");
  HM_CHECK_ARG(h && dh && dhw && sw && sg && loss_disc && loss_gen && R > 0 && G > 0, "hm_adv_loss_pair: bad argument");
  DISPATCH_T(dtype, (adv_loss_pair_kernel<T><<<ew_grid(R, 256, 2), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)h, (T*)dh, (T*)dhw, sw, sg, R, G, out_act, lsgan, relu_head, gscale, loss_disc,
                        loss_gen)));
  HM_CHECK_LAUNCH("hm_adv_loss_pair");
  return HM_OK;
}

extern "C" int hm_scale_rows(const void* src, const float* scale, void* dst, int dtype, long long R, long long L,
                             void* stream) {
  CHECK_DTYPE(dtype, "hm_scale_rows");
  HM_CHECK_ARG(src && scale && dst && R > 0 && L > 0, "hm_scale_rows: bad argument");
  DISPATCH_T(dtype, (scale_rows_kernel<T><<<ew_grid(R * L), 256, 0, (cudaStream_t)stream>>>((const T*)src, scale, (T*)dst,
                                                                                           R, L)));
  HM_CHECK_LAUNCH("hm_scale_rows");
  return HM_OK;
}

extern "C" int hm_recon_loss(const void* p, const void* y, void* dp, int dtype, long long n, int l2,
                             float weight, float gscale, int accumulate, float* loss, void* stream) {
  CHECK_DTYPE(dtype, "hm_recon_loss");
  HM_CHECK_ARG(p && y && loss && n > 0, "hm_recon_loss: bad argument");
  DISPATCH_T(dtype, (recon_loss_kernel<T><<<ew_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)p, (const T*)y, (T*)dp, n, l2, weight, gscale, accumulate, loss)));
  HM_CHECK_LAUNCH("hm_recon_loss");
  return HM_OK;
}

extern "C" int hm_rmsprop(float* p, const float* g, float* acc, long long n, const float* lr, float rho,
                          float eps, float gscale, void* stream) {
  HM_CHECK_ARG(p && g && acc && lr && n > 0, "hm_rmsprop: bad argument");
  rmsprop_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, acc, n, lr, rho, eps, gscale);
  HM_CHECK_LAUNCH("hm_rmsprop");
  return HM_OK;
}

extern "C" int hm_adam(float* p, const float* g, float* m, float* v, long long n, const float* lr, float b1,
                       float b2, float eps, int t, float gscale, void* stream) {
  HM_CHECK_ARG(p && g && m && v && lr && n > 0 && t > 0, "hm_adam: bad argument");
  float corr = sqrtf(1.f - powf(b2, (float)t)) / (1.f - powf(b1, (float)t));
  adam_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, b1, b2, eps, corr, gscale);
  HM_CHECK_LAUNCH("hm_adam");
  return HM_OK;
}

extern "C" int hm_inc_i32(int* counter, void* stream) {
  HM_CHECK_ARG(counter, "hm_inc_i32: null pointer");
  inc_i32_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter);
  HM_CHECK_LAUNCH("hm_inc_i32");
  return HM_OK;
}

extern "C" int hm_adam_dev(float* p, const float* g, float* m, float* v, long long n, const float* lr, float b1,
                           float b2, float eps, const int* t_dev, float gscale, void* stream) {
  HM_CHECK_ARG(p && g && m && v && lr && t_dev && n > 0, "hm_adam_dev: bad argument");
  adam_dev_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, b1, b2, eps, t_dev, gscale);
  HM_CHECK_LAUNCH("hm_adam_dev");
  return HM_OK;
}
