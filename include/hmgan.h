/* hmgan.h — C ABI of the B200 (sm_100a) kernel library behind the gan-heightmaps
 * training step.
 *
 * What this boundary replaces.  The reference has no native code: every number
 * on its hot path is produced by a compiled Theano function,
 * train_fn(Z,X,Y) (reference pix2pix.py:142), built from Lasagne layer objects
 * (architectures/dcgan.py:14-58, architectures/p2p.py:20-27,126-292,
 * architectures/layers.py:13-26) and Lasagne update rules (pix2pix.py:131-141).
 * A Theano/Lasagne maintainer would bind these entry points as the `perform`
 * bodies of the ops those layers lower to; each function
 * below names the Lasagne/Theano op it stands in for.
 *
 * Conventions (all functions):
 *   - every pointer is a BORROWED DEVICE pointer (the host side, PyTorch's caching
 *     allocator, owns the memory); the library never allocates or frees;
 *   - activations are NHWC, dtype HM_F32 (parity mode) or HM_F16 (fast mode);
 *     statistics, losses, master parameters, gradients and optimiser state are
 *     always float32 (double where stated);
 *   - launches are asynchronous on `stream` (a cudaStream_t passed as void*), no
 *     internal synchronisation, so a whole step can be captured in a CUDA graph;
 *   - return value: 0 on success, a negative HmStatus on error; nothing throws
 *     across the ABI; hm_last_error_string() describes the last error of the
 *     calling thread;
 *   - one CUDA device per process (the design is one process per GPU): the SM
 *     count and the opt-in shared-memory attributes of the kernels are cached
 *     process-wide on first use, for the device current at that moment.  Calls
 *     from several host threads on different streams are safe (those caches are
 *     idempotent, the error string is thread-local); the HMGAN_* environment
 *     variables the kernels' host wrappers consult are diagnostic knobs.
 */
#ifndef HMGAN_H_
#define HMGAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum HmStatus {
  HM_OK = 0,
  HM_ERR_BAD_ARG = -1,
  HM_ERR_UNSUPPORTED = -2,
  HM_ERR_CUDA = -3,
  HM_ERR_ALIGN = -4
} HmStatus;

/* HM_BF16X3 (tensor-core kernels only): the operands are three-plane bfloat16 splits of float32 tensors written by
 * hm_split_bf16x3 and the results are float32 -- float32-grade contractions on the tcgen05 pipe ("tc32" mode). */
typedef enum HmDType { HM_F32 = 0, HM_F16 = 1, HM_BF16X3 = 2 } HmDType;

typedef enum HmAct {
  HM_ACT_LINEAR = 0,  /* lasagne.nonlinearities.linear                       */
  HM_ACT_LRELU = 1,   /* LeakyRectify(slope): dcgan.py:24,45 (0.2); p2p 0.01 */
  HM_ACT_RELU = 2,    /* rectify (Conv2DLayer default, dcgan.py:50)          */
  HM_ACT_SIGMOID = 3, /* dcgan.py:32                                         */
  HM_ACT_TANH = 4     /* p2p.py:275                                          */
} HmAct;

typedef enum HmUp {
  HM_UP_NONE = 0,
  HM_UP_NEAREST2 = 1, /* lasagne Upscale2DLayer(2), dcgan.py:31                      */
  HM_UP_BILINEAR2 = 2 /* theano bilinear_upsampling(ratio=2), layers.py:21-26        */
} HmUp;

/* One gather-GEMM convolution problem.  The same descriptor drives forward,
 * input-gradient and weight-gradient kernels.
 *
 * Forward (transposed=0):  for every logical output site (n,oy,ox), oy<Ho, ox<Wo
 *   acc[co] = sum_{r<kh,s<kw,ci<C1+C2} src(n, oy*stride-pad+r, ox*stride-pad+s, ci) * w[(r*kw+s)*(C1+C2)+ci][co]
 * where src is the channel-concatenation of x1 (C1 ch) and x2 (C2 ch, may be 0)
 * seen through the virtual upsampling `up` (physical tensors are [B,H,W,C*];
 * the virtual grid is [B,H<<(up!=0),W<<(up!=0)]); out-of-range taps read 0.
 * The result is written to y[n, oy*os+ou, ox*os+ov, co] of a physical
 * [B,oH,oW,Cout] tensor after  +bias, activation.
 *
 * Input gradient (transposed=1): the logical grid (Ho,Wo) is the grid of the
 * tensor receiving the gradient, src is the upstream gradient [B,H,W,C1]:
 *   acc[co] = sum_{r,s,ci} src(n,(oy+pad-r)/stride,(ox+pad-s)/stride,ci) * w[...]   (only where divisible)
 * and channels [0,split) of the result go to y, channels [split,Cout) to y2.
 *
 * Weights are "packed" by hm_pack_conv_weight so that all kernels do plain
 * correlation (Lasagne's filter flip is applied at pack time).
 */
typedef struct HmConvDesc {
  int32_t dtype;               /* HmDType of x1,x2,y,y2 and packed weights              */
  int32_t B, H, W, C1, C2;     /* physical source tensors                               */
  int32_t up;                  /* HmUp, forward gather only                             */
  int32_t kh, kw, stride, pad;
  int32_t transposed;          /* 0 forward gather, 1 input-gradient gather             */
  int32_t Ho, Wo, Cout;        /* logical output grid and GEMM N                        */
  int32_t oH, oW, os, ou, ov;  /* physical output tensor and scatter                    */
  int32_t split;               /* channels written to y (rest to y2); ==Cout if no y2   */
  int32_t act;                 /* HmAct epilogue                                        */
  float slope;                 /* LeakyRectify leakiness                                */
  int32_t accumulate;          /* bit0: y += result, bit1: y2 += result                 */
} HmConvDesc;

/* ---- library -------------------------------------------------------------- */
int hm_version(void);
const char* hm_last_error_string(void);
/* 1 if the tcgen05/TMA path can run on the current device (cc 10.x), else 0. */
int hm_device_supported(void);

/* ---- gather-GEMM convolutions (SIMT, any shape; fp32 accumulate) ----------
 * Stand in for theano CorrMM / CorrMM_gradInputs / CorrMM_gradWeights as used by
 * lasagne Conv2DLayer / Deconv2DLayer / DenseLayer (dcgan.py:16,22,32,42,50;
 * p2p.py:20-24). */
int hm_conv_gather(const HmConvDesc* d, const void* x1, const void* x2, const void* w_packed,
                   const float* bias, void* y, void* y2, void* stream);
/* dWp[(r*kw+s)*(C1+C2)+ci][co] (+)= sum_sites src(...) * dy[n,oy,ox,co]; fp32 output,
 * atomically accumulated into `dw_packed` (caller zeroes it).  `dy` is a dense
 * [B,Ho,Wo,Cout] tensor unless os>1 (then read through the same scatter as y). */
int hm_conv_wgrad(const HmConvDesc* d, const void* x1, const void* x2, const void* dy,
                  float* dw_packed, void* stream);

/* ---- tcgen05 / TMA implicit-GEMM convolution (fast mode) -------------------
 * The tensor-core implementation of the same contract as hm_conv_gather for the GEMM-shaped layers:
 * fp16 NHWC activations, stride 1, no virtual upsampling, C1 and C2 multiples of 64, Cout a multiple of 16,
 * dense output.  `w_tc` is the K-major pack [kh*kw][Cout][C1+C2] (hm_pack_conv_weight modes 5/6).
 * hm_tc_conv_supported() returns 1 when a descriptor qualifies; hm_tc_conv() returns HM_ERR_UNSUPPORTED
 * otherwise (the caller then uses hm_conv_gather). */
int hm_tc_conv_supported(const HmConvDesc* d);
int hm_tc_conv(const HmConvDesc* d, const void* x1, const void* x2, const void* w_tc, const float* bias,
               void* y, void* y2, void* stream);
/* Convolution + bias + activation + MaxPool2DLayer(2) in ONE pass (dcgan.py:42-47): y_pooled[B,Ho/2,Wo/2,Cout] and
 * idx[B,Ho/2,Wo/2,Cout] (argmax 0..3, first maximum wins: the layout of hm_maxpool2_fwd); the un-pooled activation is
 * never written.  Plain stride-1 fp16 convolutions with rows of at least 128 pixels, even Ho / Wo, Cout % 32 == 0 and a
 * monotonic activation (hm_tc_conv_pool_supported); the maximum is taken over the fp32 values before rounding. */
int hm_tc_conv_pool_supported(const HmConvDesc* d);
int hm_tc_conv_pool(const HmConvDesc* d, const void* x1, const void* x2, const void* w_tc, const float* bias,
                    void* y_pooled, uint8_t* idx, void* stream);
/* hm_tc_conv with a caller-provided scratch workspace of ws_bytes >= hm_tc_conv_ws_bytes(d) bytes (no initialisation
 * needed, contents undefined afterwards; one buffer serves every call issued on a stream).  With it, layers whose
 * tiles fill less than half of the SMs split K over the idle ones: every K slice stores its partial sums to its own
 * plane of the workspace and one finishing pass adds the planes in a fixed order (deterministic) and applies bias /
 * activation exactly as hm_tc_conv does.  hm_tc_conv_ws_bytes() answers 0 -- and hm_tc_conv_ws() behaves as
 * hm_tc_conv() -- for shapes that do not profit or when HMGAN_TC_SPLITK=0.  ws may be null. */
long long hm_tc_conv_ws_bytes(const HmConvDesc* d);
int hm_tc_conv_ws(const HmConvDesc* d, const void* x1, const void* x2, const void* w_tc, const float* bias,
                  void* y, void* y2, void* ws, long long ws_bytes, void* stream);
/* Weight gradient on the tensor cores, same contract as hm_conv_wgrad (fp32 [kh*kw*(C1+C2)][Cout], atomically
 * accumulated, caller zeroes): fp16, stride 1, no virtual upsampling, C1, C2 and Cout multiples of 64, dense dy. */
int hm_tc_wgrad_supported(const HmConvDesc* d);
int hm_tc_wgrad(const HmConvDesc* d, const void* x1, const void* x2, const void* dy, float* dw_packed,
                void* stream);

/* float32-grade contractions on the same tcgen05 kernels (dtype HM_BF16X3, the "tc32" parity mode; csrc/split_bf16.cu).
 * hm_split_bf16x3 splits a float32 [rows][C] tensor into three bfloat16 planes  a = h + m + l  (h = bf16(a),
 * m = bf16(a - h), l = bf16(a - h - m); exact to 2^-24) and writes SIX planes along the axis the GEMM reduces over:
 *   layout 0: dst[rows][6C] = [h|h|m|h|l|m]  -- sources of hm_tc_conv (forward and input-gradient forms); pass C1, C2 = 6x;
 *   layout 1: dst[rows][6C] = [h|m|h|l|h|m]  -- the float32 K-major weight pack (hm_pack_conv_weight with dst_dtype HM_F32,
 *             rows = elements / K); c1 < C splits the two ConcatLayer segments [0,c1) and [c1,C) separately;
 *   layout 2: dst[6][rows][C] = h;h;m;h;l;m  -- x of hm_tc_wgrad, stacked along the batch axis (pass B = 6x);
 *   layout 3: dst[6][rows][C] = h;m;h;l;h;m  -- dy of hm_tc_wgrad.
 * With these operands the kernels accumulate h.h + h.m + m.h + h.l + l.h + m.m in fp32 TMEM (every product exact, the
 * dropped terms <= 2^-24 relative) and hm_tc_conv stores FLOAT32 results (y, y2 are float tensors; `accumulate` adds to
 * float values). */
int hm_split_bf16x3(const float* src, void* dst, long long rows, int C, int c1, int layout, void* stream);

/* Re-layouts that put the two thin layers of the DCGAN on the tensor cores (fp16 only):
 *  hm_im2col_c1: xc[B,H,W,64], xc[p][t] = x[p + tap t - pad] for the kh*kw taps of a ONE-channel image, 0 beyond; the
 *    first discriminator convolution (dcgan.py:42, Cin = 1) is then the 1x1 convolution xc x pack(mode 11) and its
 *    weight gradient hm_tc_wgrad on (xc, dy), whose first kh*kw rows are the usual packed gradient;
 *  hm_s2d_pad64: out[B,h,w,64], out[q][ph*Co+co] = dy[2q+phase ph][co], 0 beyond 4*Co; the weight gradient of the
 *    generator's last layer (dcgan.py:31-32, Cout = 1) is hm_tc_wgrad(3x3 on the low-res source, out) folded with
 *    hm_unpack_conv_wgrad(mode 10). */
int hm_im2col_c1(const void* x, void* xc, int B, int H, int W, int kh, int kw, int pad, void* stream);
int hm_s2d_pad64(const void* dy, void* out, int B, int h, int w, int Co, void* stream);
/* hm_im2col_thin: the same for 1..4 source channels (optionally split over two tensors = a ConcatLayer, p2p.py:279-281) and
 * stride 1 or 2: xc[B,Ho,Wo,64], xc[p][(r*kw+s)*C + c] = src[p*stride + (r,s) - pad][c], 0 beyond kh*kw*C <= 64.  The
 * PatchGAN's and the U-Net's first convolutions (p2p.py:145,285; 3x3 stride 2, 4 and 1 input channels) are then 1x1
 * tensor-core GEMMs over xc (pack / unpack mode 19). */
int hm_im2col_thin(const void* x1, const void* x2, void* xc, int B, int H, int W, int C1, int C2, int kh, int kw,
                   int stride, int pad, int Ho, int Wo, void* stream);

/* Weight gradient of (Upscale2DLayer(2) -> 5x5 'same' Conv2DLayer) with <= 4 output channels (the generator's last
 * layer, dcgan.py:31-32) as four 3x3 problems on the low-res source; `d` is the layer's forward descriptor
 * (up = HM_UP_NEAREST2).  dw_phases is fp32 [4][9*Cin][Cout], atomically accumulated (caller zeroes); fold onto the
 * 5x5 filter with hm_unpack_conv_wgrad(mode 9). */
int hm_up2conv_wgrad_phases(const HmConvDesc* d, const void* x, const void* dy, float* dw_phases, void* stream);

/* One-channel-source convolution with the im2col operand built in shared memory (tcgen05, fp16 only; csrc/c1_conv.cu).
 * x is a ONE-channel image [B,H,W] (H, W even); every element q of the half-resolution grid reads the 6x6 patch
 * A[q][u*6+v] = x[2qy-2+u][2qx-2+v] (zero outside) and y[B,H/2,W/2,64] = act(A . wk^T + bias):
 *   ncols == 64 : wk = [64][64] fp16 (hm_pack_conv_weight mode 14: the input gradient of the generator's last layer,
 *                 nearest-2x -> conv5x5(64 -> 1), dcgan.py:31-32, straight from dy[B,H,W] to dx[B,H/2,W/2,64]); idx NULL;
 *   ncols == 256: wk = [(d,co)][64] (mode 15: the discriminator's first layer conv5x5(1 -> 64) + activation + 2x2
 *                 max-pool, dcgan.py:42-47, in one pass): y = act(max_d + bias[co]), idx[B,H/2,W/2,64] = argmax d
 *                 (d = 2*dy+dx, first maximum) in bits 0-1, the layout hm_maxpool2_fwd writes, plus bit 2 = the stored y
 *                 is on the slope-1 side of the activation (y >= 0; y > 0 for ReLU) for hm_c1s2_bwd.
 * act: HM_ACT_LINEAR, HM_ACT_LRELU or HM_ACT_RELU (monotonic, so it commutes with the max). */
int hm_c1s2_conv(const void* x, const void* wk, const float* bias, void* y, uint8_t* idx, int B, int H, int W,
                 int ncols, int act, float slope, void* stream);
/* Backward of the pooled form from g = d loss / d (pooled output) [B,H/2,W/2,64]; `pooled`, `idx` as written by the
 * forward pass; `pooled` may be NULL, act' is then taken from bit 2 of idx (one tensor less to read: what the engine
 * does).  With G'[w][(d,co)] = g[w][co] * act'(pooled[w][co]) * [idx[w][co] == d]:
 *   dwk != NULL: dwk[(d,co)][k] (fp32 [256][64], caller zeroes) += sum_w G'[w][(d,co)] * A[w][k], A[w][36] = 1;
 *                hm_c1s2_bwd_fold turns it into dW[co][0][5][5] (Lasagne layout) and db[co];
 *   u   != NULL: u[B,H/2,W/2,64], u[w][k] = sum_(d,co) G'[w][(d,co)] * wk2[d*64+k][co] (wk2 = pack mode 16);
 *                hm_c1s2_col2im sums the patch contributions into dx[B,H,W] (one channel).
 * The full-resolution gradient of the un-pooled activation is never materialised. */
int hm_c1s2_bwd(const void* x, const void* g, const void* pooled, const uint8_t* idx, const void* wk2, float* dwk,
                void* u, const float* img_scale /* [B] or NULL: g of image b times img_scale[b] */, int B, int H, int W,
                int act, float slope, void* stream);
int hm_c1s2_bwd_fold(const float* dwk, float* dw, float* db, int cout, void* stream);
/* Weight gradient of the generator's last layer (nearest-2x -> conv5x5 'same', 64 -> 1 channel; dcgan.py:31-32) from the
 * one-channel dy[B,H,W] and the low-res source x[B,H/2,W/2,64]: the kernel of hm_c1s2_bwd with the roles swapped (6x6
 * stride-2 patches of dy against the source rows), dwk[ci][u*6+v] (fp32 [64][64], caller zeroes) +=
 * sum_q x[q][ci] * dy[2q-2+(u,v)] = the gradient of pack mode 14's operand; hm_unpack_conv_wgrad(mode 14) folds it onto
 * the 5x5 filter.  x is read once; no regrouped copy of dy is materialised. */
int hm_c1s2_wgrad(const void* dy, const void* x, float* dwk, int B, int H, int W, void* stream);
int hm_c1s2_col2im(const void* u, void* dx, int B, int H, int W, void* stream);

/* Weight (un)packing between Lasagne master layout and the packed [K][Cout] layout.
 *  mode 0: Conv2DLayer W (Cout,Cin,kh,kw)      -> Wp[(r*kw+s)*Cin+ci][co] = W[co][ci][kh-1-r][kw-1-s]   (forward)
 *  mode 1: Conv2DLayer W                       -> Wp[(r*kw+s)*Cout+co][ci] = W[co][ci][kh-1-r][kw-1-s]  (input gradient)
 *  mode 2: Deconv2DLayer W (Cin,Cout,kh,kw), tap (u,v) -> Wp[ci][co] = W[ci][co][kh-1-u][kw-1-v]        (forward, one output phase)
 *  mode 3: Deconv2DLayer W                     -> Wp[(u*kw+v)*Cout+co][ci] = W[ci][co][kh-1-u][kw-1-v]  (input gradient = strided conv)
 *  mode 4: DenseLayer W (in,out)               -> Wp = W (dtype cast only)
 *  mode 5: Conv2DLayer W, tcgen05 forward pack (K-major)  -> Wt[(r*kw+s)][co][ci] = W[co][ci][kh-1-r][kw-1-s]
 *  mode 6: Conv2DLayer W, tcgen05 input-gradient pack     -> Wt[(r*kw+s)][ci][co] = W[co][ci][r][s]
 *          (the input gradient of a stride-1 'same' convolution is the forward correlation of dy with this
 *           pack and pad' = k-1-pad)
 *  mode 8: nearest-2x + 5x5 as four 3x3 phase filters, mode 11: one-channel input over the im2col tensor, modes 14/15/16: hm_c1s2_* operands, mode 19: thin-source convolution over hm_im2col_thin's tensor, modes 17/18: Deconv2DLayer 2x2 stride 2 on the tensor cores (all phases; its input gradient over hm_s2d_pad64), mode 12: input
 *          gradient of a 3x3 stride-2 convolution as a 2x2-tap phase convolution of dy (see csrc/simt_conv.cu)
 *  mode 20: input gradient of (Upscale2DLayer(2) -> 5x5 'same' conv, dcgan.py:31-32 / :21-22) as ONE 6x6 stride-2 pad-2
 *          convolution of dy landing directly on the low-res source grid -> Wt[(u*6+v)][ci][co] (K-major, K = co): the
 *          adjoint of the four 3x3 phase filters of mode 8; 36 taps on H x W pixels instead of 25 taps on 2H x 2W
 *          followed by hm_upsample2_bwd (mode 14 is its Cout == 1 case for hm_c1s2_conv)
 *  mode 22: Deconv2DLayer W (Cin,Cout,2,2) -> Wt[ci][(u*2+v)*Cout+co] = W[ci][co][1-u][1-v] (K-major, K = 4*Cout): mode
 *          18 without the padding.  A 2x2 deconvolution of a 1x1 input (the U-Net bottleneck, p2p.py:197-198) is a dense
 *          layer Cin -> (u,v,co): forward = 1x1 hm_tc_conv with pack mode 17 (bias tiled over the four positions), input
 *          gradient = 1x1 hm_tc_conv over dy[B,1,1,4*Cout] with this pack, weight gradient = 1x1 hm_tc_wgrad, whose
 *          [ci][4*Cout] result hm_unpack_conv_wgrad(mode 17) takes with leading dimension 4*Cout (64 when 4*Cout <= 64)
 *  mode 21: DenseLayer W (in,out) -> Wt[co][ci] (K-major): the layer as a 1x1 tensor-core convolution over a [B,1,1,in]
 *          tensor; `in` may be any multiple of 8 (hm_tc_conv zero-fills the last 64-channel slice, e.g. latent_dim 1000)
 *  mode 7: Conv2DLayer W, the same input-gradient-as-forward form in the gather layout
 *          -> Wp[(r*kw+s)*Cout+co][ci] = W[co][ci][r][s]   (used when dy has <= 4 channels: thin-input kernel)
 * `dst_dtype` is the HmDType of the packed copy.  hm_unpack_conv_wgrad applies the
 * inverse index map of mode 0 / 2(all taps) / 4 to a packed fp32 gradient and
 * (over)writes the master-layout gradient; modes 8 / 9 / 10 / 14 / 17 apply the ADJOINT of that pack (they sum). */
int hm_pack_conv_weight(const float* w, void* wp, int mode, int cout, int cin, int kh, int kw,
                        int u, int v, int dst_dtype, void* stream);
/* All packs of a network in one launch.  jobs_dev: n_jobs HmPackJob records in DEVICE memory (same fields as the
 * arguments of hm_pack_conv_weight; n = hm_pack_conv_weight_count(mode, ...) elements), max_n = the largest n. */
typedef struct HmPackJob {
  const float* w;
  void* wp;
  int32_t mode, cout, cin, kh, kw, u, v, pad_;
  long long n;
} HmPackJob;
long long hm_pack_conv_weight_count(int mode, int cout, int cin, int kh, int kw);
int hm_pack_conv_weight_multi(const HmPackJob* jobs_dev, int n_jobs, long long max_n, int dst_dtype, void* stream);
int hm_unpack_conv_wgrad(const float* dwp, float* dw, int mode, int cout, int cin, int kh, int kw,
                         void* stream);

/* ---- BatchNormLayer (dcgan.py:17,23; p2p.py:146-268) ----------------------- */
/* sums[0..C) = sum x, sums[C..2C) = sum x^2 over the M rows of x[M,C]; double, caller zeroes. */
int hm_bn_stats(const void* x, int dtype, long long M, int C, double* sums, void* stream);
/* From the sums: mean, inv_std = 1/sqrt(var+eps) (biased var), scale = gamma*inv_std,
 * shift = beta - mean*scale; if update_running: running <- (1-alpha)*running + alpha*batch
 * for BOTH mean and inv_std (Lasagne averages inv_std).  If sums==NULL (deterministic):
 * scale/shift come from the running statistics. */
int hm_bn_finalize(const double* sums, long long M, int C, const float* gamma, const float* beta,
                   float* running_mean, float* running_inv_std, float eps, float alpha,
                   int update_running, float* mean, float* inv_std, float* scale, float* shift,
                   void* stream);
/* a = act(x*scale[c]+shift[c]) */
int hm_bn_apply_act(const void* x, void* a, int dtype, long long M, int C, const float* scale,
                    const float* shift, int act, float slope, void* stream);
/* Backward, pass 1: with dyh = da * act'(a):  red[0..C) += sum dyh, red[C..2C) += sum dyh*xhat,
 * xhat = (x-mean)*inv_std.  double, caller zeroes. */
int hm_bn_bwd_reduce(const void* da, const void* a, const void* x, int dtype, long long M, int C,
                     const float* mean, const float* inv_std, int act, float slope, double* red,
                     void* stream);
/* Backward, pass 2: dx = gamma*inv_std*(dyh - red0/M - xhat*red1/M); also dgamma=red1, dbeta=red0. */
int hm_bn_bwd_apply(const void* da, const void* a, const void* x, void* dx, int dtype, long long M,
                    int C, const float* mean, const float* inv_std, const float* gamma, int act,
                    float slope, const double* red, float* dgamma, float* dbeta, void* stream);

/* The same two passes without reading x where xhat can be recovered from the layer's OUTPUT a = act(gamma*xhat + beta):
 * xhat = (act^-1(a) - beta)/gamma for linear / leaky-rectify (slope > 0) layers and channels with |gamma| >= 2^-10 (other
 * channels, activations and shapes read x exactly as above).  One tensor less per pass; results agree with the x-based
 * form to the rounding of a.  Used by the fp16 fast mode. */
int hm_bn_bwd_reduce_a(const void* da, const void* a, const void* x, int dtype, long long M, int C, const float* mean,
                       const float* inv_std, const float* gamma, const float* beta, int act, float slope, double* red,
                       void* stream);
int hm_bn_bwd_apply_a(const void* da, const void* a, const void* x, void* dx, int dtype, long long M, int C,
                      const float* mean, const float* inv_std, const float* gamma, const float* beta, int act, float slope,
                      const double* red, float* dgamma, float* dbeta, void* stream);

/* ---- elementwise / pooling / resampling ----------------------------------- */
/* dx = dy * act'(y)   (NonlinearityLayer backward expressed through the OUTPUT y) */
int hm_act_bwd(const void* dy, const void* y, void* dx, int dtype, long long n, int act, float slope,
               int accumulate, void* stream);
/* db[c] (+)= sum_rows dy[M,C]  (conv / dense bias gradient) */
int hm_col_sum(const void* dy, int dtype, long long M, int C, float* db, void* stream);
/* MaxPool2DLayer(2) (dcgan.py:47): p = max over 2x2, idx = argmax (0..3, first max wins) */
int hm_maxpool2_fwd(const void* x, void* p, uint8_t* idx, int dtype, int B, int H, int W, int C,
                    void* stream);
/* dx[.. 2x2 ..] = (k==idx) ? dp * act'(p) : 0   — unpool fused with the backward of the
 * (monotonic) activation that preceded the pool.  db (optional, fp32 [C], caller zeroes) += sum over all pixels of
 * dx: the bias gradient of the convolution that produced the pooled tensor (saves a pass over dx). */
int hm_maxpool2_bwd(const void* dp, const void* p, const uint8_t* idx, void* dx, int dtype, int B,
                    int H, int W, int C, int act, float slope, float* db, void* stream);
/* The same with a per-image weight: dx as above, dxs = scale[image] * dx (a second tensor of dx's shape) and
 * db += scale[image] * (sum of dx).  Used by the single-pass discriminator backward (hm_adv_loss_pair): the
 * input-gradient chain runs on dx, weight and bias gradients on dxs.  Needs C % 8 == 0 and 256 % (C/8) == 0.  dx may be
 * NULL (only dxs and db are produced: the half of the work that belongs on the weight-gradient stream). */
int hm_maxpool2_bwd_scaled(const void* dp, const void* p, const uint8_t* idx, void* dx, void* dxs, const float* scale,
                           int dtype, int B, int H, int W, int C, int act, float slope, float* db, void* stream);
/* dst[r][:] = scale[r] * src[r][:]  (R rows of L elements). */
int hm_scale_rows(const void* src, const float* scale, void* dst, int dtype, long long R, long long L, void* stream);
/* Adjoint of the virtual upsampling: dx[B,H,W,C] (+)= U^T dy[B,2H,2W,C]; mode = HmUp. */
int hm_upsample2_bwd(const void* dy, void* dx, int dtype, int B, int H, int W, int C, int mode,
                     int accumulate, void* stream);
/* Materialised upsampling (fast path feeding the tensor-core conv): y[B,2H,2W,C] = U x. */
int hm_upsample2_fwd(const void* x, void* y, int dtype, int B, int H, int W, int C, int mode,
                     void* stream);
/* Boundary layout changes: NCHW float32 (the reference's numpy convention) <-> NHWC dtype. */
int hm_nchw_to_nhwc(const float* src, void* dst, int dtype, int B, int C, int H, int W, void* stream);
int hm_nhwc_to_nchw(const void* src, float* dst, int dtype, int B, int C, int H, int W, void* stream);
int hm_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, void* stream);
/* uint8 image data in the reference's on-disk NHWC layout -> dtype NHWC, normalised on the device as the reference's
 * iterator does on the host in float32 (util.py:33-35): tanh_range 0: x/255 (grayscale images), 1: (x-127.5)/127.5.
 * n = number of elements (B*H*W*C).  The float32 result is bit-identical to numpy's. */
int hm_u8_normalize(const uint8_t* src, void* dst, int dtype, long long n, int tanh_range, void* stream);
/* dst[M][nc] (+)= src[M][C] channels [c0, c0+nc): the part of a ConcatLayer's gradient that belongs to one input
 * (p2p.py:279-281). */
int hm_slice_channels(const void* src, void* dst, int dtype, long long M, int C, int c0, int nc, int accumulate,
                      void* stream);
/* ReshapeLayer((-1,C,H,W)) after the generator's DenseLayer (dcgan.py:18): the flat feature vector is in
 * NCHW order, the convolutions want NHWC.  inverse=0: src[B][C][H][W] -> dst[B][H][W][C]; inverse=1 the
 * other way (the backward pass).  Same dtype on both sides. */
int hm_permute(const void* src, void* dst, int dtype, int B, int C, int H, int W, int inverse, void* stream);

/* ---- losses (pix2pix.py:102-121) ------------------------------------------ */
/* Adversarial loss on a discriminator head h[R, G] (R rows of G values; the DCGAN head
 * average-pools G=rf*rf post-ReLU values per sample, dcgan.py:50-56; PatchGAN has G=1):
 *   out_r = out_act(mean_g h[r,g])  (out_act: HM_ACT_LINEAR or HM_ACT_SIGMOID);  lsgan: l_r=(out_r-target)^2 ; else BCE(out_r,target)
 *   loss[0] += weight * mean_r l_r ;  dh[r,g] (+)= gscale * weight * dl_r/dout_r /(R*G) * (relu_head ? h>0 : 1)
 * dh may be NULL (loss only). */
int hm_adv_loss(const void* h, void* dh, int dtype, long long R, int G, int out_act, float target,
                int lsgan, int relu_head, float weight, float gscale, int accumulate, float* loss,
                void* stream);
/* Both adversarial losses of the FAKE half of a scalar-output discriminator in one go (pix2pix.py:107-108: the
 * generator loss, target 1, and the fake term of the discriminator loss, target 0, are functions of the same D(G(z))).
 * A discriminator without BatchNorm is sample-wise independent and its backward pass is linear in the output
 * gradient, so for sample r the two backward passes differ only by the scalar factors a_r = dl(out_r,0)/dout and
 * b_r = dl(out_r,1)/dout.  This writes ONE output gradient dh[r,g] = c_r * mask (c_r = the larger of a_r, b_r in
 * magnitude, so nothing underflows in fp16), the weights sw[r] = a_r/c_r (apply to this sample's contribution to the
 * discriminator's weight gradients) and sg[r] = b_r/c_r (apply to its input gradient, which goes to the generator),
 * dhw[r,g] = a_r * mask (= sw * dh, the head convolution's weight-gradient operand), and adds mean_r l(out_r,0) to
 * loss_disc[0], mean_r l(out_r,1) to loss_gen[0].  Arguments as hm_adv_loss. */
int hm_adv_loss_pair(const void* h, void* dh, void* dhw, float* sw, float* sg, int dtype, long long R, int G,
                     int out_act, int lsgan, int relu_head, float gscale, float* loss_disc, float* loss_gen,
                     void* stream);
/* Reconstruction loss: l1: mean|p-y| , l2: mean (p-y)^2 ; dp (+)= gscale*weight*dl/dp. */
int hm_recon_loss(const void* p, const void* y, void* dp, int dtype, long long n, int l2,
                  float weight, float gscale, int accumulate, float* loss, void* stream);

/* ---- optimisers (lasagne.updates, experiments.py:116-117) ------------------ */
/* rmsprop: acc = rho*acc + (1-rho)*g^2 ; p -= lr*g/sqrt(acc+eps)   (eps inside the sqrt).
 * g is multiplied by gscale first (1/world_size after a sum all-reduce, 1/loss_scale). */
int hm_rmsprop(float* p, const float* g, float* acc, long long n, const float* lr, float rho,
               float eps, float gscale, void* stream);
/* adam (lasagne 0.2.dev1): t is the NEW step count. */
int hm_adam(float* p, const float* g, float* m, float* v, long long n, const float* lr, float b1,
            float b2, float eps, int t, float gscale, void* stream);
/* The same with the step count in device memory: *t_dev is the NEW step count (>= 1), advanced on the stream by
 * hm_inc_i32 before the call, so a captured CUDA graph of the step keeps counting on every replay. */
int hm_adam_dev(float* p, const float* g, float* m, float* v, long long n, const float* lr, float b1,
                float b2, float eps, const int* t_dev, float gscale, void* stream);
int hm_inc_i32(int* counter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HMGAN_H_ */
