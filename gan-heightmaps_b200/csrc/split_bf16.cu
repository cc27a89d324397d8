// float32 -> bfloat16 three-plane operand splits for the HM_BF16X3 ("tc32") mode of the tcgen05 kernels.
//
// The reference computes its convolutions in float32 (Theano CorrMM / cuDNN behind lasagne Conv2DLayer, reference
// architectures/dcgan.py:22,42; architectures/p2p.py:20-21), and BASELINE.json asks for 1e-3 relative parity with
// it.  The tensor pipe has no float32 operand type, so the parity gate runs the SAME tcgen05 kernels as the fp16
// benchmark path (tc_conv.cu, tc_wgrad.cu; kind::f16 with BF16 operands) on operands split into three bf16 planes
//     a = h + m + l,   h = bf16(a),  m = bf16(a - h),  l = bf16(a - h - m)      (3 x 8 significand bits: exact to 2^-24)
// and evaluates  a.b ~= h.h + h.m + m.h + h.l + l.h + m.m  (everything down to 2^-16 relative; the dropped m.l, l.m
// and l.l terms are <= 2^-24) inside ONE fp32 TMEM accumulation by concatenating the six products along the GEMM's
// reduction axis: channels for the forward / input-gradient convolutions, the batch (pixel) axis for the weight
// gradient.  Every bf16 x bf16 product is exact in the fp32 accumulator, so the contraction is float32-grade (measured
// against the SIMT fp32 kernels in tests/test_tc_gpu.py; a two-plane split, 2^-17, is not enough: it flips enough
// max-pool ties of the discriminator to move the generator's gradients by 2e-3) while every MMA, TMA box, descriptor
// and TMEM epilogue is the one the fast mode times.
#include <cuda_bf16.h>

#include "hm_common.cuh"

namespace hm {

// plane order along the reduction axis (six planes of the source's extent):
//   "a" side (layouts 0, 2: activations / x):   h h m h l m
//   "b" side (layouts 1, 3: weights / dy):      h m h l h m
// layouts 0/1: dst[row][6C], planes side by side inside a row, per channel segment ([0,c1) and [c1,C): the two sources
//              of a ConcatLayer in a packed weight are split separately, giving [6*c1 | 6*(C-c1)]);
// layouts 2/3: dst[6][rows][C], planes stacked along the batch axis (weight gradient).
__global__ void split_bf16x3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int C,
                                    int c1, int layout) {
  const long long total = rows * (long long)C;
  const bool bside = layout & 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float a = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    const float r1 = a - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
    const __nv_bfloat16 pl[6] = {h, bside ? m : h, bside ? h : m, bside ? l : h, bside ? h : l, m};
    if (layout >= 2) {
#pragma unroll
      for (int k = 0; k < 6; k++) dst[k * total + i] = pl[k];
    } else {
      const long long row = i / C;
      const int c = (int)(i - row * C);
      const int seg0 = c < c1 ? 0 : 6 * c1, w = c < c1 ? c1 : C - c1, cc = c < c1 ? c : c - c1;
      __nv_bfloat16* d = dst + row * 6LL * C + seg0 + cc;
#pragma unroll
      for (int k = 0; k < 6; k++) d[k * w] = pl[k];
    }
  }
}

}  // namespace hm

using namespace hm;

extern "C" int hm_split_bf16x3(const float* src, void* dst, long long rows, int C, int c1, int layout, void* stream) {
  HM_CHECK_ARG(src && dst && rows > 0 && C > 0, "hm_split_bf16x3: bad argument");
  HM_CHECK_ARG(c1 > 0 && c1 <= C, "hm_split_bf16x3: bad first-segment width %d of %d", c1, C);
  HM_CHECK_ARG(layout >= 0 && layout <= 3, "hm_split_bf16x3: bad layout %d", layout);
  const long long total = rows * (long long)C;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  split_bf16x3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, rows, C, c1, layout);
  HM_CHECK_LAUNCH("hm_split_bf16x3");
  return HM_OK;
}
