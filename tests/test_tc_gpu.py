"""tcgen05/TMA implicit-GEMM convolution (fast mode) against the SIMT gather convolution on identical fp16
inputs and weights, through the C ABI.  Both accumulate in fp32 and round the result to fp16 once, so the
tolerance is 3e-3 of the output scale (one fp16 ulp at the top of the range plus summation-order noise)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import tc_probe   # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", tc_probe.CASES, ids=[c[0] for c in tc_probe.CASES])
def test_tc_conv_matches_simt(case):
    rel, line = tc_probe.run_case(*case)
    assert rel <= 3e-3, line


@pytest.mark.parametrize("case", tc_probe.WGRAD_CASES, ids=[c[0] for c in tc_probe.WGRAD_CASES])
def test_tc_wgrad_matches_simt(case):
    rel, line = tc_probe.run_wgrad_case(*case)
    assert rel <= 3e-3, line


@pytest.mark.parametrize("case", tc_probe.UP2_CASES, ids=[c[0] for c in tc_probe.UP2_CASES])
def test_tc_up2conv_matches_simt_gather(case):
    """nearest-2x upsampling + 5x5 convolution evaluated as four 3x3 phase convolutions on the low-res tensor
    (summed fp16 phase filters) against the gather through the virtual upsampling: 5e-3 of the output scale."""
    rel, line = tc_probe.run_up2_case(*case)
    assert rel <= 5e-3, line


@pytest.mark.parametrize("case", tc_probe.S2_CASES, ids=[c[0] for c in tc_probe.S2_CASES])
def test_tc_stride2_conv_fwd_wgrad_dgrad(case):
    """3x3 stride-2 convolutions (pix2pix encoder / PatchGAN): TMA element-stride forward and weight gradient,
    phase-decomposed input gradient, each against the SIMT kernel on the same fp16 data (3e-3 of the scale)."""
    rel, line = tc_probe.run_s2_case(*case)
    assert rel <= 3e-3, line


@pytest.mark.parametrize("case", tc_probe.C1_CASES, ids=[c[0] for c in tc_probe.C1_CASES])
def test_c1s2_conv_pooled_first_layer_and_last_layer_input_gradient(case):
    """hm_c1s2_conv (in-kernel im2col of a one-channel image, tcgen05): the discriminator's conv5x5(1->64)+LeakyReLU+
    max-pool in one pass (values and argmax) and the generator's last-layer input gradient, against torch float32 on
    the same fp16 data: 3e-3 of the output scale (fp16 operands and output, fp32 accumulation)."""
    rel, line = tc_probe.run_c1_case(*case)
    assert rel <= 3e-3, line


@pytest.mark.parametrize("case", tc_probe.C1B_CASES, ids=[c[0] for c in tc_probe.C1B_CASES])
def test_c1s2_bwd_pooled_first_layer_gradients(case):
    """hm_c1s2_bwd + hm_c1s2_bwd_fold + hm_c1s2_col2im: weight, bias and input gradient of the discriminator's
    conv5x5(1->64)+LeakyReLU+max-pool straight from the pooled tensor's gradient, against float32 torch adjoints
    on the same fp16 data: 3e-3 of each gradient's scale (g*act' and the patch-space gradient are rounded to fp16)."""
    rel, line = tc_probe.run_c1bwd_case(*case)
    assert rel <= 3e-3, line


@pytest.mark.parametrize("case", tc_probe.C1B_CASES, ids=[c[0].replace("c1bwd", "c1wg") for c in tc_probe.C1B_CASES])
def test_c1s2_wgrad_generator_output_layer(case):
    """hm_c1s2_wgrad + hm_unpack_conv_wgrad(mode 14): the weight gradient of the generator's last layer (nearest-2x ->
    conv5x5, 64 -> 1; reference architectures/dcgan.py:31-32) as 6x6 stride-2 patches of the one-channel dy against the
    low-res source rows, against torch autograd in float32 on the same fp16 data: 1e-3 of the gradient's scale (fp16
    operands are exact, fp32 accumulation over B*H*W/4 terms in tile order)."""
    rel, line = tc_probe.run_c1wg_case(*case)
    assert rel <= 1e-3, line


@pytest.mark.parametrize("case", tc_probe.DC2_CASES, ids=[c[0] for c in tc_probe.DC2_CASES])
def test_tc_deconv_2x2_stride2_all_phases(case):
    """Deconv2DLayer 2x2 stride 2 (pix2pix U-Net output layer, reference architectures/p2p.py:272) as one tensor-core
    launch (1x1 GEMM with N = (phase, co) + depth-to-space epilogue, concat source, tanh) and its input gradient over
    hm_s2d_pad64, against torch conv_transpose2d / its adjoint in float32: 3e-3 of the output scale."""
    rel, line = tc_probe.run_dc2_case(*case)
    assert rel <= 3e-3, line


# ---- HM_BF16X3 ("tc32" precision): float32-grade contractions on the same tcgen05 kernels ---------------------------------
# Operands are three-plane bf16 splits of float32 tensors (hm_split_bf16x3), six exact products accumulated in fp32 TMEM,
# float32 results; compared with the SIMT float32 kernels: 5e-5 of the output scale (measured <= 2.5e-5: summation-order noise of up to
# ~10^4-term fp32 sums; a dropped or misplaced plane shows up at >= 4e-3).
TC32_TOL = 5e-5


@pytest.mark.parametrize("case", tc_probe.CASES, ids=[c[0] for c in tc_probe.CASES])
def test_tc32_conv_matches_simt_fp32(case):
    rel, line = tc_probe.run_case32(*case)
    assert rel <= TC32_TOL, line


@pytest.mark.parametrize("case", tc_probe.WGRAD_CASES, ids=[c[0] for c in tc_probe.WGRAD_CASES])
def test_tc32_wgrad_matches_simt_fp32(case):
    rel, line = tc_probe.run_wgrad_case32(*case)
    assert rel <= TC32_TOL, line


@pytest.mark.parametrize("case", [c for c in tc_probe.UP2_CASES if c[5] % 32 == 0 or c[5] <= 4],
                         ids=[c[0] for c in tc_probe.UP2_CASES if c[5] % 32 == 0 or c[5] <= 4])
def test_tc32_up2conv_matches_simt_fp32(case):
    rel, line = tc_probe.run_up2_case32(*case)
    assert rel <= TC32_TOL, line


@pytest.mark.parametrize("case", tc_probe.S2_CASES, ids=[c[0] for c in tc_probe.S2_CASES])
def test_tc32_stride2_fwd_wgrad_dgrad_match_simt_fp32(case):
    rel, line = tc_probe.run_s2_case32(*case)
    assert rel <= TC32_TOL, line


@pytest.mark.parametrize("tc32", [False, True], ids=["fp16", "tc32"])
@pytest.mark.parametrize("case", tc_probe.UP2_BWD_CASES, ids=[c[0] for c in tc_probe.UP2_BWD_CASES])
def test_tc_up2conv_backward_on_the_low_res_grid(case, tc32):
    """Backward of nearest-2x + 5x5 (the generator's layers, reference architectures/dcgan.py:21-22,31-32) evaluated on the
    LOW-res grid: input gradient as one 6x6 stride-2 convolution of dy (pack mode 20), weight gradient as the gradient of
    the four 3x3 phase filters (unpack mode 8) -- 36 instead of 100 low-res taps each -- against the SIMT kernels through
    the virtual upsampling: 3e-3 of the scale on fp16 data (the high-res fp16 input gradient is rounded once more before
    hm_upsample2_bwd sums it), 1e-4 in tc32 mode (measured <= 5.5e-5: two float32 summation orders of up to 9216 terms)."""
    rel, line = tc_probe.run_up2_bwd_case(*case, tc32=tc32)
    assert rel <= (1e-4 if tc32 else 3e-3), line


@pytest.mark.parametrize("case", tc_probe.POOL_CASES, ids=[c[0] for c in tc_probe.POOL_CASES])
def test_tc_conv_with_fused_maxpool(case):
    """hm_tc_conv_pool (the discriminator's conv + LeakyReLU + 2x2 max-pool, reference architectures/dcgan.py:42-47, in the
    tensor-core epilogue) against hm_tc_conv + hm_maxpool2_fwd: identical pooled values, and an argmax whose un-pooled value
    IS the pooled value (ragged widths, two-source, Cout = 64 .. 256 in 128-column tiles)."""
    bad, line = tc_probe.run_pool_case(*case)
    assert bad == 0.0, line
