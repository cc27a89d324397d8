"""Raw uint8 image batches on the GPU: hm_u8_normalize through the C ABI against numpy's float32 arithmetic of the
reference's iterator (util.py:33-35) -- bit-exact for float32 output, bit-exact after one float32->float16 rounding for
fp16 -- and the training step fed raw bytes against the step fed the host-normalised float32 batches."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import step as S

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
import _lib   # noqa: E402
import util   # noqa: E402

from test_engine_cpu import build_pair   # noqa: E402

pytestmark = pytest.mark.gpu
TD = {0: torch.float32, 1: torch.float16}


@pytest.mark.parametrize("n,offset", [(256 * 256 * 3, 0), (4 * 512 * 512, 0), (1000003, 0), (4096, 3), (1, 0), (16, 0)])
@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("tanh_range", [0, 1])
def test_u8_normalize_is_numpy_float32_arithmetic(n, offset, dtype, tanh_range):
    """Vector path (n % 16 == 0, aligned), scalar tail path (odd n, unaligned source), every byte value present."""
    r = np.random.RandomState(n % 997 + offset)
    a = r.randint(0, 256, n + offset).astype(np.uint8)
    a[offset:offset + min(n, 256)] = np.arange(min(n, 256), dtype=np.uint8)
    src = torch.from_numpy(a).cuda()
    dst = torch.full((n + 32,), -7.0, dtype=TD[dtype], device="cuda")
    _lib.call("hm_u8_normalize", src.data_ptr() + offset, dst.data_ptr(), dtype, n, tanh_range, None)
    torch.cuda.synchronize()
    f = a[offset:].astype(np.float32)
    ref = (f - np.float32(127.5)) / np.float32(127.5) if tanh_range else f / np.float32(255.0)
    ref = ref.astype(np.float32 if dtype == 0 else np.float16)
    out = dst.cpu().numpy()
    np.testing.assert_array_equal(out[:n], ref)
    assert np.all(out[n:] == -7.0)            # nothing written past the end


def test_u8_normalize_rejects_bad_arguments():
    src = torch.zeros(16, dtype=torch.uint8, device="cuda")
    dst = torch.zeros(16, device="cuda")
    for args in ((0, dst.data_ptr(), 0, 16, 0), (src.data_ptr(), dst.data_ptr(), 0, 0, 0),
                 (src.data_ptr(), dst.data_ptr(), 5, 16, 0), (src.data_ptr(), dst.data_ptr(), 0, 16, 2)):
        with pytest.raises(Exception):
            _lib.call("hm_u8_normalize", *args, None)


def test_step_from_raw_bytes_matches_step_from_float_batches():
    """64-px gate, float32 parity mode, four steps (the last two replay the captured CUDA graphs): the losses of the
    model fed uint8 NHWC bytes equal those of the model fed util.normalise_uint8 of the same bytes to 1e-4 relative
    (the inputs are bit-identical on the device; only the summation order of atomics differs between runs)."""
    cfg = S.experiment_kwargs('gate64')
    r = np.random.RandomState(4)
    Xu = r.randint(0, 256, (4, 64, 64, 1)).astype(np.uint8)
    Yu = np.repeat(Xu, 3, axis=3)
    Z = np.random.RandomState(1).rand(4, cfg['latent_dim']).astype(np.float32)
    out = []
    for raw in (False, True):
        _, m = build_pair(cfg, 'dcgan', with_p2p=False, device="cuda")
        X, Y = (Xu, Yu) if raw else (util.normalise_uint8(Xu, True), util.normalise_uint8(Yu, False))
        out.append(np.array([m.train_fn(Z, X, Y) for _ in range(4)]))
    assert np.all(np.isfinite(out[1]))
    np.testing.assert_allclose(out[1], out[0], rtol=1e-4, atol=1e-6)
