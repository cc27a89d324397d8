"""Torch-CPU restatement of the Lasagne/Theano ops the reference's hot path uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

Every function cites the reference call site it serves (file:line under
/root/reference) and the upstream semantics it restates.  Tensors are NCHW,
dtype is whatever the caller passes (float32 for the oracle proper, float64 for
the tight cross-checks against numpy_ref).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# layers
# --------------------------------------------------------------------------- #

def conv2d(x, W, b, stride=1, pad="same"):
    """lasagne.layers.Conv2DLayer(..., flip_filters=True) — a TRUE convolution.

    Call sites: architectures/dcgan.py:22,32,42,50 (5x5, stride 1, pad='same');
    architectures/p2p.py:20-21 (3x3 stride 2 'same'; 2x2 stride 1 'valid').
    W is (Cout, Cin, kh, kw).  pad='same' means k//2 on each side ('half');
    output side = (in + 2*pad - k)//stride + 1.
    """
    kh, kw = W.shape[2], W.shape[3]
    if pad == "same":
        p = (kh // 2, kw // 2)
    elif pad == "valid":
        p = (0, 0)
    else:
        p = (int(pad), int(pad))
    return F.conv2d(x, torch.flip(W, dims=(2, 3)), b, stride=stride, padding=p)


def deconv2d(x, W, b, stride=2):
    """lasagne.layers.Deconv2DLayer (= TransposedConv2DLayer), crop=0,
    flip_filters=False.

    Call sites: architectures/p2p.py:23-24 via :197,:205-267 (non-bilinear), :272.
    W is (Cin, Cout, kh, kw).  Lasagne evaluates it as
    AbstractConv2d_gradInputs(filter_flip = not flip_filters = True): the
    gradient, with respect to its input, of a TRUE convolution whose kernel is W.
    torch.conv_transpose2d is the input-gradient of a correlation, hence the
    spatial flip.  Output side = (in-1)*stride + k.
    """
    return F.conv_transpose2d(x, torch.flip(W, dims=(2, 3)), b, stride=stride)


def dense(x, W, b):
    """lasagne.layers.DenseLayer; W is (in, out).  architectures/dcgan.py:16."""
    return x @ W + b


def batch_norm(x, beta, gamma, mean, inv_std, deterministic, eps=1e-4, alpha=0.1):
    """lasagne.layers.BatchNormLayer(axes='auto', epsilon=1e-4, alpha=0.1).

    Call sites: architectures/dcgan.py:17,23,44; architectures/p2p.py:146-268.
    Train: batch mean / biased variance over all axes but 1, and BOTH running
    statistics move as r <- (1-alpha)*r + alpha*batch (note: the running
    *inverse std* is averaged, not the variance).  Returns (y, new_mean,
    new_inv_std); the new_* are None in deterministic mode.
    """
    axes = tuple(i for i in range(x.dim()) if i != 1)
    shape = [1] * x.dim()
    shape[1] = -1
    if deterministic:
        m, s = mean, inv_std
        new_mean = new_inv_std = None
    else:
        m = x.mean(dim=axes)
        v = x.var(dim=axes, unbiased=False)
        s = 1.0 / torch.sqrt(v + eps)
        new_mean = (1 - alpha) * mean + alpha * m.detach()
        new_inv_std = (1 - alpha) * inv_std + alpha * s.detach()
    y = (x - m.view(shape)) * (gamma * s).view(shape) + beta.view(shape)
    return y, new_mean, new_inv_std


def leaky_rectify(x, slope):
    """lasagne.nonlinearities.LeakyRectify(leakiness): 0.2 in dcgan.py:24,45;
    the `leaky_rectify` instance (0.01) everywhere in p2p.py."""
    return torch.where(x >= 0, x, x * slope)


def upscale2d(x, factor=2):
    """lasagne.layers.Upscale2DLayer(mode='repeat'); dcgan.py:31."""
    return x.repeat_interleave(factor, dim=2).repeat_interleave(factor, dim=3)


def bilinear_upsample(x, ratio=2):
    """theano.tensor.nnet.abstract_conv.bilinear_upsampling(use_1D_kernel=True),
    wrapped by BilinearUpsample2DLayer (architectures/layers.py:13-26).

    Upstream construction, restated literally: replicate-pad one pixel on each
    side, then two separable transposed convolutions (rows, then columns) with
    the triangle kernel [1..ratio..1]/ratio, stride `ratio`, cropping
    pad = 2*ratio - (ratio-1)//2 - 1 on the leading side and whatever is left
    on the trailing side so that the output is exactly ratio*n.  For ratio 2:
    y[2m] = x[m], y[2m+1] = (x[m] + x[min(m+1, n-1)])/2.
    """
    B, C, H, W = x.shape
    up = x.reshape(B * C, 1, H, W)
    up = torch.cat([up[:, :, :1, :], up, up[:, :, -1:, :]], dim=2)
    up = torch.cat([up[:, :, :, :1], up, up[:, :, :, -1:]], dim=3)
    pad = 2 * ratio - (ratio - 1) // 2 - 1
    half = torch.arange(1, ratio + 1, dtype=x.dtype)
    kern = torch.cat([half, half[:-1].flip(0)]) / ratio  # [1..r..1]/r
    k = kern.numel()
    # rows: full transposed conv then crop to H*ratio starting at `pad`
    rows = F.conv_transpose2d(up, kern.view(1, 1, k, 1), stride=(ratio, 1))
    rows = rows[:, :, pad:pad + H * ratio, :]
    cols = F.conv_transpose2d(rows, kern.view(1, 1, 1, k), stride=(1, ratio))
    cols = cols[:, :, :, pad:pad + W * ratio]
    return cols.reshape(B, C, H * ratio, W * ratio)


def max_pool(x, size=2):
    """lasagne.layers.MaxPool2DLayer(pool_size=2) (stride=pool_size,
    ignore_border=True); dcgan.py:47."""
    return F.max_pool2d(x, size, size)


def avg_pool_inc_pad(x, size):
    """lasagne.layers.Pool2DLayer(mode='average_inc_pad'), pad 0; dcgan.py:49,52."""
    return F.avg_pool2d(x, size, size)


def apply_nonlinearity(x, name):
    """`name` is one of the strings the host maps Lasagne callables to."""
    if name in (None, "linear"):
        return x
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "rectify":
        return torch.relu(x)
    raise ValueError("unknown nonlinearity %r" % (name,))


# --------------------------------------------------------------------------- #
# objectives / updates / init
# --------------------------------------------------------------------------- #

def squared_error(a, t):
    """lasagne.objectives.squared_error; pix2pix.py:103."""
    return (a - t) ** 2


def binary_crossentropy(p, t):
    """lasagne.objectives.binary_crossentropy -> theano.tensor.nnet.binary_crossentropy;
    pix2pix.py:105."""
    return -(t * torch.log(p) + (1.0 - t) * torch.log(1.0 - p))


def rmsprop_update(p, g, acc, lr, rho=0.9, eps=1e-6):
    """lasagne.updates.rmsprop: acc <- rho*acc + (1-rho)*g^2;
    p <- p - lr*g/sqrt(acc + eps)   (epsilon INSIDE the sqrt).  experiments.py:116."""
    acc_new = rho * acc + (1 - rho) * g * g
    p_new = p - lr * g / torch.sqrt(acc_new + eps)
    return p_new, acc_new


def adam_update(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """lasagne.updates.adam (0.2.dev1 form): t <- t+1;
    a_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMAs; p <- p - a_t*m/(sqrt(v)+eps).
    pix2pix.py:30 (default opt)."""
    t_new = t + 1
    a_t = lr * math.sqrt(1 - b2 ** t_new) / (1 - b1 ** t_new)
    m_new = b1 * m + (1 - b1) * g
    v_new = b2 * v + (1 - b2) * g * g
    p_new = p - a_t * m_new / (torch.sqrt(v_new) + eps)
    return p_new, m_new, v_new, t_new


def glorot_uniform(rng, shape):
    """lasagne.init.GlorotUniform(gain=1): bound sqrt(6/((n1+n2)*receptive_field))
    with n1,n2 = shape[:2] (valid for conv (Cout,Cin,kh,kw), deconv
    (Cin,Cout,kh,kw) and dense (in,out)).  The reference draws from the global
    unseeded np.random; the oracle takes an explicit RandomState."""
    n1, n2 = shape[0], shape[1]
    rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    a = math.sqrt(6.0 / ((n1 + n2) * rf))
    return rng.uniform(-a, a, size=shape).astype(np.float32)
