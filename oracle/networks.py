"""Torch-CPU restatement of the four network factories on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

Each network is a pair of functions: ``*_init(rng, ...)`` returns the parameter
list in ``lasagne.layers.get_all_param_values`` order and
``*_forward(params, x, ...)`` evaluates ``lasagne.layers.get_output`` on it.
Parameters are torch tensors (so autograd can give the oracle its gradients);
BatchNorm running statistics ride in the same list, as in the checkpoints.

forward functions return ``(out, bn_updates)`` where ``bn_updates`` maps a
parameter index to its new running value (Theano default_update of
BatchNormLayer, applied by every compiled function built from a
non-deterministic graph; pix2pix.py:92,99).
"""
import numpy as np
import torch

from . import lasagne_ops as L


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #

def _bn_params(n):
    """BatchNormLayer params in registration order: beta, gamma, mean, inv_std."""
    return [np.zeros(n, np.float32), np.ones(n, np.float32),
            np.zeros(n, np.float32), np.ones(n, np.float32)]


def _conv_params(rng, cout, cin, k):
    return [L.glorot_uniform(rng, (cout, cin, k, k)), np.zeros(cout, np.float32)]


def _deconv_params(rng, cin, cout, k):
    return [L.glorot_uniform(rng, (cin, cout, k, k)), np.zeros(cout, np.float32)]


class _Cursor(object):
    """Walks a parameter list in order and records BN running-stat updates."""

    def __init__(self, params, deterministic):
        self.p = params
        self.i = 0
        self.det = deterministic
        self.updates = {}

    def take(self, n):
        out = self.p[self.i:self.i + n]
        self.i += n
        return out

    def bn(self, x):
        base = self.i
        beta, gamma, mean, inv_std = self.take(4)
        y, nm, ns = L.batch_norm(x, beta, gamma, mean, inv_std, self.det)
        if not self.det:
            self.updates[base + 2] = nm
            self.updates[base + 3] = ns
        return y

    def done(self):
        assert self.i == len(self.p), (self.i, len(self.p))


def trainable_mask(params_meta):
    return [m != "stat" for m in params_meta]


# --------------------------------------------------------------------------- #
# DCGAN generator  (architectures/dcgan.py:14-33)
# --------------------------------------------------------------------------- #

def generator_init(rng, latent_dim, is_a_grayscale, nch=512, h=5, initial_size=4,
                   final_size=512, div=(2, 2, 4, 4, 8, 8, 16), num_repeats=0,
                   dropout_p=0., bilinear_upsample=False):
    units = nch * initial_size * initial_size
    params = [L.glorot_uniform(rng, (latent_dim, units)), np.zeros(units, np.float32)]
    meta = ["W", "b"]
    params += _bn_params(units)
    meta += ["beta", "gamma", "stat", "stat"]
    cin = nch
    for d in div:
        n = nch // d                               # py2 int division, dcgan.py:19
        for _ in range(num_repeats + 1):
            params += _conv_params(rng, n, cin, h) + _bn_params(n)
            meta += ["W", "b", "beta", "gamma", "stat", "stat"]
            cin = n
    params += _conv_params(rng, 1 if is_a_grayscale else 3, cin, h)
    meta += ["W", "b"]
    return params, meta


def generator_forward(params, z, deterministic=False, nch=512, initial_size=4,
                      div=(2, 2, 4, 4, 8, 8, 16), num_repeats=0, bilinear_upsample=False,
                      **_unused):
    c = _Cursor(params, deterministic)
    W, b = c.take(2)
    x = L.dense(z, W, b)                                        # dcgan.py:16 (linear)
    x = c.bn(x)                                                 # dcgan.py:17
    x = x.reshape(-1, nch, initial_size, initial_size)          # dcgan.py:18
    for _ in div:
        for _r in range(num_repeats + 1):
            W, b = c.take(2)
            x = L.conv2d(x, W, b, 1, "same")                    # dcgan.py:22
            x = c.bn(x)                                         # dcgan.py:23
            x = L.leaky_rectify(x, 0.2)                         # dcgan.py:24
        if bilinear_upsample:
            x = L.bilinear_upsample(x, 2)                       # dcgan.py:28
        else:
            x = L.upscale2d(x, 2)                               # dcgan.py:31
    W, b = c.take(2)
    x = torch.sigmoid(L.conv2d(x, W, b, 1, "same"))             # dcgan.py:32
    c.done()
    return x, c.updates


# --------------------------------------------------------------------------- #
# DCGAN discriminator  (architectures/dcgan.py:35-58)
# --------------------------------------------------------------------------- #

def discriminator_init(rng, in_shp, is_a_grayscale, nch=512, h=5,
                       div=(8, 4, 4, 2, 2, 1, 1), num_repeats=0, bn=False,
                       pool_mode='max', nonlinearity='sigmoid'):
    params, meta = [], []
    cin = 1 if is_a_grayscale else 3
    for d in div:
        n = nch // d
        for _ in range(num_repeats + 1):
            params += _conv_params(rng, n, cin, h)
            meta += ["W", "b"]
            if bn:
                params += _bn_params(n)
                meta += ["beta", "gamma", "stat", "stat"]
            cin = n
    params += _conv_params(rng, 1, cin, h)
    meta += ["W", "b"]
    return params, meta


def discriminator_forward(params, x, deterministic=False, nch=512,
                          div=(8, 4, 4, 2, 2, 1, 1), num_repeats=0, bn=False,
                          pool_mode='max', nonlinearity='sigmoid', **_unused):
    c = _Cursor(params, deterministic)
    for _ in div:
        for _r in range(num_repeats + 1):
            W, b = c.take(2)
            x = L.conv2d(x, W, b, 1, "same")                    # dcgan.py:42
            if bn:
                x = c.bn(x)                                     # dcgan.py:44
            x = L.leaky_rectify(x, 0.2)                         # dcgan.py:45
        x = L.max_pool(x, 2) if pool_mode == 'max' else L.avg_pool_inc_pad(x, 2)
    W, b = c.take(2)
    x = torch.relu(L.conv2d(x, W, b, 1, "same"))                # dcgan.py:50 (default rectify)
    rf = nch // (2 ** len(div))                                 # dcgan.py:51
    x = L.avg_pool_inc_pad(x, rf)                               # dcgan.py:52
    x = x.reshape(-1, 1)                                        # dcgan.py:55
    x = L.apply_nonlinearity(x, nonlinearity)                   # dcgan.py:56
    c.done()
    return x, c.updates


# --------------------------------------------------------------------------- #
# pix2pix U-Net generator  (architectures/p2p.py:126-276)
# --------------------------------------------------------------------------- #

_UNET_ENC = (1, 2, 4, 8, 8, 8, 8, 8)            # conv1..conv8 multipliers of nf
_UNET_DEC = (8, 8, 8, 8, 4, 2, 1)               # dconv2..dconv8


def g_unet_init(rng, in_shp, is_a_grayscale, is_b_grayscale, nf=64, act='tanh',
                dropout=False, num_repeats=0, bilinear_upsample=False):
    assert in_shp in [512] or True   # p2p.py:137 asserts 512; the oracle also runs 256-divisible sizes
    assert num_repeats == 0 and not dropout, "unused by every experiment (SURVEY.md App. A)"
    params, meta = [], []
    six = ["W", "b", "beta", "gamma", "stat", "stat"]
    cin = 1 if is_a_grayscale else 3
    enc_ch = []
    for m in _UNET_ENC:
        params += _conv_params(rng, nf * m, cin, 3) + _bn_params(nf * m)
        meta += six
        cin = nf * m
        enc_ch.append(cin)
    params += _conv_params(rng, nf * 8, cin, 2) + _bn_params(nf * 8)          # conv9
    meta += six
    params += _deconv_params(rng, nf * 8, nf * 8, 2) + _bn_params(nf * 8)      # dconv1
    meta += six
    cin = nf * 8 + enc_ch[7]
    for j, m in enumerate(_UNET_DEC):
        if bilinear_upsample:
            params += _conv_params(rng, nf * m, cin, 3)
        else:
            params += _deconv_params(rng, cin, nf * m, 2)
        params += _bn_params(nf * m)
        meta += six
        cin = nf * m + enc_ch[6 - j]
    params += _deconv_params(rng, cin, 1 if is_b_grayscale else 3, 2)          # dconv9
    meta += ["W", "b"]
    return params, meta


def g_unet_forward(params, x, deterministic=False, act='tanh', bilinear_upsample=False,
                   **_unused):
    c = _Cursor(params, deterministic)
    skips = []
    for _ in _UNET_ENC:
        W, b = c.take(2)
        pre = c.bn(L.conv2d(x, W, b, 2, "same"))                # p2p.py:145-146 ...
        skips.append(pre)                                       # post-BN, pre-activation
        x = L.leaky_rectify(pre, 0.01)
    W, b = c.take(2)
    x = L.leaky_rectify(c.bn(L.conv2d(x, W, b, 1, "valid")), 0.01)      # conv9, p2p.py:193-195
    W, b = c.take(2)
    d = c.bn(L.deconv2d(x, W, b, 1))                            # dconv1, p2p.py:197-199
    x = L.leaky_rectify(torch.cat([d, skips[7]], 1), 0.01)      # p2p.py:202-203
    for j in range(len(_UNET_DEC)):
        W, b = c.take(2)
        if bilinear_upsample:
            d = L.conv2d(L.bilinear_upsample(x, 2), W, b, 1, "same")    # p2p.py:208-209
        else:
            d = L.deconv2d(x, W, b, 2)                          # p2p.py:206
        d = c.bn(d)
        x = L.leaky_rectify(torch.cat([d, skips[6 - j]], 1), 0.01)
    W, b = c.take(2)
    x = L.apply_nonlinearity(L.deconv2d(x, W, b, 2), act)       # p2p.py:272-275
    c.done()
    return x, c.updates


# --------------------------------------------------------------------------- #
# pix2pix PatchGAN discriminator  (architectures/p2p.py:278-292)
# --------------------------------------------------------------------------- #

def patch_discriminator_init(rng, in_shp, is_a_grayscale, is_b_grayscale, nf=32,
                             act='sigmoid', mul_factor=(1, 2, 4, 8), num_repeats=0, bn=False):
    assert num_repeats == 0
    params, meta = [], []
    cin = (1 if is_a_grayscale else 3) + (1 if is_b_grayscale else 3)
    for m in mul_factor:
        params += _conv_params(rng, nf * m, cin, 3)
        meta += ["W", "b"]
        if bn:
            params += _bn_params(nf * m)
            meta += ["beta", "gamma", "stat", "stat"]
        cin = nf * m
    params += _conv_params(rng, 1, cin, 3)
    meta += ["W", "b"]
    return params, meta


def patch_discriminator_forward(params, a, b_img, deterministic=False, act='sigmoid',
                                mul_factor=(1, 2, 4, 8), bn=False, **_unused):
    c = _Cursor(params, deterministic)
    x = torch.cat([a, b_img], 1)                                # p2p.py:281
    for _ in mul_factor:
        W, b = c.take(2)
        x = L.leaky_rectify(L.conv2d(x, W, b, 2, "same"), 0.01)  # p2p.py:285-286
        if bn:
            x = c.bn(x)                                         # p2p.py:288 (after the activation)
    W, b = c.take(2)
    x = L.apply_nonlinearity(L.conv2d(x, W, b, 2, "same"), act)  # p2p.py:289-290
    c.done()
    return x, c.updates
