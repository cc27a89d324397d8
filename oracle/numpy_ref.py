"""First-principles numpy (float64) restatement of the layer arithmetic, written
from the mathematical definitions rather than through torch.nn.functional, so
that a mis-remembered Lasagne/Theano convention in lasagne_ops.py shows up as a
disagreement between two independent restatements.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.  Small shapes only.
"""
import numpy as np


def conv2d(x, W, b, stride=1, pad="same"):
    """True convolution (Conv2DLayer, flip_filters=True):
    y[n,o,i,j] = b[o] + sum_{c,u,v} x_pad[n,c,i*s+u,j*s+v] * W[o,c,kh-1-u,kw-1-v]."""
    B, C, H, Wd = x.shape
    O, _, kh, kw = W.shape
    p = (kh // 2, kw // 2) if pad == "same" else (0, 0)
    xp = np.zeros((B, C, H + 2 * p[0], Wd + 2 * p[1]), np.float64)
    xp[:, :, p[0]:p[0] + H, p[1]:p[1] + Wd] = x
    Ho = (H + 2 * p[0] - kh) // stride + 1
    Wo = (Wd + 2 * p[1] - kw) // stride + 1
    y = np.zeros((B, O, Ho, Wo), np.float64)
    for u in range(kh):
        for v in range(kw):
            patch = xp[:, :, u:u + (Ho - 1) * stride + 1:stride, v:v + (Wo - 1) * stride + 1:stride]
            y += np.einsum('nchw,oc->nohw', patch, W[:, :, kh - 1 - u, kw - 1 - v])
    return y + b.reshape(1, -1, 1, 1)


def deconv2d(x, W, b, stride=2):
    """Deconv2DLayer(crop=0): the adjoint (input-gradient) of the true convolution
    with kernel W (Cin,Cout,kh,kw) read as (Cout_of_fwd=Cin, Cin_of_fwd=Cout).
    Forward conv: a[n,ci,i,j] = sum y[n,co,i*s+u,j*s+v] * W[ci,co,kh-1-u,kw-1-v];
    adjoint: y[n,co,i*s+u,j*s+v] += x[n,ci,i,j] * W[ci,co,kh-1-u,kw-1-v]."""
    B, Ci, H, Wd = x.shape
    _, Co, kh, kw = W.shape
    y = np.zeros((B, Co, (H - 1) * stride + kh, (Wd - 1) * stride + kw), np.float64)
    for u in range(kh):
        for v in range(kw):
            contrib = np.einsum('nchw,co->nohw', x, W[:, :, kh - 1 - u, kw - 1 - v])
            y[:, :, u:u + (H - 1) * stride + 1:stride, v:v + (Wd - 1) * stride + 1:stride] += contrib
    return y + b.reshape(1, -1, 1, 1)


def bilinear_upsample2(x):
    """Theano bilinear_upsampling, ratio 2:
    y[2m] = x[m], y[2m+1] = (x[m] + x[min(m+1,n-1)])/2 along each spatial axis."""
    def up(a, axis):
        n = a.shape[axis]
        nxt = np.take(a, np.minimum(np.arange(n) + 1, n - 1), axis=axis)
        out_shape = list(a.shape)
        out_shape[axis] = 2 * n
        out = np.zeros(out_shape, np.float64)
        ev = [slice(None)] * a.ndim
        od = [slice(None)] * a.ndim
        ev[axis] = slice(0, None, 2)
        od[axis] = slice(1, None, 2)
        out[tuple(ev)] = a
        out[tuple(od)] = 0.5 * (a + nxt)
        return out
    return up(up(x.astype(np.float64), 2), 3)


def batch_norm_train(x, beta, gamma, eps=1e-4):
    axes = tuple(i for i in range(x.ndim) if i != 1)
    sh = [1] * x.ndim
    sh[1] = -1
    m = x.mean(axes)
    v = ((x - m.reshape(sh)) ** 2).mean(axes)
    s = 1.0 / np.sqrt(v + eps)
    return (x - m.reshape(sh)) * (gamma * s).reshape(sh) + beta.reshape(sh), m, s


def max_pool2(x):
    B, C, H, W = x.shape
    return x[:, :, :H // 2 * 2, :W // 2 * 2].reshape(B, C, H // 2, 2, W // 2, 2).max((3, 5))


def upscale2(x):
    return np.repeat(np.repeat(x, 2, 2), 2, 3)
