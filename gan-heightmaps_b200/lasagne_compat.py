"""Layer-spec vocabulary with Lasagne's names.

The reference builds its networks out of ``lasagne.layers`` objects and hands the
last layer to ``Pix2Pix`` (reference pix2pix.py:73-77).  Here the same
constructors build a light, framework-free description of the graph; nothing is
computed until ``engine.lower`` turns the description into a program of
sm_100a kernel launches.  Only what the four hot-path factories use is provided
(SURVEY.md §8a); names, argument meaning and defaults follow Lasagne 0.2.dev1.
"""
import math

import numpy as np


# --------------------------------------------------------------------------- #
# nonlinearities (lasagne.nonlinearities) — identity objects, never evaluated on
# the host; the kernels implement them (include/hmgan.h HmAct).
# --------------------------------------------------------------------------- #
class _Nonlinearity(object):
    def __init__(self, name, slope=0.0):
        self.name, self.slope = name, float(slope)

    def __repr__(self):
        return "<%s%s>" % (self.name, "(%g)" % self.slope if self.name == "leaky_rectify" else "")

    def __call__(self, x):
        raise TypeError("nonlinearities are symbolic in this framework; %r is evaluated "
                        "by the CUDA kernels" % self)


linear = identity = _Nonlinearity("linear")
rectify = _Nonlinearity("rectify")
sigmoid = _Nonlinearity("sigmoid")
tanh = _Nonlinearity("tanh")


def LeakyRectify(leakiness=0.01):
    return _Nonlinearity("leaky_rectify", leakiness)


leaky_rectify = LeakyRectify(0.01)
very_leaky_rectify = LeakyRectify(1. / 3)


def as_nonlinearity(f):
    if f is None:
        return linear
    if isinstance(f, _Nonlinearity):
        return f
    if isinstance(f, str):   # dcgan.default_discriminator's default is the STRING 'sigmoid'
        return {"linear": linear, "sigmoid": sigmoid, "tanh": tanh, "rectify": rectify}[f]
    raise TypeError("unsupported nonlinearity %r" % (f,))


# --------------------------------------------------------------------------- #
# lasagne.utils / theano.shared stand-ins used by experiments.py:116-117
# --------------------------------------------------------------------------- #
def floatX(x):
    return np.asarray(x, dtype=np.float32)


class SharedScalar(object):
    """theano.shared(floatX(v)) as used for the learning rate (pix2pix.py:156,259)."""

    def __init__(self, value):
        self._v = np.float32(value)
        self._listeners = []

    def get_value(self):
        return self._v

    def set_value(self, v):
        self._v = np.float32(v)
        for f in self._listeners:
            f(self._v)


def shared(value):
    return SharedScalar(value)


# --------------------------------------------------------------------------- #
# lasagne.updates — the optimiser is passed around as a callable
# (pix2pix.py:30,132); here it is a tag the engine dispatches on.
# --------------------------------------------------------------------------- #
class _Optimiser(object):
    def __init__(self, name, **defaults):
        self.name, self.defaults = name, defaults

    def __repr__(self):
        return "<lasagne.updates.%s>" % self.name


rmsprop = _Optimiser("rmsprop", learning_rate=1.0, rho=0.9, epsilon=1e-6)
adam = _Optimiser("adam", learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8)


# --------------------------------------------------------------------------- #
# layers
# --------------------------------------------------------------------------- #
class Param(object):
    def __init__(self, name, shape, kind, trainable=True):
        self.name, self.shape, self.kind, self.trainable = name, tuple(shape), kind, trainable

    @property
    def size(self):
        return int(np.prod(self.shape))


class Layer(object):
    def __init__(self, incoming, name=None):
        self.input_layer = incoming
        self.input_shape = incoming.output_shape if incoming is not None else None
        self.params = []
        self.name = name

    @property
    def output_shape(self):
        return self.get_output_shape_for(self.input_shape)

    def get_output_shape_for(self, s):
        return s

    def inputs(self):
        return [] if self.input_layer is None else [self.input_layer]

    def __repr__(self):
        return "<%s>" % type(self).__name__


class InputLayer(Layer):
    def __init__(self, shape, **kw):
        super(InputLayer, self).__init__(None, **kw)
        self.shape = tuple(shape)

    @property
    def output_shape(self):
        return self.shape


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class DenseLayer(Layer):
    def __init__(self, incoming, num_units, nonlinearity=rectify, **kw):
        super(DenseLayer, self).__init__(incoming, **kw)
        self.num_units = int(num_units)
        self.nonlinearity = as_nonlinearity(nonlinearity)
        n_in = int(np.prod(self.input_shape[1:]))
        self.params = [Param("W", (n_in, self.num_units), "W"), Param("b", (self.num_units,), "b")]

    def get_output_shape_for(self, s):
        return (s[0], self.num_units)


class BatchNormLayer(Layer):
    """axes='auto', epsilon=1e-4, alpha=0.1; params beta, gamma, mean, inv_std."""

    def __init__(self, incoming, epsilon=1e-4, alpha=0.1, **kw):
        super(BatchNormLayer, self).__init__(incoming, **kw)
        self.epsilon, self.alpha = float(epsilon), float(alpha)
        c = self.input_shape[1]
        self.params = [Param("beta", (c,), "beta"), Param("gamma", (c,), "gamma"),
                       Param("mean", (c,), "mean", False), Param("inv_std", (c,), "inv_std", False)]


class ReshapeLayer(Layer):
    def __init__(self, incoming, shape, **kw):
        super(ReshapeLayer, self).__init__(incoming, **kw)
        self.shape = tuple(shape)

    def get_output_shape_for(self, s):
        known = int(np.prod([d for d in self.shape if d != -1]))
        if s[0] is None:
            return (None,) + tuple(d for d in self.shape[1:])
        total = int(np.prod(s))
        return tuple(total // known if d == -1 else d for d in self.shape)


class NonlinearityLayer(Layer):
    def __init__(self, incoming, nonlinearity=rectify, **kw):
        super(NonlinearityLayer, self).__init__(incoming, **kw)
        self.nonlinearity = as_nonlinearity(nonlinearity)


class DropoutLayer(Layer):
    def __init__(self, incoming, p=0.5, **kw):
        super(DropoutLayer, self).__init__(incoming, **kw)
        self.p = p


class Conv2DLayer(Layer):
    """flip_filters=True (true convolution), W (num_filters, Cin, kh, kw), b=0."""

    def __init__(self, incoming, num_filters, filter_size, stride=(1, 1), pad=0,
                 nonlinearity=rectify, **kw):
        super(Conv2DLayer, self).__init__(incoming, **kw)
        self.num_filters = int(num_filters)
        self.filter_size = _pair(filter_size)
        self.stride = _pair(stride)
        if pad == 'same':
            if self.filter_size[0] % 2 == 0:
                raise NotImplementedError("`same` padding requires odd filter size.")
            self.pad = (self.filter_size[0] // 2, self.filter_size[1] // 2)
        elif pad == 'valid':
            self.pad = (0, 0)
        else:
            self.pad = _pair(pad)
        self.nonlinearity = as_nonlinearity(nonlinearity)
        cin = self.input_shape[1]
        self.params = [Param("W", (self.num_filters, cin) + self.filter_size, "W"),
                       Param("b", (self.num_filters,), "b")]

    def get_output_shape_for(self, s):
        hw = tuple((s[2 + i] + 2 * self.pad[i] - self.filter_size[i]) // self.stride[i] + 1
                   for i in range(2))
        return (s[0], self.num_filters) + hw


class TransposedConv2DLayer(Layer):
    """crop=0, flip_filters=False; W (Cin, num_filters, kh, kw)."""

    def __init__(self, incoming, num_filters, filter_size, stride=(1, 1), crop=0,
                 nonlinearity=rectify, **kw):
        super(TransposedConv2DLayer, self).__init__(incoming, **kw)
        self.num_filters = int(num_filters)
        self.filter_size = _pair(filter_size)
        self.stride = _pair(stride)
        if crop not in (0, 'valid'):
            raise NotImplementedError("only crop=0 is used on the hot path")
        self.nonlinearity = as_nonlinearity(nonlinearity)
        cin = self.input_shape[1]
        self.params = [Param("W", (cin, self.num_filters) + self.filter_size, "W"),
                       Param("b", (self.num_filters,), "b")]

    def get_output_shape_for(self, s):
        hw = tuple((s[2 + i] - 1) * self.stride[i] + self.filter_size[i] for i in range(2))
        return (s[0], self.num_filters) + hw


Deconv2DLayer = TransposedConv2DLayer


class Upscale2DLayer(Layer):
    def __init__(self, incoming, scale_factor, mode='repeat', **kw):
        super(Upscale2DLayer, self).__init__(incoming, **kw)
        self.scale_factor = _pair(scale_factor)
        if self.scale_factor != (2, 2) or mode != 'repeat':
            raise NotImplementedError("only 2x repeat upscaling is used on the hot path")

    def get_output_shape_for(self, s):
        return s[:2] + (s[2] * 2, s[3] * 2)


class Pool2DLayer(Layer):
    def __init__(self, incoming, pool_size, stride=None, pad=(0, 0), ignore_border=True,
                 mode='max', **kw):
        super(Pool2DLayer, self).__init__(incoming, **kw)
        self.pool_size = _pair(pool_size)
        self.stride = self.pool_size if stride is None else _pair(stride)
        self.mode = mode
        if self.stride != self.pool_size or _pair(pad) != (0, 0):
            raise NotImplementedError("only non-overlapping unpadded pooling is used on the hot path")

    def get_output_shape_for(self, s):
        return s[:2] + (s[2] // self.pool_size[0], s[3] // self.pool_size[1])


class MaxPool2DLayer(Pool2DLayer):
    def __init__(self, incoming, pool_size, **kw):
        super(MaxPool2DLayer, self).__init__(incoming, pool_size, mode='max', **kw)


class ConcatLayer(Layer):
    def __init__(self, incomings, axis=1, **kw):
        super(ConcatLayer, self).__init__(incomings[0], **kw)
        self.input_layers = list(incomings)
        if axis != 1:
            raise NotImplementedError("only channel concatenation is used on the hot path")

    @property
    def output_shape(self):
        shapes = [l.output_shape for l in self.input_layers]
        return (shapes[0][0], sum(s[1] for s in shapes)) + tuple(shapes[0][2:])

    def inputs(self):
        return list(self.input_layers)


# --------------------------------------------------------------------------- #
# lasagne.layers helpers
# --------------------------------------------------------------------------- #
def get_all_layers(layer):
    """Topological order, depth-first through the inputs in declaration order
    (lasagne.layers.get_all_layers) — this fixes the checkpoint parameter order."""
    out, seen = [], set()

    def visit(l):
        if id(l) in seen:
            return
        seen.add(id(l))
        for i in l.inputs():
            visit(i)
        out.append(l)

    visit(layer)
    return out


def get_all_params(layer, trainable=None):
    out = []
    for l in get_all_layers(layer):
        for p in l.params:
            if trainable is None or p.trainable == trainable:
                out.append(p)
    return out


def count_params(layer, trainable=None):
    return sum(p.size for p in get_all_params(layer, trainable))


def glorot_uniform(rng, shape):
    """lasagne.init.GlorotUniform(gain=1) from an explicit RandomState."""
    n1, n2 = shape[0], shape[1]
    rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    a = math.sqrt(6.0 / ((n1 + n2) * rf))
    return rng.uniform(-a, a, size=shape).astype(np.float32)


def init_param(rng, p):
    if p.kind == "W":
        return glorot_uniform(rng, p.shape)
    if p.kind in ("gamma", "inv_std"):
        return np.ones(p.shape, np.float32)
    return np.zeros(p.shape, np.float32)
