"""DCGAN heightmap generator / discriminator factories.

Signatures, defaults and layer order are those of the reference
(architectures/dcgan.py:14 ``default_generator`` and :35 ``default_discriminator``)
so that ``Pix2Pix(gen_fn_dcgan=dcgan.default_generator, ...)`` drops in; the
returned objects are layer specifications (lasagne_compat), lowered to sm_100a
kernel programs by engine.lower.
"""
from lasagne_compat import (InputLayer, DenseLayer, BatchNormLayer, ReshapeLayer, Conv2DLayer,
                            NonlinearityLayer, DropoutLayer, Upscale2DLayer, MaxPool2DLayer,
                            Pool2DLayer, LeakyRectify, linear, sigmoid)
from .layers import BilinearUpsample2DLayer


def _widths(nch, div):
    # the reference divides with Python-2 integer semantics (dcgan.py:19,39)
    return [nch // d for d in div]


def default_generator(latent_dim, is_a_grayscale, nch=512, h=5, initial_size=4, final_size=512,
                      div=[2, 2, 4, 4, 8, 8, 16], num_repeats=0, dropout_p=0.,
                      bilinear_upsample=False):
    """z -> dense -> BN -> (nch, s0, s0) -> [conv hxh -> BN -> LReLU(0.2)]*(r+1) -> 2x up ... -> conv -> sigmoid.
    ``final_size`` is accepted and ignored, as in the reference."""
    net = InputLayer((None, latent_dim))
    net = BatchNormLayer(DenseLayer(net, num_units=nch * initial_size * initial_size, nonlinearity=linear))
    net = ReshapeLayer(net, (-1, nch, initial_size, initial_size))
    for width in _widths(nch, div):
        for _ in range(num_repeats + 1):
            net = BatchNormLayer(Conv2DLayer(net, num_filters=width, filter_size=h, pad='same',
                                             nonlinearity=linear))
            net = NonlinearityLayer(net, nonlinearity=LeakyRectify(0.2))
            if dropout_p > 0.:
                net = DropoutLayer(net, p=dropout_p)
        net = BilinearUpsample2DLayer(net, factor=2) if bilinear_upsample else Upscale2DLayer(net, scale_factor=2)
    out_ch = 1 if is_a_grayscale else 3
    return Conv2DLayer(net, num_filters=out_ch, filter_size=h, pad='same', nonlinearity=sigmoid)


def default_discriminator(in_shp, is_a_grayscale, nch=512, h=5, div=[8, 4, 4, 2, 2, 1, 1], num_repeats=0,
                          bn=False, pool_mode='max', nonlinearity='sigmoid'):
    """x -> [conv hxh (-> BN) -> LReLU(0.2)]*(r+1) -> pool2 ... -> conv(1)+rectify -> avg-pool -> (-1,1) -> nonlinearity.
    The average-pool size is derived from ``nch`` (not ``in_shp``), as in the
    reference (dcgan.py:51): the network is well formed only for nch == in_shp."""
    net = InputLayer((None, 1 if is_a_grayscale else 3, in_shp, in_shp))
    widths = _widths(nch, div)
    for width in widths:
        for _ in range(num_repeats + 1):
            net = Conv2DLayer(net, num_filters=width, filter_size=h, pad='same', nonlinearity=linear)
            if bn:
                net = BatchNormLayer(net)
            net = NonlinearityLayer(net, nonlinearity=LeakyRectify(0.2))
        if pool_mode == 'max':
            net = MaxPool2DLayer(net, pool_size=2)
        else:
            net = Pool2DLayer(net, pool_size=2, mode='average_inc_pad')
    net = Conv2DLayer(net, num_filters=1, filter_size=h, pad='same')      # default nonlinearity: rectify
    rf = nch // (2 ** len(widths))
    net = Pool2DLayer(net, pool_size=(rf, rf), mode='average_inc_pad')
    net = ReshapeLayer(net, (-1, 1))
    return NonlinearityLayer(net, nonlinearity)
