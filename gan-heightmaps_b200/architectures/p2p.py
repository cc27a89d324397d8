"""pix2pix U-Net generator and PatchGAN discriminator factories.

Signatures, defaults, layer order (hence checkpoint parameter order) follow the
reference: ``g_unet`` architectures/p2p.py:126-276, ``discriminator`` :278-292,
helpers ``Convolution``/``Deconvolution``/``concatenate_layers`` :20-27.  The
nine-level encoder/decoder of the reference is written out longhand there; here
it is generated from the two channel tables below.
"""
from lasagne_compat import (InputLayer, BatchNormLayer, Conv2DLayer, Deconv2DLayer, NonlinearityLayer,
                            DropoutLayer, ConcatLayer, leaky_rectify, linear, sigmoid, tanh)
from .layers import BilinearUpsample2DLayer

ENCODER_MULT = (1, 2, 4, 8, 8, 8, 8, 8)     # conv1..conv8, each 3x3 stride 2 (512 -> 2)
DECODER_MULT = (8, 8, 8, 8, 4, 2, 1)        # dconv2..dconv8, each 2x up (2 -> 256)


def Convolution(layer, f, k=3, s=2, border_mode='same', **kwargs):
    return Conv2DLayer(layer, num_filters=f, filter_size=(k, k), stride=(s, s), pad=border_mode,
                       nonlinearity=linear)


def Deconvolution(layer, f, k=2, s=2, **kwargs):
    return Deconv2DLayer(layer, num_filters=f, filter_size=(k, k), stride=(s, s), nonlinearity=linear)


def concatenate_layers(layers, **kwargs):
    return ConcatLayer(layers, axis=1)


def g_unet(in_shp, is_a_grayscale, is_b_grayscale, nf=64, act=tanh, dropout=False, num_repeats=0,
           bilinear_upsample=False):
    """U-Net: 8 x [conv3x3 s2 -> BN -> LReLU], 2x2 valid bottleneck, then mirrored
    up path whose blocks are [bilinear 2x -> conv3x3] (or deconv 2x2 s2) -> BN ->
    concat(skip) -> LReLU; skips are the post-BN pre-activation encoder tensors."""
    assert in_shp in [512]

    def repeat_block(width, x):
        for _ in range(num_repeats):
            x = NonlinearityLayer(BatchNormLayer(Convolution(x, width, s=1, k=3)), nonlinearity=leaky_rectify)
        return x

    x = InputLayer((None, 1 if is_a_grayscale else 3, in_shp, in_shp))
    skips = []
    for mult in ENCODER_MULT:
        pre = BatchNormLayer(Convolution(x, nf * mult))
        skips.append(pre)
        x = repeat_block(nf * mult, NonlinearityLayer(pre, nonlinearity=leaky_rectify))
    bottleneck = BatchNormLayer(Convolution(x, nf * 8, k=2, s=1, border_mode='valid'))
    x = NonlinearityLayer(bottleneck, nonlinearity=leaky_rectify)
    up = BatchNormLayer(Deconvolution(x, nf * 8, k=2, s=1))
    if dropout:
        up = DropoutLayer(up, p=0.5)
    x = NonlinearityLayer(concatenate_layers([up, skips[7]]), nonlinearity=leaky_rectify)
    for level, mult in enumerate(DECODER_MULT):
        if bilinear_upsample:
            up = Convolution(BilinearUpsample2DLayer(x, 2), nf * mult, s=1)
        else:
            up = Deconvolution(x, nf * mult)
        up = BatchNormLayer(up)
        if dropout and level < 2:
            up = DropoutLayer(up, p=0.5)
        x = NonlinearityLayer(concatenate_layers([up, skips[6 - level]]), leaky_rectify)
    out = Deconvolution(x, 1 if is_b_grayscale else 3)
    return NonlinearityLayer(out, act)


def discriminator(in_shp, is_a_grayscale, is_b_grayscale, nf=32, act=sigmoid, mul_factor=[1, 2, 4, 8],
                  num_repeats=0, bn=False):
    """PatchGAN on concat(A, B): [conv3x3 (s2 first) -> LReLU (-> BN)] per multiplier,
    then conv3x3 s2 -> 1 channel -> act.  Returns {"inputs": [A, B], "out": layer}."""
    i_a = InputLayer((None, 1 if is_a_grayscale else 3, in_shp, in_shp))
    i_b = InputLayer((None, 1 if is_b_grayscale else 3, in_shp, in_shp))
    x = concatenate_layers([i_a, i_b])
    for m in mul_factor:
        for r in range(num_repeats + 1):
            x = NonlinearityLayer(Convolution(x, nf * m, s=2 if r == 0 else 1), leaky_rectify)
            if bn:
                x = BatchNormLayer(x)
    out = NonlinearityLayer(Convolution(x, 1), act)
    return {"inputs": [i_a, i_b], "out": out}
