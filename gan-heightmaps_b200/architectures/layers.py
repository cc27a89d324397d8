"""Custom layer of the reference (architectures/layers.py:13-26)."""
from lasagne_compat import Layer


class BilinearUpsample2DLayer(Layer):
    """2x upsampling with Theano's ``bilinear_upsampling`` semantics (even ratio:
    y[2m] = x[m], y[2m+1] = (x[m] + x[min(m+1, n-1)]) / 2 on each axis).  The
    engine never materialises it in parity mode: the consumer convolution gathers
    through it (HM_UP_BILINEAR2 in include/hmgan.h)."""

    def __init__(self, incoming, factor, **kwargs):
        super(BilinearUpsample2DLayer, self).__init__(incoming, **kwargs)
        if factor != 2:
            raise NotImplementedError("only factor 2 is used on the hot path (p2p.py:208)")
        self.factor = factor

    def get_output_shape_for(self, s):
        return s[:2] + (s[2] * self.factor, s[3] * self.factor)
