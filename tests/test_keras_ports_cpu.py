"""ReduceLROnPlateau (reference keras_ports.py:7-111) against the reference's own self-test (keras_ports.py:113-123:
lr 0.01, values 1.45, 1.43, 1.41 x10 -> one reduction, lr 0.001) and the behaviours its code implies."""
import os
import sys
import warnings

import numpy as np
import pytest

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
from keras_ports import ReduceLROnPlateau   # noqa: E402
from lasagne_compat import shared, floatX   # noqa: E402


def test_reference_self_test_sequence():
    lr = shared(0.01)
    cb = ReduceLROnPlateau(lr, verbose=0)
    cb.on_train_begin()
    cb.on_epoch_end(1.45, 1)
    cb.on_epoch_end(1.43, 2)
    cb.on_epoch_end(1.41, 3)
    seen = []
    for i in range(1, 10):
        cb.on_epoch_end(1.41, 3 + i)
        seen.append(float(lr.get_value()))
    # default mode 'auto' monitors for INCREASE: best stays 1.45; the rate drops at the 11th epoch without improvement
    np.testing.assert_allclose(seen[:-1], [0.01] * 8, rtol=1e-6)
    np.testing.assert_allclose(seen[-1], 0.001, rtol=1e-6)


def test_min_mode_cooldown_and_floor():
    lr = shared(floatX(1.0))
    cb = ReduceLROnPlateau(lr, factor=0.5, patience=1, mode='min', cooldown=2, min_lr=0.2)
    vals = []
    for e, v in enumerate([1.0, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]):
        cb.on_epoch_end(v, e)
        vals.append(round(float(lr.get_value()), 6))
    # improvements at epochs 0,1; wait reaches patience at epoch 3 -> 0.5; two cooldown epochs hold the counter at
    # zero; the next drops follow the same rhythm and stop at min_lr
    assert vals[:3] == [1.0, 1.0, 1.0] and vals[3] == 0.5
    assert vals[4] == 0.5 and vals[5] == 0.5
    assert min(vals) == 0.2 and vals[-1] == 0.2
    assert sorted(set(vals), reverse=True) == [1.0, 0.5, 0.25, 0.2]


def test_bad_arguments():
    with pytest.raises(ValueError):
        ReduceLROnPlateau(shared(0.1), factor=1.0)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        cb = ReduceLROnPlateau(shared(0.1), mode='sideways')
        assert cb.mode == 'auto' and any(issubclass(x.category, RuntimeWarning) for x in w)
        cb.on_epoch_end(None, 0)
        assert len(w) >= 2
