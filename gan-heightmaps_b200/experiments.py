"""Experiment registry / CLI with the reference's contract (reference experiments.py:20-131):

    python experiments.py <experiment_name> <mode>        mode in {train, interp, gen}

The closures and their keyword sets are those of the reference; `theano.shared`
and the Lasagne names it star-imports come from lasagne_compat.  `get_iterators`
(reference experiments.py:10-18) opens the HDF5 file with h5py when it is installed and
wraps its xt/yt/xv/yv datasets in util.Hdf5Iterator (flips and arbitrary-angle rotation with reflect
fill as augmentation, util.RotateFlipAugmenter; raw uint8 batches normalised on the device); set HMGAN_SYNTHETIC=<N> to train on N
seeded synthetic 512x512 pairs instead (what bench.py and the tests use).
"""
import os
import sys

import numpy as np

from pix2pix import Pix2Pix
from lasagne_compat import *          # noqa: F401,F403  (linear, tanh, rmsprop, adam, floatX, shared ...)
from lasagne_compat import floatX, shared, linear, tanh, rmsprop
from util import SyntheticIterator, Hdf5Iterator, FlipAugmenter, RotateFlipAugmenter


def get_iterators(dataset, batch_size, is_a_grayscale, is_b_grayscale, da=True):
    n_syn = int(os.environ.get("HMGAN_SYNTHETIC", "0"))
    if n_syn > 0:
        return SyntheticIterator(n_syn, batch_size, 512, 0), SyntheticIterator(n_syn, batch_size, 512, 1000)
    try:
        import h5py
    except ImportError:
        raise RuntimeError("h5py is not installed: cannot open %s.  Set HMGAN_SYNTHETIC=<N> to run on N synthetic "
                           "pairs (the HDF5 data path is out of this round's scope)." % dataset)
    f = h5py.File(dataset, "r")
    imgen = None
    if da:          # reference experiments.py:13 (flips + arbitrary rotation with reflect fill); HMGAN_DA=flip: flips only
        imgen = (FlipAugmenter(True, True) if os.environ.get("HMGAN_DA", "") == "flip"
                 else RotateFlipAugmenter(True, True, rotation_range=360, fill_mode="reflect"))
    raw = os.environ.get("HMGAN_DEVICE_NORMALISE", "1") != "0"     # raw uint8 batches, normalised on the device
    it_train = Hdf5Iterator(f['xt'], f['yt'], batch_size, imgen, is_a_grayscale=is_a_grayscale,
                            is_b_grayscale=is_b_grayscale, device_normalise=raw)
    it_val = Hdf5Iterator(f['xv'], f['yv'], batch_size, imgen, is_a_grayscale=is_a_grayscale,
                          is_b_grayscale=is_b_grayscale, device_normalise=raw)
    return it_train, it_val


def _model(gen_params_p2p, train_mode='both'):
    from architectures import p2p, dcgan
    return Pix2Pix(
        gen_fn_dcgan=dcgan.default_generator,
        disc_fn_dcgan=dcgan.default_discriminator,
        gen_params_dcgan={'num_repeats': 0, 'div': [2, 2, 4, 4, 8, 8, 8]},
        disc_params_dcgan={'num_repeats': 0, 'bn': False, 'nonlinearity': linear, 'div': [8, 4, 4, 4, 2, 2, 2]},
        gen_fn_p2p=p2p.g_unet,
        disc_fn_p2p=p2p.discriminator,
        gen_params_p2p=gen_params_p2p,
        disc_params_p2p={'nf': 64, 'bn': False, 'num_repeats': 0, 'act': linear, 'mul_factor': [1, 2, 4, 8]},
        in_shp=512,
        latent_dim=1000,
        is_a_grayscale=True,
        is_b_grayscale=False,
        lsgan=True,
        opt=rmsprop,
        opt_args={'learning_rate': shared(floatX(1e-4))},
        train_mode=train_mode)


DESERT_H5 = "/data/lisa/data/cbeckham/textures_v2_brown500.h5"


def test1_nobn(mode):
    assert mode in ["train", "interp", "gen"]
    model = _model({'nf': 64, 'act': tanh, 'num_repeats': 0})
    bs = 4
    name = "test1_repeatnod_fixp2p_nobn"
    if mode == "train":
        it_train, it_val = get_iterators(DESERT_H5, bs, True, False, True)
        model.train(it_train, it_val, batch_size=bs, num_epochs=1000, out_dir="output/%s" % name,
                    model_dir="models/%s" % name)
    elif mode == "interp":
        model.load_model("models/%s/600.model.bak" % name)
        zs = model.sampler(2, model.latent_dim)
        model.generate_interpolation("/tmp/test.png", floatX(zs[0]), floatX(zs[1]), mode='matrix')
    elif mode == "gen":
        model.load_model("models/%s/600.model.bak" % name)
        model.generate_gz(100, 10, "deleteme")


def test1_nobn_finetunep2p_bilin(mode):
    assert mode in ["train", "interp", "gen"]
    model = _model({'nf': 64, 'act': tanh, 'num_repeats': 0, 'bilinear_upsample': True}, train_mode='p2p')
    name = "test1_repeatnod_fixp2p_nobn_finetunep2p_bilin"
    bs = 4
    if mode == "train":
        model.load_model("models/test1_repeatnod_fixp2p_nobn/1000.model.bak", mode='dcgan')
        it_train, it_val = get_iterators(DESERT_H5, bs, True, False, True)
        model.train(it_train, it_val, batch_size=bs, num_epochs=1000, out_dir="output/%s" % name,
                    model_dir="models/%s" % name)
    elif mode == "interp":
        model.load_model("models/test1_repeatnod_fixp2p_nobn/1000.model.bak", mode='dcgan')
        model.load_model("models/%s/1000.model.bak" % name, mode='p2p')
        model.generate_interpolation_clip(100, 4, "output/%s/interp_clip_600_concat_bothdet/" % name, concat=True,
                                          deterministic=True)


def test1_nobn_bilin_both(mode):
    assert mode in ["train", "interp", "gen"]
    model = _model({'nf': 64, 'act': tanh, 'num_repeats': 0, 'bilinear_upsample': True}, train_mode='both')
    bs = 4
    name = "test1_nobn_bilin_both_deleteme"
    if mode == "train":
        it_train, it_val = get_iterators(DESERT_H5, bs, True, False, True)
        model.train(it_train, it_val, batch_size=bs, num_epochs=int(os.environ.get("HMGAN_EPOCHS", "1000")),
                    out_dir="output/%s" % name, model_dir="models/%s" % name,
                    quick_run=bool(int(os.environ.get("HMGAN_QUICK", "0"))))


if __name__ == '__main__':
    globals()[sys.argv[1]](sys.argv[2])
