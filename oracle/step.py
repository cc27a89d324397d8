"""Torch-CPU restatement of Pix2Pix.__init__'s compiled functions
(pix2pix.py:87-147): train_fn, loss_fn, gen_fn(_det), z_fn(_det).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

Gradients come from torch autograd on the CPU restatement of the forward graph;
the update rule, its simultaneity (all four gradients at the OLD parameters, one
merged update, pix2pix.py:131-141) and the BatchNorm running-average side
effects of every non-deterministic function are restated explicitly.
"""
import numpy as np
import torch

from . import lasagne_ops as L
from . import networks as N

TRAIN_KEYS = ['dcgan_gen', 'dcgan_disc', 'p2p_gen', 'p2p_recon', 'p2p_disc']   # pix2pix.py:157


def _to_t(arrs, dtype):
    return [torch.tensor(np.asarray(a), dtype=dtype) for a in arrs]


class OracleModel(object):
    """State + the six callables.  ``nets`` maps 'G','D','P','Dp' to
    (params, meta, forward_kwargs); any of P/Dp (or G/D) may be absent, in which
    case its losses are reported as 0 (a new-repo extension for the 64-px
    DCGAN-only gate)."""

    def __init__(self, nets, alpha=100., opt='rmsprop', lr=1e-4, train_mode='both',
                 reconstruction='l1', lsgan=True, dtype=torch.float32):
        self.dtype = dtype
        self.kw = {}
        self.params = {}
        self.meta = {}
        for k, (p, m, kw) in nets.items():
            self.params[k] = _to_t(p, dtype)
            self.meta[k] = list(m)
            self.kw[k] = dict(kw)
        self.alpha, self.opt, self.lr = alpha, opt, lr
        self.train_mode, self.reconstruction, self.lsgan = train_mode, reconstruction, lsgan
        self.state = {k: [dict() for _ in v] for k, v in self.params.items()}
        self.last_grads = {}

    # ---- forwards -------------------------------------------------------- #
    def _fwd(self, key, params, *inputs, **extra):
        f = {'G': N.generator_forward, 'D': N.discriminator_forward,
             'P': N.g_unet_forward, 'Dp': N.patch_discriminator_forward}[key]
        kw = dict(self.kw[key])
        kw.update(extra)
        return f(params, *inputs, **kw)

    def _adv(self, out, target):
        if self.lsgan:
            return L.squared_error(out, target).mean()              # pix2pix.py:103
        t = torch.full_like(out, target)
        return L.binary_crossentropy(out, t).mean()                 # pix2pix.py:105

    def _graph(self, Z, X, Y, params):
        """pix2pix.py:91-121.  Returns the five losses, G's/P's outputs and the
        BN updates keyed by (net, index)."""
        zero = torch.zeros((), dtype=self.dtype)
        upd = {}
        out = {}
        losses = dict(gen_dcgan=zero, disc_dcgan=zero, gen_p2p=zero, recon=zero,
                      gen_total_p2p=zero, disc_p2p=zero)
        if 'G' in params:
            gz, u = self._fwd('G', params['G'], Z)                  # :92
            upd.update({('G', i): v for i, v in u.items()})
            d_real, _ = self._fwd('D', params['D'], X)              # :94
            d_fake, _ = self._fwd('D', params['D'], gz)             # :95
            losses['gen_dcgan'] = self._adv(d_fake, 1.)                                  # :107
            losses['disc_dcgan'] = self._adv(d_real, 1.) + self._adv(d_fake, 0.)         # :108
            out['gz'] = gz
        if 'P' in params:
            dp_real, _ = self._fwd('Dp', params['Dp'], X, Y)        # :98
            px, u = self._fwd('P', params['P'], X)                  # :99
            upd.update({('P', i): v for i, v in u.items()})
            dp_fake, _ = self._fwd('Dp', params['Dp'], X, px)       # :101
            losses['gen_p2p'] = self._adv(dp_fake, 1.)                                   # :110
            if self.reconstruction == 'l2':
                losses['recon'] = L.squared_error(px, Y).mean()                          # :113
            else:
                losses['recon'] = torch.abs(px - Y).mean()                               # :115
            losses['gen_total_p2p'] = losses['gen_p2p'] + self.alpha * losses['recon']   # :117
            losses['disc_p2p'] = self._adv(dp_real, 1.) + self._adv(dp_fake, 0.)         # :121
            out['px'] = px
        return losses, out, upd

    def _five(self, losses):
        return [np.float32(losses[k].item()) for k in
                ('gen_dcgan', 'disc_dcgan', 'gen_p2p', 'recon', 'disc_p2p')]

    def _apply_bn(self, upd):
        for (k, i), v in upd.items():
            self.params[k][i] = v.detach()

    # ---- compiled functions ----------------------------------------------- #
    def train_fn(self, Z, X, Y):
        Z, X, Y = [None if a is None else torch.tensor(a, dtype=self.dtype) for a in (Z, X, Y)]
        live = {k: [p.detach().clone().requires_grad_(m != "stat") for p, m in zip(ps, self.meta[k])]
                for k, ps in self.params.items()}
        losses, _, upd = self._graph(Z, X, Y, live)
        pairs = []
        if self.train_mode in ('both', 'dcgan') and 'G' in live:
            pairs += [('G', losses['gen_dcgan']), ('D', losses['disc_dcgan'])]
        if self.train_mode in ('both', 'p2p') and 'P' in live:
            pairs += [('P', losses['gen_total_p2p']), ('Dp', losses['disc_p2p'])]
        grads = {}
        for k, loss in pairs:                       # all gradients at the OLD parameters
            tr = [p for p in live[k] if p.requires_grad]
            g = torch.autograd.grad(loss, tr, retain_graph=True, allow_unused=True)
            grads[k] = [torch.zeros_like(p) if gi is None else gi for p, gi in zip(tr, g)]
        self.last_grads = {k: [g.numpy().copy() for g in v] for k, v in grads.items()}
        for k, g in grads.items():                  # one simultaneous update
            it = iter(g)
            for i, m in enumerate(self.meta[k]):
                if m == "stat":
                    continue
                self._update(k, i, next(it))
        self._apply_bn(upd)
        return self._five(losses)

    def _update(self, k, i, g):
        p, st = self.params[k][i], self.state[k][i]
        if self.opt == 'rmsprop':
            acc = st.get('acc', torch.zeros_like(p))
            p, acc = L.rmsprop_update(p, g, acc, self.lr)
            st['acc'] = acc
        elif self.opt == 'adam':
            m = st.get('m', torch.zeros_like(p))
            v = st.get('v', torch.zeros_like(p))
            t = st.get('t', 0)
            p, m, v, t = L.adam_update(p, g, m, v, t, self.lr)
            st.update(m=m, v=v, t=t)
        else:
            raise ValueError(self.opt)
        self.params[k][i] = p

    def loss_fn(self, Z, X, Y):
        Z, X, Y = [None if a is None else torch.tensor(a, dtype=self.dtype) for a in (Z, X, Y)]
        with torch.no_grad():
            losses, _, upd = self._graph(Z, X, Y, self.params)
        self._apply_bn(upd)                         # non-deterministic graph: stats still move
        return self._five(losses)

    def _gen(self, key, A, deterministic):
        A = torch.tensor(A, dtype=self.dtype)
        with torch.no_grad():
            out, upd = self._fwd(key, self.params[key], A, deterministic=deterministic)
        if not deterministic:
            self._apply_bn({(key, i): v for i, v in upd.items()})
        return out.numpy()

    def gen_fn(self, X):
        return self._gen('P', X, False)

    def gen_fn_det(self, X):
        return self._gen('P', X, True)

    def z_fn(self, Z):
        return self._gen('G', Z, False)

    def z_fn_det(self, Z):
        return self._gen('G', Z, True)

    def get_all_param_values(self, key):
        return [p.detach().numpy().astype(np.float32) for p in self.params[key]]


# --------------------------------------------------------------------------- #
# Standard configurations + synthetic inputs
# --------------------------------------------------------------------------- #

def experiment_kwargs(name):
    """Keyword sets of the experiments (experiments.py:102-119) and of the
    new-repo 64-px DCGAN-only gate."""
    if name == 'test1_nobn_bilin_both':
        return dict(
            in_shp=512, latent_dim=1000,
            G=dict(num_repeats=0, div=[2, 2, 4, 4, 8, 8, 8]),
            D=dict(num_repeats=0, bn=False, nonlinearity='linear', div=[8, 4, 4, 4, 2, 2, 2]),
            P=dict(nf=64, act='tanh', num_repeats=0, bilinear_upsample=True),
            Dp=dict(nf=64, bn=False, num_repeats=0, act='linear', mul_factor=[1, 2, 4, 8]))
    if name == 'gate64':
        return dict(
            in_shp=64, latent_dim=100,
            G=dict(nch=128, num_repeats=0, div=[2, 2, 4, 4]),
            D=dict(nch=64, num_repeats=0, bn=False, nonlinearity='linear', div=[8, 4, 2, 1]))
    if name == 'tiny512':
        # the full test1_nobn_bilin_both topology (all four networks, 512x512, 7+7+17+5 layers) at toy widths: what
        # the joint-step parity tests and tests/golden/joint_tiny512.npz run
        return dict(
            in_shp=512, latent_dim=16,
            G=dict(nch=64, num_repeats=0, div=[2, 2, 4, 4, 8, 8, 8]),
            D=dict(nch=512, num_repeats=0, bn=False, nonlinearity='linear', div=[128, 64, 64, 64, 32, 32, 32]),
            P=dict(nf=4, act='tanh', num_repeats=0, bilinear_upsample=True),
            Dp=dict(nf=4, bn=False, num_repeats=0, act='linear', mul_factor=[1, 2, 4, 8]))
    raise KeyError(name)


def build_nets(cfg, seed=2, which=('G', 'D', 'P', 'Dp')):
    rng = np.random.RandomState(seed)
    nets = {}
    if 'G' in which and 'G' in cfg:
        p, m = N.generator_init(rng, cfg['latent_dim'], True, **cfg['G'])
        nets['G'] = (p, m, cfg['G'])
        p, m = N.discriminator_init(rng, cfg['in_shp'], True, **cfg['D'])
        nets['D'] = (p, m, cfg['D'])
    if 'P' in which and 'P' in cfg:
        p, m = N.g_unet_init(rng, cfg['in_shp'], True, False, **cfg['P'])
        nets['P'] = (p, m, cfg['P'])
        p, m = N.patch_discriminator_init(rng, cfg['in_shp'], True, False, **cfg['Dp'])
        nets['Dp'] = (p, m, cfg['Dp'])
    return nets


def synthetic_batch(B, latent_dim, S, seed=0):
    """Z ~ U[0,1) (pix2pix.py:31,206); X a smooth heightmap-like field in [0,1];
    Y in [-1,1] (util.py:31-36 normalisations)."""
    Z = np.random.RandomState(seed).rand(B, latent_dim).astype(np.float32)
    r = np.random.RandomState(seed + 1)
    lo = max(S // 8, 1)
    base = r.rand(B, 1, lo, lo).astype(np.float32)
    X = np.kron(base, np.ones((1, 1, S // lo, S // lo), np.float32))
    X = 0.75 * X + 0.25 * r.rand(B, 1, S, S).astype(np.float32)
    X = (X ** 2).astype(np.float32)
    basey = r.rand(B, 3, lo, lo).astype(np.float32)
    Y = np.kron(basey, np.ones((1, 1, S // lo, S // lo), np.float32))
    Y = (1.5 * Y + 0.5 * r.rand(B, 3, S, S).astype(np.float32) - 1.0).clip(-1, 1).astype(np.float32)
    return Z, X, Y
