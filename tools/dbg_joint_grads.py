# which P/Dp gradient arrays deviate in fast mode at full width (GPU)
import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from oracle import step as S
from test_engine_cpu import build_pair
cfg = S.experiment_kwargs('test1_nobn_bilin_both')
om, m = build_pair(cfg, 'both', device="cuda", precision="fast")
Z, X, Y = S.synthetic_batch(2, cfg['latent_dim'], 512, seed=2)
lo = om.train_fn(Z, X, Y); lm = m.train_fn(Z, X, Y)
print(lo, lm)
sc = 1.0 / m.rt.loss_scale
for k, net in (('Dp', m.Dp), ('P', m.P)):
    tr = [q for q in net.params if q.trainable]
    for i, (a, b, q) in enumerate(zip(net.get_grads(), om.last_grads[k], tr)):
        if q.kind != "W": continue
        rel = float(np.linalg.norm((a * sc - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))
        print(k, i, q.shape, 'l2rel %.3f' % rel)
