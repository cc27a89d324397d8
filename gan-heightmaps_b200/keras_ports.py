"""Learning-rate schedule helper with the reference's surface (reference keras_ports.py:7-111).

The reference carries a Lasagne-side ``ReduceLROnPlateau`` that multiplies a shared learning rate by ``factor`` once a
monitored value has not improved for ``patience`` epochs.  It is dormant upstream (its only call site is commented
out, pix2pix.py:242); it is kept here because ``Pix2Pix.lr`` is a shared scalar precisely so that such a callback can
drive it: the new value reaches the kernels through the device-side learning rate the captured CUDA graphs read.

Behaviour kept from the reference, quirks included:
  * ``mode='min'`` improves when ``value < best - epsilon``; every other accepted mode -- 'max' AND the default 'auto'
    -- improves when ``value > best + epsilon`` (keras_ports.py:70-75: 'auto' is not inferred from a metric name);
  * an unknown mode falls back to 'auto' with a RuntimeWarning;
  * the patience test comes before the increment, so the rate drops on the (patience+1)-th epoch without improvement;
  * after a drop, ``cooldown`` epochs pass during which the wait counter is held at zero;
  * the rate is only lowered while it is above ``min_lr`` (by more than min_lr*1e-4) and never below ``min_lr``.
"""
import warnings

import numpy as np


class ReduceLROnPlateau(object):
    def __init__(self, learning_rate, factor=0.1, patience=10, verbose=0, mode='auto', epsilon=1e-4, cooldown=0,
                 min_lr=0):
        if factor >= 1.0:
            raise ValueError('ReduceLROnPlateau does not support a factor >= 1.0.')
        self.learning_rate = learning_rate          # shared scalar: get_value() / set_value()
        self.factor, self.patience, self.verbose = factor, patience, verbose
        self.mode, self.epsilon, self.cooldown, self.min_lr = mode, epsilon, cooldown, min_lr
        self._reset()

    def _reset(self):
        if self.mode not in ('auto', 'min', 'max'):
            warnings.warn('Learning Rate Plateau Reducing mode %s is unknown, fallback to auto mode.' % self.mode,
                          RuntimeWarning)
            self.mode = 'auto'
        if self.mode == 'min':
            self.monitor_op = lambda value, best: bool(np.less(value, best - self.epsilon))
            self.best = np.inf
        else:
            self.monitor_op = lambda value, best: bool(np.greater(value, best + self.epsilon))
            self.best = -np.inf
        self.cooldown_counter = 0
        self.wait = 0
        self.lr_epsilon = self.min_lr * 1e-4

    def on_train_begin(self, logs=None):
        self._reset()

    def in_cooldown(self):
        return self.cooldown_counter > 0

    def on_epoch_end(self, monitor, epoch, logs=None):
        if monitor is None:
            warnings.warn('Learning Rate Plateau Reducing requires a monitored value', RuntimeWarning)
            return
        if self.in_cooldown():
            self.cooldown_counter -= 1
            self.wait = 0
        if self.monitor_op(monitor, self.best):
            self.best = monitor
            self.wait = 0
            return
        if self.in_cooldown():
            return
        if self.wait >= self.patience:
            old_lr = float(self.learning_rate.get_value())
            if old_lr > self.min_lr + self.lr_epsilon:
                new_lr = max(old_lr * self.factor, self.min_lr)
                self.learning_rate.set_value(new_lr)
                if self.verbose > 0:
                    print('\nEpoch %05d: reducing learning rate to %s.' % (epoch, new_lr))
                self.cooldown_counter = self.cooldown
                self.wait = 0
        self.wait += 1
