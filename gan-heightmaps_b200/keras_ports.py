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
        if not factor < 1.0:
            raise ValueError('ReduceLROnPlateau needs a factor < 1.0, got %r' % (factor,))
        self.learning_rate = learning_rate          # shared scalar: get_value() / set_value()
        self.factor, self.patience, self.verbose = factor, patience, verbose
        self.mode, self.epsilon, self.cooldown, self.min_lr = mode, epsilon, cooldown, min_lr
        self._reset()

    # -- state ------------------------------------------------------------- #
    def _reset(self):
        """Fresh counters; validates the mode (unknown -> 'auto' with a RuntimeWarning)."""
        if self.mode not in ('auto', 'min', 'max'):
            warnings.warn('ReduceLROnPlateau: unknown mode %r, using auto' % (self.mode,), RuntimeWarning)
            self.mode = 'auto'
        self._minimise = self.mode == 'min'          # 'auto' behaves as 'max' (see the module docstring)
        self.best = np.inf if self._minimise else -np.inf
        self.wait = 0
        self.cooldown_counter = 0
        self.lr_epsilon = 1e-4 * self.min_lr

    def monitor_op(self, value, best):
        """True if `value` improves on `best` by more than epsilon in the monitored direction."""
        return bool(value < best - self.epsilon) if self._minimise else bool(value > best + self.epsilon)

    def in_cooldown(self):
        return self.cooldown_counter > 0

    # -- callbacks --------------------------------------------------------- #
    def on_train_begin(self, logs=None):
        self._reset()

    def on_epoch_end(self, monitor, epoch, logs=None):
        if monitor is None:
            warnings.warn('ReduceLROnPlateau: no monitored value given', RuntimeWarning)
            return
        cooling = self.in_cooldown()
        if cooling:
            self.cooldown_counter -= 1
            self.wait = 0
            cooling = self.in_cooldown()
        if self.monitor_op(monitor, self.best):
            self.best, self.wait = monitor, 0
        elif not cooling:
            if self.wait >= self.patience:
                self._reduce(epoch)
            self.wait += 1

    def _reduce(self, epoch):
        lr = float(self.learning_rate.get_value())
        if lr <= self.min_lr + self.lr_epsilon:
            return
        lr = max(lr * self.factor, self.min_lr)
        self.learning_rate.set_value(lr)
        if self.verbose > 0:
            print('\nEpoch %05d: reducing learning rate to %s.' % (epoch, lr))
        self.cooldown_counter = self.cooldown
        self.wait = 0
