"""CPU oracle for the gan-heightmaps training step.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic that the reference
(christopher-beckham/gan-heightmaps) delegates to Theano + Lasagne 0.2.dev1.
It exists to CHECK the CUDA path; it is never on the product path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.

PARITY UNPINNED: the reference cannot run here (Python-2-only, needs
Theano/Lasagne/Keras; none installed, no network) and ships no golden tensors,
seeds or saved weights.  The only known answers it records are
structural (parameter counts 22 882 243 / 391 009 and layer shape listings in
``g_unet.ipynb``); those are checked in ``tests/test_oracle.py``.  Numerical
fixtures under ``tests/golden/`` are produced BY this oracle
(``oracle/make_golden.py``), cross-checked against an independent
first-principles numpy restatement (``oracle/numpy_ref.py``).

Third-party algorithm sources restated here (not vendored in /root/reference):
  * Lasagne 0.2.dev1 (version evidence: lasagne/notebooks/gaussian_blur.ipynb:600)
    - layers.Conv2DLayer / TransposedConv2DLayer / DenseLayer / BatchNormLayer /
      Pool2DLayer / Upscale2DLayer, updates.rmsprop / adam, init.GlorotUniform,
      objectives.squared_error / binary_crossentropy.
  * Theano >= 0.9 tensor.nnet.abstract_conv.bilinear_upsampling.
"""
