"""Network factories with the reference's signatures (architectures/dcgan.py,
architectures/p2p.py, architectures/layers.py)."""
