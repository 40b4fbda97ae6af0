"""Empty stand-in: the reference imports matplotlib.pyplot but the hot path never plots."""
