"""Information-form Gaussian container (mirror of the reference's utils/gaussian.py:4-16)."""
import numpy as np


class NdimGaussian:
    def __init__(self, dimensionality, eta=None, lam=None):
        self.dim = dimensionality
        ok_eta = eta is not None and len(eta) == self.dim
        ok_lam = lam is not None and getattr(lam, "shape", None) == (self.dim, self.dim)
        self.eta = eta if ok_eta else np.zeros(self.dim)
        self.lam = lam if ok_lam else np.zeros([self.dim, self.dim])
