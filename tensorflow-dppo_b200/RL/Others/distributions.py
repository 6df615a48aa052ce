"""Probability-distribution surface of the reference (/root/reference/Others/distributions.py)
for the one family the DPPO update path uses: the diagonal Gaussian selected by make_pdtype for
Box action spaces (distributions.py:231-235).

Same class and method names (Pd / PdType / DiagGaussianPd / DiagGaussianPdType / CategoricalPd /
CategoricalPdType / make_pdtype).
The reference methods build TensorFlow graph ops; here a DiagGaussianPd wraps a [B,2A] device
tensor and every method is one launch of the sm_100a kernel behind dppo_diag_gaussian().  Inside
the training step these formulas are not called one by one: they are fused into the loss kernel.
"""
import itertools

_counter = itertools.count(1)


class Pd(object):
    """A particular probability distribution (distributions.py:8-26)."""
    def flatparam(self):
        raise NotImplementedError
    def mode(self):
        raise NotImplementedError
    def neglogp(self, x):
        raise NotImplementedError
    def kl(self, other):
        raise NotImplementedError
    def entropy(self):
        raise NotImplementedError
    def sample(self):
        raise NotImplementedError
    def logp(self, x):
        return - self.neglogp(x)                       # distributions.py:25-26


class PdType(object):
    """Parametrized family of probability distributions (distributions.py:28-46)."""
    def pdclass(self):
        raise NotImplementedError
    def pdfromflat(self, flat, engine=None):
        return self.pdclass()(flat, engine)
    def param_shape(self):
        raise NotImplementedError
    def sample_shape(self):
        raise NotImplementedError
    def sample_dtype(self):
        raise NotImplementedError


class DiagGaussianPdType(PdType):
    """distributions.py:77-87."""
    def __init__(self, size):
        self.size = size
    def pdclass(self):
        return DiagGaussianPd
    def param_shape(self):
        return [2 * self.size]
    def sample_shape(self):
        return [self.size]
    def sample_dtype(self):
        import torch
        return torch.float32


class DiagGaussianPd(Pd):
    """distributions.py:184-208 over a [B,2A] float32 CUDA tensor: flat = [mean | logstd]."""
    def __init__(self, flat, engine=None):
        if engine is None:
            raise RuntimeError("DiagGaussianPd needs the DppoEngine that owns the device (no CPU path)")
        self.flat = flat
        self.engine = engine
        A = flat.shape[-1] // 2
        self.mean = flat[..., :A]                       # :187 (views; the kernels read `flat`)
        self.logstd = flat[..., A:]
        self.seed = 0
    def flatparam(self):
        return self.flat
    def mode(self):
        return self.engine.diag_gaussian("mode", self.flat)
    def neglogp(self, x):
        return self.engine.diag_gaussian("neglogp", self.flat, x)
    def kl(self, other):
        assert isinstance(other, DiagGaussianPd)
        return self.engine.diag_gaussian("kl", self.flat, other.flat)
    def entropy(self):
        return self.engine.diag_gaussian("entropy", self.flat)
    def sample(self):
        return self.engine.diag_gaussian("sample", self.flat, seed=self.seed, offset=next(_counter))
    @classmethod
    def fromflat(cls, flat, engine=None):
        return cls(flat, engine)


class CategoricalPdType(PdType):
    """distributions.py:48-58."""
    def __init__(self, ncat):
        self.ncat = ncat
    def pdclass(self):
        return CategoricalPd
    def param_shape(self):
        return [self.ncat]
    def sample_shape(self):
        return []
    def sample_dtype(self):
        import torch
        return torch.int32


class CategoricalPd(Pd):
    """distributions.py:124-159 over a [B,ncat] float32 CUDA logits tensor; x / samples / mode are int32 [B]."""
    def __init__(self, logits, engine=None):
        if engine is None:
            raise RuntimeError("CategoricalPd needs the DppoEngine that owns the device (no CPU path)")
        self.logits = logits
        self.engine = engine
        self.seed = 0
    def flatparam(self):
        return self.logits
    def mode(self):
        return self.engine.categorical("mode", self.logits)
    def neglogp(self, x):
        return self.engine.categorical("neglogp", self.logits, x)
    def kl(self, other):
        assert isinstance(other, CategoricalPd)
        return self.engine.categorical("kl", self.logits, other.logits)
    def entropy(self):
        return self.engine.categorical("entropy", self.logits)
    def sample(self):
        return self.engine.categorical("sample", self.logits, seed=self.seed, offset=next(_counter))
    @classmethod
    def fromflat(cls, flat, engine=None):
        return cls(flat, engine)


def make_pdtype(ac_space):
    """distributions.py:231-243.  Box -> DiagGaussianPdType, Discrete -> CategoricalPdType; the MultiCategorical /
    Bernoulli families of the reference are outside the accelerated path."""
    if hasattr(ac_space, "n"):
        return CategoricalPdType(ac_space.n)
    shape = getattr(ac_space, "shape", None)
    if shape is not None:
        assert len(shape) == 1
        return DiagGaussianPdType(shape[0])
    raise NotImplementedError("only Box (DiagGaussianPd) and Discrete (CategoricalPd) action spaces are built")
