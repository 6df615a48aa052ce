"""Eager numpy stand-in for the handful of TensorFlow-1.x ops the reference's hot path calls.

TEST INFRASTRUCTURE (oracle side).  Purpose: let ``oracle/make_golden.py`` execute the
reference's OWN source (Others/distributions.py, Others/tf_util.py, Model.py and the loss
lines of PPO.py) in this container, where TensorFlow is not installed, so that golden vectors
come from the reference's composition of ops rather than from our restatement of it.  Each op
below restates the documented TF-1.x semantics of the op of the same name; nothing here is
imported by the product.
"""
import types
import numpy as np

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
uint8 = np.uint8
_DEFAULT = np.float64          # goldens are generated in fp64 unless the caller feeds fp32


class _Var(np.ndarray):
    """trainable variable: an ndarray that remembers its scope name"""
    pass


class Tensor(np.ndarray):
    """ndarray with the static-shape accessor CategoricalPd.neglogp uses (distributions.py:135)."""
    def get_shape(self):
        shp = self.shape
        return types.SimpleNamespace(as_list=lambda: list(shp))


def as_tensor(x, dtype=None):
    return np.asarray(x, dtype).view(Tensor)


_state = types.SimpleNamespace(scope=[], variables=[])


def reset_default_graph():
    _state.scope = []
    _state.variables = []


class variable_scope:
    def __init__(self, name, reuse=None):
        self.name = name
    def __enter__(self):
        _state.scope.append(self.name)
        return self
    def __exit__(self, *a):
        _state.scope.pop()


class GraphKeys:
    TRAINABLE_VARIABLES = "trainable_variables"
    GLOBAL_VARIABLES = "variables"


def get_collection(key, scope=None):
    return [v for (s, v) in _state.variables if scope is None or s.startswith(scope)]


def constant(x, dtype=None):
    return np.asarray(x, dtype)


def expand_dims(x, axis):
    return np.expand_dims(x, axis)


def split(axis, num_or_size_splits, value):
    return np.split(value, num_or_size_splits, axis=axis)


def exp(x): return np.exp(x)
def log(x): return np.log(x)
def square(x): return np.square(x)
def sqrt(x): return np.sqrt(x)
def minimum(x, y): return np.minimum(x, y)
def maximum(x, y): return np.maximum(x, y)
def clip_by_value(x, lo, hi): return np.minimum(np.maximum(x, lo), hi)
def to_float(x): return np.asarray(x, np.float64)
def shape(x): return np.asarray(np.shape(x))
def concat(values=None, axis=0, **kw): return np.concatenate(values, axis=axis)


def reduce_sum(x, axis=None, keep_dims=False):
    ax = None if axis is None else tuple(np.atleast_1d(axis).tolist())
    return np.sum(x, axis=ax, keepdims=keep_dims)


def reduce_mean(x, axis=None, keep_dims=False):
    ax = None if axis is None else tuple(np.atleast_1d(axis).tolist())
    return np.mean(x, axis=ax, keepdims=keep_dims)


_rng = np.random.default_rng(0)


def random_normal(shape_, dtype=None):
    return _rng.standard_normal(tuple(int(s) for s in shape_))


def reduce_max(x, axis=None, keep_dims=False):
    if isinstance(axis, (list, tuple)):
        axis = tuple(axis)
    return np.max(x, axis=axis, keepdims=keep_dims)


def argmax(x, axis=None): return np.argmax(x, axis=axis)


def one_hot(x, depth):
    return (np.asarray(x)[..., None] == np.arange(depth)).astype(np.float64)


def random_uniform(shape_, dtype=None):
    return _rng.random(tuple(int(s) for s in shape_))


class _nn:
    @staticmethod
    def relu(x): return np.maximum(x, 0)

    @staticmethod
    def softmax_cross_entropy_with_logits(logits=None, labels=None):
        """TF op semantics (restated): -sum(labels * log_softmax(logits)) over the last axis, max-shifted."""
        z = logits - np.max(logits, axis=-1, keepdims=True)
        lsm = z - np.log(np.sum(np.exp(z), axis=-1, keepdims=True))
        return -np.sum(labels * lsm, axis=-1)


nn = _nn


class _layers:
    @staticmethod
    def dense(x, units, activation=None, kernel_initializer=None, name=None):
        """tf.layers.dense: kernel [in,out] via initializer, zero bias, tensordot on last axis."""
        fan_in = x.shape[-1]
        idx = sum(1 for (s, _) in _state.variables if s.startswith("/".join(_state.scope))) // 2
        sc = "/".join(_state.scope) + ("/dense" if idx == 0 else "/dense_%d" % idx)
        k = np.asarray(kernel_initializer((fan_in, units))).astype(x.dtype)
        b = np.zeros(units, x.dtype)
        _state.variables.append((sc + "/kernel", k))
        _state.variables.append((sc + "/bias", b))
        y = np.tensordot(x, k, axes=[[x.ndim - 1], [0]]) + b
        return activation(y) if activation is not None else y


layers = _layers


def install():
    """Register this module as ``tensorflow`` (+ the one submodule distributions.py imports)."""
    import sys
    me = sys.modules[__name__]
    sys.modules["tensorflow"] = me
    py = types.ModuleType("tensorflow.python"); ops = types.ModuleType("tensorflow.python.ops")
    mo = types.ModuleType("tensorflow.python.ops.math_ops")
    mo.less = np.less
    ops.math_ops = mo; py.ops = ops
    sys.modules["tensorflow.python"] = py
    sys.modules["tensorflow.python.ops"] = ops
    sys.modules["tensorflow.python.ops.math_ops"] = mo
    return me
