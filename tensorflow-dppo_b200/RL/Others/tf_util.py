"""The two tf_util helpers the update path touches (/root/reference/Others/tf_util.py:14-24,
286-291).  sum/mean/clip are graph-op aliases there; here the reductions and clips live inside
the fused CUDA loss kernel, so only the host-side initialiser remains."""
import numpy as np


def normc_initializer(std=1.0):
    """Gaussian kernel whose columns are L2-normalised to `std` (tf_util.py:286-291)."""
    def _initializer(shape, dtype=None, partition_info=None):
        out = np.random.randn(*shape).astype(np.float32)
        out *= std / np.sqrt(np.square(out).sum(axis=0, keepdims=True))
        return out
    return _initializer
