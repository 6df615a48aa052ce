"""The tf_util helpers the update path touches (/root/reference/Others/tf_util.py:14-24, 271-279,
286-291).  sum/mean/clip are graph-op aliases there; here the reductions and clips live inside
the fused CUDA loss kernel, so the host-side initialiser and save_state / load_state remain."""
import os

import numpy as np

_DEFAULT_PPO = None      # the chief's PPO object: what tf.train.Saver() would find in the default session


def set_default_ppo(ppo):
    global _DEFAULT_PPO
    _DEFAULT_PPO = ppo


def save_state(fname, ppo=None):
    """tf_util.py:276-279: every variable of the session to `fname` -- here the chief's pi / oldpi parameter vectors,
    its Adam slots and optimiser step, through dppo_save_state (one file, written atomically)."""
    ppo = ppo or _DEFAULT_PPO
    if ppo is None:
        raise RuntimeError("save_state: no chief PPO object exists yet")
    d = os.path.dirname(fname)
    if d:
        os.makedirs(d, exist_ok=True)
    ppo.backend.save_state(fname, ppo.pipara.flat, ppo.oldpipara.flat, ppo._m, ppo._v, ppo._t)


def load_state(fname, ppo=None):
    """tf_util.py:271-273: restore what save_state wrote (fails loudly on a different network)."""
    ppo = ppo or _DEFAULT_PPO
    if ppo is None:
        raise RuntimeError("load_state: no chief PPO object exists yet")
    ppo._t = ppo.backend.load_state(fname, ppo.pipara.flat, ppo.oldpipara.flat, ppo._m, ppo._v)


def normc_initializer(std=1.0):
    """Gaussian kernel whose columns are L2-normalised to `std` (tf_util.py:286-291)."""
    def _initializer(shape, dtype=None, partition_info=None):
        out = np.random.randn(*shape).astype(np.float32)
        out *= std / np.sqrt(np.square(out).sum(axis=0, keepdims=True))
        return out
    return _initializer
