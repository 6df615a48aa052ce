"""Flat parameter layout of one actor-critic net (host-side bookkeeping only).

Order = creation order of the trainable variables in the reference
(/root/reference/Model.py:12-14,17): trunk kernels/biases, value head, policy head; kernels are
[in, out] row-major (tf.layers.dense `kernel`).  The trunk is generalised from the reference's
single 16-wide layer to L hidden layers.
"""
from dataclasses import dataclass
from typing import List, Sequence, Tuple


@dataclass(frozen=True)
class ParamEntry:
    name: str
    shape: Tuple[int, ...]
    offset: int

    @property
    def size(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


class ParamLayout:
    def __init__(self, obs_dim: int, hidden: Sequence[int], act_dim: int, pd_width: int = None):
        """pd_width = pdtype.param_shape()[0] (Model.py:14): 2*act_dim for DiagGaussianPd (default), ncat for CategoricalPd."""
        self.obs_dim, self.hidden, self.act_dim = int(obs_dim), tuple(int(h) for h in hidden), int(act_dim)
        pw = self.pd_width = 2 * self.act_dim if pd_width is None else int(pd_width)
        entries: List[ParamEntry] = []
        off, fan_in = 0, self.obs_dim
        for l, h in enumerate(self.hidden):
            entries.append(ParamEntry(f"W{l + 1}", (fan_in, h), off)); off += fan_in * h
            entries.append(ParamEntry(f"b{l + 1}", (h,), off)); off += h
            fan_in = h
        entries.append(ParamEntry("Wv", (fan_in, 1), off)); off += fan_in          # Model.py:13
        entries.append(ParamEntry("bv", (1,), off)); off += 1
        entries.append(ParamEntry("Wp", (fan_in, pw), off)); off += fan_in * pw                         # :14
        entries.append(ParamEntry("bp", (pw,), off)); off += pw
        self.entries = entries
        self.n_params = off

    def offsets(self) -> List[int]:
        return [e.offset for e in self.entries]

    def views(self, flat):
        """name -> view of a flat tensor/array (torch or numpy)."""
        return {e.name: flat[e.offset:e.offset + e.size].reshape(e.shape) for e in self.entries}
