"""Actor-critic MLP of the reference (/root/reference/Model.py).

`Model().FC(scope, s, action_space)` keeps the reference signature (and accepts the `state=`
spelling PPO.py:21-22 uses).  Instead of adding TF ops to a graph it allocates the
flat fp32 parameter vector of that scope on the device, in the reference's variable creation
order (trunk kernel/bias, value head, policy head -- Model.py:12-14,17), initialised with
normc_initializer(0.01) kernels and zero biases (Model.py:10), and returns handles:
    predv -> value head output,  pd -> PdHandle (flatparam of the DiagGaussianPd),  para -> ParamSet.
The forward pass itself is the CUDA path (dppo_forward / dppo_loss_grad)."""
import numpy as np
import torch

from RL.Others.distributions import make_pdtype
from RL.Others import tf_util as U
from dppo_b200.layout import ParamLayout


class ParamSet:
    """tf.get_collection(TRAINABLE_VARIABLES, scope) stand-in: one flat device vector + named views."""
    def __init__(self, scope, layout: ParamLayout, flat: torch.Tensor):
        self.scope, self.layout, self.flat = scope, layout, flat

    def __iter__(self):                         # iterating yields the per-variable views, Model.py:17 order
        return iter(self.layout.views(self.flat).values())

    def __len__(self):
        return len(self.layout.entries)

    def assign(self, other: "ParamSet"):
        """localp.assign(chiefp) for every variable (main.py:48-50, Chief.py:67-70, PPO.py:47)."""
        self.flat.copy_(other.flat)             # device memcpy; plumbing, not compute
        return self


class Model(object):
    HIDDEN = (16,)                              # Model.py:12: one dense(16, relu) trunk layer

    def __init__(self, hidden=None):
        self.hidden = tuple(hidden) if hidden is not None else self.HIDDEN

    def FC(self, scope, s=None, action_space=None, state=None, device="cuda"):
        s = state if s is None else s
        obs_dim = int(s.shape[-1])
        pdtype = make_pdtype(action_space)
        pd_width = pdtype.param_shape()[0]                      # Model.py:14
        discrete = hasattr(action_space, "n")
        act_dim = pd_width if discrete else pd_width // 2
        layout = ParamLayout(obs_dim, self.hidden, act_dim, pd_width)
        initializer = U.normc_initializer(0.01)                 # Model.py:10
        theta = np.zeros(layout.n_params, np.float32)
        for e in layout.entries:
            if e.name[0] == "W":
                theta[e.offset:e.offset + e.size] = initializer(e.shape).ravel()
        para = ParamSet(scope, layout, torch.from_numpy(theta).to(device))
        return ("predv", scope), (pdtype, scope), para
