# -*- coding: utf-8 -*-
"""PPO graph object of the reference (/root/reference/PPO.py), re-hosted on the sm_100a engine.

Same constructor `PPO(scope, parameter_dict, env, workerLists=None)` and the same attribute
surface: placeholders `s, l_mul, a, adv, etr`; fetchables `pipredv, ca, policyLoss, entropyLoss,
valueLoss, total_loss, summary_op, gradient, sync_pis`; chief-only `train`; parameter lists
`pipara, oldpipara`; `_average_gradients`.  Evaluating a fetchable (Session.run) launches CUDA:

    policyLoss / valueLoss / entropyLoss / total_loss / gradient  -> dppo_loss_grad   (PPO.py:21-46)
    ca / pipredv                                                   -> dppo_act         (PPO.py:23)
    sync_pis                                                       -> dppo_sync_old    (PPO.py:47)
    train                                                          -> dppo_loss_grad over the towers'
                         rollouts + dppo_allreduce_avg + dppo_adam_apply              (PPO.py:49-65)

The spellings Worker.py / Chief.py use for the placeholders (`action`, `advantage`,
`estimatedReturn`, `act`, `str` -- SURVEY D4) are kept as aliases.
"""
import os

import numpy as np
import torch

from RL.Model import Model
from RL.session import Placeholder, Fetch
from dppo_b200.dist import stable_seed

_ENGINES = {}


def _get_backend(parameter_dict, obs_dim, act_dim, hidden, discrete=False):
    """One engine per process / GPU, shared by every PPO object (they shared one tf.Session, main.py:31)."""
    backend = parameter_dict.get('BACKEND')
    if backend is not None:
        return backend
    rows = int(parameter_dict['MAX_EPOCH_STEPS']) * int(parameter_dict.get('NUM_ENVS', 1)) \
        * int(parameter_dict['NUM_WORKERS'])
    key = (obs_dim, act_dim, tuple(hidden), rows, parameter_dict.get('DEVICE_INDEX', 0), bool(discrete))
    if key not in _ENGINES:
        from dppo_b200 import DppoEngine                   # raises if the CUDA library is missing
        _ENGINES[key] = DppoEngine(obs_dim, act_dim, hidden, max(rows, 128), device=key[4],
                                   entcoeff=parameter_dict['ENTCOEFF'], vcoeff=parameter_dict['VCOEFF'],
                                   clip_param=parameter_dict['CLIP_PARAM'], pd_type=1 if discrete else 0)
    return _ENGINES[key]


class PPO(object):
    _LOSS_KINDS = ('policyLoss', 'valueLoss', 'entropyLoss', 'total_loss', 'summary_op', 'gradient')

    def __init__(self, scope, parameter_dict, env, workerLists = None):
        self.scope = scope
        self.LEARNING_RATE = parameter_dict['LEARNING_RATE']
        self.CLIP_PARAM = parameter_dict['CLIP_PARAM']
        self.ENTCOEFF = parameter_dict['ENTCOEFF']
        self.VCOEFF = parameter_dict['VCOEFF']
        self.NUM_WORKERS = parameter_dict['NUM_WORKERS']
        hidden = tuple(parameter_dict.get('HIDDEN', Model.HIDDEN))
        observation_space = env.observation_space
        action_space = env.action_space
        self.obs_dim = observation_space.shape[0]
        self.discrete = hasattr(action_space, 'n')                       # spaces.Discrete -> CategoricalPd
        self.act_dim = action_space.n if self.discrete else action_space.shape[0]
        self.act_cols = 1 if self.discrete else self.act_dim             # PPO.py:26: a is int32 [None, 1]
        self.backend = _get_backend(parameter_dict, self.obs_dim, self.act_dim, hidden, self.discrete)
        self.device = getattr(self.backend, 'device', 'cuda')
        model = Model(hidden)
        self.s = Placeholder(scope + '/s', [None, self.obs_dim])
        self.l_mul = Placeholder(scope + '/learningRate_multiplier', [])
        _, _, self.pipara = model.FC(scope=scope + 'pi', state=self.s, action_space=action_space, device=self.device)
        _, _, self.oldpipara = model.FC(scope=scope + 'oldpi', state=self.s, action_space=action_space,
                                        device=self.device)
        self.pipredv = Fetch(self, 'pipredv')
        self.ca = Fetch(self, 'ca')
        self._seed = stable_seed(scope, parameter_dict.get('SEED', 0), os.environ.get('RANK', 0))
        self._draws = 0

        if scope != "Chief":
            self.a = Placeholder(scope + '/a', [None, self.act_cols])
            self.adv = Placeholder(scope + '/adv', [None, 1])
            self.etr = Placeholder(scope + '/etr', [None, 1])
            self.action = self.act = self.a                              # Worker.py:106, Chief.py:60
            self.advantage = self.adv                                    # Worker.py:107
            self.estimatedReturn = self.str = self.etr                   # Worker.py:108, Chief.py:62
            self.policyLoss = Fetch(self, 'policyLoss')
            self.entropyLoss = Fetch(self, 'entropyLoss')
            self.valueLoss = Fetch(self, 'valueLoss')
            self.total_loss = Fetch(self, 'total_loss')
            self.summary_op = Fetch(self, 'summary_op')
            self.gradient = Fetch(self, 'gradient')
            self.sync_pis = [Fetch(self, 'sync_pis')]
        else:
            self.workerLists = workerLists
            n = self.pipara.flat.numel()
            self._m = torch.zeros(n, dtype=torch.float32, device=self.device)   # Adam slots (PPO.py:20)
            self._v = torch.zeros(n, dtype=torch.float32, device=self.device)
            self._t = 0
            self.train = Fetch(self, 'train')
            # feed `update_steps: k` next to `train` and the chief performs k optimisation steps on the same feed in ONE
            # fused engine call (what Chief.py:64 spells as a loop of k sess.run calls); absent -> one step per run
            self.update_steps = Placeholder(scope + '/update_steps', [])
            from RL.Others import tf_util as U
            U.set_default_ppo(self)            # what tf.train.Saver() (tf_util.py:271-279) would save / restore
            self.last_losses = None

    # -------------------------------------------------------------------------------------------
    def _dev(self, x, cols=None):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), dtype=torch.float32)
        t = t.to(device=self.device, dtype=torch.float32)
        if cols is not None:
            t = t.reshape(-1, cols)
        return t.contiguous()

    def _host_feeds(self, feed_dict):
        """This tower's feed as contiguous HOST tensors (no copy when the worker queued fp32 / int32 arrays, as it does):
        what dppo_train_host uploads itself, chunked and overlapped with compute."""
        def host(x, dtype, cols=None):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            t = t.detach().to(device='cpu', dtype=dtype)
            if cols is not None:
                t = t.reshape(-1, cols)
            return t.contiguous()
        s = host(feed_dict[self.s], torch.float32, self.obs_dim)
        a = host(feed_dict[self.a], torch.int32 if self.discrete else torch.float32, self.act_cols)
        adv = host(feed_dict[self.adv], torch.float32).reshape(-1)
        etr = host(feed_dict[self.etr], torch.float32).reshape(-1)
        return s, a, adv, etr

    def _feeds(self, feed_dict):
        s = self._dev(feed_dict[self.s], self.obs_dim)
        if self.discrete:
            a = torch.as_tensor(np.asarray(feed_dict[self.a]).astype(np.int32)).reshape(-1, 1).to(self.device).contiguous()
        else:
            a = self._dev(feed_dict[self.a], self.act_dim)
        adv = self._dev(feed_dict[self.adv]).reshape(-1)
        etr = self._dev(feed_dict[self.etr]).reshape(-1)
        return s, a, adv, etr

    def _average_gradients(self, tower_grads):
        """PPO.py:55-65.  Towers of one process are batched into ONE loss_grad call (equal-sized
        towers: mean of tower means == mean over the concatenated rows); across processes the mean
        is ncclAllReduce(ncclAvg) on the flat buffer.  `tower_grads` is that flat buffer."""
        if getattr(self.backend, 'world', 1) > 1:
            self.backend.allreduce_avg(tower_grads)
        return tower_grads

    def _run(self, kinds, feed_dict):
        lock = getattr(self.backend, 'lock', None)
        if lock is None:
            return self._run_locked(kinds, feed_dict)
        with lock:                       # all PPO objects of a process share ONE engine / workspace
            return self._run_locked(kinds, feed_dict)

    def _run_locked(self, kinds, feed_dict):
        vals = {}
        be = self.backend
        if any(k in self._LOSS_KINDS for k in kinds):
            s, a, adv, etr = self._feeds(feed_dict)
            l_mul = float(feed_dict[self.l_mul])
            losses, g = be.loss_grad(self.pipara.flat, self.oldpipara.flat, s, a, adv, etr, l_mul)
            lo = [float(x) for x in losses.tolist()]
            vals.update(policyLoss=lo[0], valueLoss=lo[1], entropyLoss=lo[2], total_loss=lo[3], gradient=g,
                        summary_op=dict(entropyLoss=lo[2], policyLoss=lo[0], valueLoss=lo[1], total_loss=lo[3]))
        if 'ca' in kinds or 'pipredv' in kinds:
            s = self._dev(feed_dict[self.s], self.obs_dim)
            self._draws += 1
            a, v = be.act(self.pipara.flat, s, self._seed, self._draws)
            vals['ca'] = a.cpu().numpy()[:, None, :]             # [B,1,A] (or [B,1,1] class index)  (Model.py:11 rank-3 convention)
            vals['pipredv'] = v.cpu().numpy()[:, None, None]     # [B,1,1]
        if 'sync_pis' in kinds:
            be.sync_old(self.pipara.flat, self.oldpipara.flat)   # PPO.py:47
            vals['sync_pis'] = None
        if 'train' in kinds and getattr(self, 'update_steps', None) in feed_dict and hasattr(be, 'train_host'):
            # Chief._train in ONE engine call (dppo_train_host): the towers' host batches are uploaded once (chunked,
            # overlapped with the first step), UPDATE_STEPS x [fwd, loss, bwd, tower mean, Adam] run on the device
            towers = [w.ppo for w in self.workerLists[:self.NUM_WORKERS]]
            l_mul = float(feed_dict[self.l_mul])
            k = int(feed_dict[self.update_steps])
            losses, g = be.train_host(self.pipara.flat, towers[0].oldpipara.flat, [t._host_feeds(feed_dict) for t in towers],
                                      k, l_mul, self._m, self._v, self._t + 1, self.LEARNING_RATE * l_mul)
            self._t += k
            self.last_losses = g[self.pipara.flat.numel():]
            self.step_losses = losses
            vals['train'] = None
        elif 'train' in kinds:
            # step-by-step path (a backend without the fused entry point, or `train` run without update_steps)
            towers = [w.ppo for w in self.workerLists[:self.NUM_WORKERS]]
            feeds = [t._feeds(feed_dict) for t in towers]
            s, a, adv, etr = (torch.cat([f[i] for f in feeds]) for i in range(4))
            l_mul = float(feed_dict[self.l_mul])
            for _ in range(int(feed_dict.get(getattr(self, 'update_steps', None), 1))):
                _, g = be.loss_grad(self.pipara.flat, towers[0].oldpipara.flat, s, a, adv, etr, l_mul)
                g = self._average_gradients(g)
                self._t += 1
                be.adam_apply(self.pipara.flat, g, self._m, self._v, self._t, self.LEARNING_RATE * l_mul)   # PPO.py:20,53
            self.last_losses = g[self.pipara.flat.numel():]
            vals['train'] = None
        return vals
