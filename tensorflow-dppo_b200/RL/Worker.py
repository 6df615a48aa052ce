# -*- coding: utf-8 -*-
"""The reference's `Worker` (/root/reference/Worker.py) on the sm_100a engine: roll the policy out, turn the rollout
into advantages / returns, evaluate the losses for the log and queue the batch for the chief.

Public surface kept: `Worker(scope, parameter_dict, SESS, MEMORY_DICT, COORD)`, `work`, `act`, the attributes the
reference sets (`env`, `ppo`, `name`, `CUR_EP`, `cur_lr`, `returnAction`, the upper-case hyper-parameters).
What the target design changes: one replica steps NUM_ENVS environments at once (the engine's [T, N] buffer, one
column per rollout) and the GAE of Worker.py:82-92 is ONE launch of the CUDA scan (dppo_gae) over all N rollouts."""
import os

import numpy as np
import torch

from RL import envs
from RL.PPO import PPO
from dppo_b200.dist import stable_seed

# hyper-parameters copied from parameter_dict under their own names (Worker.py:16-27)
_COPIED_KEYS = ('EPOCH_MAX', 'MAX_EPOCH_STEPS', 'MAX_AC_EXP_RATE', 'MIN_AC_EXP_RATE', 'SCHEDULE', 'GAMMA', 'LAM',
                'LOG_FILE_PATH')


class Worker(object):
    def __init__(self, scope, parameter_dict, SESS, MEMORY_DICT, COORD):
        self.name, self.sess, self.MEMORY_DICT, self.COORD = scope, SESS, MEMORY_DICT, COORD
        for key in _COPIED_KEYS:
            setattr(self, key, parameter_dict[key])
        self.AC_EXP_EPOCH = parameter_dict['AC_EXP_PERCENTAGE'] * self.EPOCH_MAX
        self.ENV_SAMPLE_ITERATIONS = parameter_dict.get('ENV_SAMPLE_ITERATIONS')      # read, never used
        self.NUM_ENVS = int(parameter_dict.get('NUM_ENVS', 1))
        self.VERBOSE = parameter_dict.get('VERBOSE', True)
        # stable across processes (str hash() is salted), distinct per rank: ranks must not replay each other's rollouts
        seed = stable_seed(scope, parameter_dict.get('SEED', 0), os.environ.get('RANK', 0))
        self.env = envs.make(parameter_dict['GAME'], self.NUM_ENVS, seed=seed % 65536)
        self.ppo = PPO(scope, parameter_dict, self.env)
        self.CUR_EP, self.cur_lr, self.returnAction = 0, 1.0, None
        self._explore_rng = np.random.default_rng(seed % (2 ** 31))

    # ------------------------------------------------------------------------------------------------ rollout
    def _rollout(self):
        """T steps of N environments (Worker.py:39-65).  Returns the [T, N, .] buffers and the returns of the episodes
        that ended inside the rollout."""
        T, N = self.MAX_EPOCH_STEPS, self.NUM_ENVS
        act_dtype = np.int32 if self.ppo.discrete else np.float32
        # pinned host memory (when a CUDA device exists): the chief's fused update uploads these batches asynchronously
        def alloc(shape, dtype):
            t = torch.empty(shape, dtype=torch.int32 if dtype == np.int32 else torch.float32,
                            pin_memory=torch.cuda.is_available())
            return t.numpy()
        buf = dict(s=alloc((T, N, self.ppo.obs_dim), np.float32), a=alloc((T, N, self.ppo.act_cols), act_dtype),
                   r=alloc((T, N), np.float32), v=alloc((T, N), np.float32), d=alloc((T, N), np.float32))
        finished, running = [], np.zeros(N)
        state = self.env.reset()
        for t in range(T):
            action, value = self.act(state)
            nxt, reward, done, _ = self.env.step(action)
            buf['s'][t], buf['a'][t], buf['r'][t], buf['v'][t], buf['d'][t] = state, action, reward, value, done
            running += reward
            if done.any():
                finished.extend(running[done].tolist())
                running[done] = 0
                nxt = self.env.reset(done)
            state = nxt
        return buf, np.asarray(finished)

    def _anneal(self):
        """Worker.py:77-80: the multiplier of the learning rate and of the clip range."""
        if self.SCHEDULE == 'linear':
            self.cur_lr = max(1.0 - float(self.CUR_EP) / self.EPOCH_MAX, 0)
        elif self.SCHEDULE == 'constant':
            self.cur_lr = 1.0
        return self.cur_lr

    def work(self, PUSH_EVENT, UPDATE_EVENT, log_writer, iterations=None):
        """Worker.py:29-138.  `iterations=k` returns after k rollouts (synchronous driver)."""
        finished_iterations = 0
        while not self.COORD.should_stop():
            if not PUSH_EVENT.is_set():                            # Worker.py:40-42: the chief is training
                PUSH_EVENT.wait()
            self.sess.run(self.ppo.sync_pis)                       # oldpi <- pi before acting (PPO.py:47)
            buf, episode_returns = self._rollout()
            self.CUR_EP += 1
            if self.VERBOSE:
                print(self.name, "finished iterator", self.CUR_EP, "\n")
            l_mul = self._anneal()

            # Worker.py:82-92 for all N rollouts at once: the GAE scan kernel
            dev = self.ppo.device
            adv, etr = self.ppo.backend.gae(*(torch.from_numpy(buf[k]).to(dev) for k in ('r', 'v', 'd')), self.GAMMA, self.LAM)
            rows = buf['r'].size
            # queued as host arrays, like the numpy batches of Worker.py:94-95 (the chief's fused update uploads them)
            adv_h, etr_h = (x.detach().to('cpu').reshape(-1, 1).numpy() if isinstance(x, torch.Tensor) else
                            np.asarray(x, np.float32).reshape(-1, 1) for x in (adv, etr))
            batch = [buf['s'].reshape(rows, -1), buf['a'].reshape(rows, -1), adv_h, etr_h, l_mul]
            tower = self.ppo
            feed = {tower.s: batch[0], tower.action: batch[1], tower.advantage: batch[2], tower.estimatedReturn: batch[3],
                    tower.l_mul: l_mul}
            if self.name == 'Worker_N0':                           # Worker.py:112-114: one worker writes the summaries
                log_writer.add_summary(self.sess.run(tower.summary_op, feed), self.CUR_EP)
            losses = self.sess.run([tower.policyLoss, tower.valueLoss, tower.entropyLoss, tower.total_loss], feed)

            if episode_returns.size:                               # Worker.py:121,135 (nothing finished -> no push)
                spread = episode_returns.std()
                score = episode_returns.mean() / spread if spread > 0 else float('inf')
                batch.append([score, episode_returns.min(), episode_returns.max(), episode_returns.mean(), *losses, self.CUR_EP])
                self.MEMORY_DICT[self.name].append(batch)
            UPDATE_EVENT.set()
            finished_iterations += 1
            if iterations is not None and finished_iterations >= iterations:
                return

    # ------------------------------------------------------------------------------------------------ acting
    def _exploration_rate(self):
        """Worker.py:141-144: linear decay from MAX_AC_EXP_RATE to MIN_AC_EXP_RATE over AC_EXP_EPOCH epochs."""
        if self.CUR_EP >= self.AC_EXP_EPOCH:
            return self.MIN_AC_EXP_RATE
        return self.MAX_AC_EXP_RATE + self.CUR_EP * (self.MIN_AC_EXP_RATE - self.MAX_AC_EXP_RATE) / self.AC_EXP_EPOCH

    def act(self, s):
        """Worker.py:140-153 for a batch of N states: sample pi (dppo_act), read the value head.  The epsilon-greedy
        override of the reference is discrete-only (`action_space.n`): applied per environment for Discrete
        spaces, skipped for Box spaces."""
        sampled, value = self.sess.run([self.ppo.ca, self.ppo.pipredv], {self.ppo.s: s})
        chosen = sampled[:, 0, :]
        if self.ppo.discrete:
            n_env = len(chosen)
            explore = self._explore_rng.random(n_env) < self._exploration_rate()            # Worker.py:149-150
            uniform = self._explore_rng.integers(0, self.env.action_space.n, n_env)
            chosen = np.where(explore[:, None], uniform[:, None], chosen).astype(np.int32)
        self.returnAction = chosen
        return chosen, value[:, 0, 0]
