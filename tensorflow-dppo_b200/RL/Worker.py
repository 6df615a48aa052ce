# -*- coding: utf-8 -*-
"""Worker of the reference (/root/reference/Worker.py): rolls the policy out, computes GAE
returns/advantages, evaluates the losses for logging and pushes the batch to the chief.

Same constructor and methods (`work`, `act`).  Differences forced by the target design: one
Worker replica steps NUM_ENVS environments at once (the [T,N] buffer of the engine) and the GAE
of Worker.py:82-92 runs as the CUDA scan (dppo_gae) over all N rollouts."""
from collections import deque

import numpy as np
import torch

from RL import envs
from RL.PPO import PPO


class Worker(object):
    def __init__(self, scope, parameter_dict, SESS, MEMORY_DICT, COORD):
        self.NUM_ENVS = int(parameter_dict.get('NUM_ENVS', 1))
        self.env = envs.make(parameter_dict['GAME'], self.NUM_ENVS, seed=abs(hash(scope)) % 65536)
        self.ppo = PPO(scope, parameter_dict, self.env)
        self.COORD = COORD
        self.MEMORY_DICT = MEMORY_DICT
        self.name = scope
        self.sess = SESS
        self.CUR_EP = 0
        self.EPOCH_MAX = parameter_dict['EPOCH_MAX']
        self.MAX_EPOCH_STEPS = parameter_dict['MAX_EPOCH_STEPS']
        self.MAX_AC_EXP_RATE = parameter_dict['MAX_AC_EXP_RATE']
        self.MIN_AC_EXP_RATE = parameter_dict['MIN_AC_EXP_RATE']
        self.AC_EXP_EPOCH = parameter_dict['AC_EXP_PERCENTAGE'] * parameter_dict['EPOCH_MAX']
        self.SCHEDULE = parameter_dict['SCHEDULE']
        self.GAMMA = parameter_dict['GAMMA']
        self.LAM = parameter_dict['LAM']
        self.ENV_SAMPLE_ITERATIONS = parameter_dict.get('ENV_SAMPLE_ITERATIONS')   # read, never used (D5)
        self.LOG_FILE_PATH = parameter_dict['LOG_FILE_PATH']
        self.VERBOSE = parameter_dict.get('VERBOSE', True)
        self._explore_rng = np.random.default_rng(abs(hash(scope)) % (2 ** 31))

    def work(self, PUSH_EVENT, UPDATE_EVENT, log_writer, iterations=None):
        """Worker.py:29-138.  `iterations=k` returns after k rollouts (synchronous driver)."""
        done_iters = 0
        while not self.COORD.should_stop():
            N, T = self.NUM_ENVS, self.MAX_EPOCH_STEPS
            buffer_s = np.empty((T, N, self.ppo.obs_dim), np.float32)
            buffer_a = np.empty((T, N, self.ppo.act_cols), np.int32 if self.ppo.discrete else np.float32)
            buffer_r = np.empty((T, N), np.float32)
            buffer_predv = np.empty((T, N), np.float32)
            buffer_done = np.empty((T, N), np.float32)
            buffer_epr = []
            if not PUSH_EVENT.is_set():                          # Worker.py:40-42
                PUSH_EVENT.wait()
            self.sess.run(self.ppo.sync_pis)
            s = self.env.reset()
            EPR = np.zeros(N)
            for index in range(T):                               # Worker.py:39-65
                action, pred_v = self.act(s)
                s_, r, done, _ = self.env.step(action)
                buffer_s[index], buffer_a[index], buffer_r[index] = s, action, r
                buffer_predv[index], buffer_done[index] = pred_v, done
                EPR += r
                s = s_
                if done.any():
                    buffer_epr.extend(EPR[done].tolist())
                    EPR[done] = 0
                    s = self.env.reset(done)
            self.CUR_EP += 1
            if self.VERBOSE:
                print(self.name, "finished iterator", self.CUR_EP, "\n")

            if self.SCHEDULE == 'constant':                      # Worker.py:77-80
                self.cur_lr = 1.0
            elif self.SCHEDULE == 'linear':
                self.cur_lr = max(1.0 - float(self.CUR_EP) / self.EPOCH_MAX, 0)

            # Worker.py:82-92 for all N rollouts: one launch of the GAE scan kernel
            dev = self.ppo.device
            buffer_adv, buffer_etr = self.ppo.backend.gae(
                torch.from_numpy(buffer_r).to(dev), torch.from_numpy(buffer_predv).to(dev),
                torch.from_numpy(buffer_done).to(dev), self.GAMMA, self.LAM)

            batchs = [buffer_s.reshape(T * N, -1), buffer_a.reshape(T * N, -1), buffer_adv.reshape(-1, 1),
                      buffer_etr.reshape(-1, 1), self.cur_lr]
            feed_dict = {
                self.ppo.s: batchs[0],
                self.ppo.action: batchs[1],
                self.ppo.advantage: batchs[2],
                self.ppo.estimatedReturn: batchs[3],
                self.ppo.l_mul: self.cur_lr,
            }
            if (self.name == 'Worker_N0'):                       # Worker.py:112-114
                summary = self.sess.run(self.ppo.summary_op, feed_dict)
                log_writer.add_summary(summary, self.CUR_EP)
            queryItem = [self.ppo.policyLoss, self.ppo.valueLoss, self.ppo.entropyLoss, self.ppo.total_loss]
            policyLoss, valueLoss, entropyLoss, totalLoss = self.sess.run(queryItem, feed_dict)

            buffer_epr = np.array(buffer_epr)
            if len(buffer_epr) > 0:                              # Worker.py:121,135 (empty -> skip)
                std = buffer_epr.std()
                score = buffer_epr.mean() / std if std > 0 else float('inf')
                logs = [score, buffer_epr.min(), buffer_epr.max(), buffer_epr.mean(), policyLoss, valueLoss,
                        entropyLoss, totalLoss, self.CUR_EP]
                batchs.append(logs)
                self.MEMORY_DICT[self.name].append(batchs)
            UPDATE_EVENT.set()
            done_iters += 1
            if iterations is not None and done_iters >= iterations:
                break

    def act(self, s):
        """Worker.py:140-153 for a batch of N states.  The epsilon-greedy override there is
        discrete-only (`action_space.n`): applied per env for Discrete spaces, dropped for Box spaces."""
        action, pred_v = self.sess.run([self.ppo.ca, self.ppo.pipredv], {self.ppo.s: s})
        self.returnAction = action[:, 0, :]
        if self.ppo.discrete:
            if self.CUR_EP >= self.AC_EXP_EPOCH:                             # Worker.py:141-144
                cur_exp_rate = self.MIN_AC_EXP_RATE
            else:
                cur_exp_rate = self.MAX_AC_EXP_RATE + self.CUR_EP * (self.MIN_AC_EXP_RATE - self.MAX_AC_EXP_RATE) / self.AC_EXP_EPOCH
            explore = self._explore_rng.random(len(self.returnAction)) < cur_exp_rate          # Worker.py:149-150
            rand = self._explore_rng.integers(0, self.env.action_space.n, len(self.returnAction))
            self.returnAction = np.where(explore[:, None], rand[:, None], self.returnAction).astype(np.int32)
        return self.returnAction, pred_v[:, 0, 0]
