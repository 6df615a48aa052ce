# -*- coding: utf-8 -*-
"""Chief of the reference (/root/reference/Chief.py): waits for rollouts, averages the towers'
gradients, applies Adam, broadcasts the weights.  Same constructor and methods (`check`, `_train`,
`_update_local_workers_weight`, `_get_data_size`, `_logs_writer`, `act`).  One Chief object lives
on every rank (there is no chief rank: every rank applies the same all-reduced update)."""
from collections import deque
import sys

import numpy as np

from RL import envs
from RL.PPO import PPO


class Chief(object):
    def __init__(self, scope, parameter_dict, SESS, MEMORY_DICT, COORD, workers):
        env = envs.make(parameter_dict['GAME'], 1)
        self.ppo = PPO(scope, parameter_dict, env, workers)
        self.sess = SESS
        self.MEMORY_DICT = MEMORY_DICT
        self.workers = workers
        self.UPDATE_STEPS = parameter_dict['UPDATE_STEPS']
        self.COORD = COORD
        self.NUM_WORKERS = parameter_dict['NUM_WORKERS']
        self.EPOCH_MAX = parameter_dict['EPOCH_MAX']

    def check(self, PUSH_EVENT, UPDATE_EVENT, once=False):
        """Chief.py:19-28.  `once=True` runs a single pass (one process per GPU drives the loop
        synchronously instead of parking the chief in a thread)."""
        while not self.COORD.should_stop():
            UPDATE_EVENT.wait()
            min_data_size, _ = self._get_data_size()
            if min_data_size >= 1:
                PUSH_EVENT.clear()
                self._train()
                self._update_local_workers_weight()
                PUSH_EVENT.set()
            UPDATE_EVENT.clear()
            if once:
                break

    def _train(self):
        data_stack = []
        _, max_data_size = self._get_data_size()
        while max_data_size > 0:                                   # Chief.py:33-48: drain every deque
            for key, value in self.MEMORY_DICT.items():
                if len(value) > 0:
                    data_stack.append(list(value.popleft()))
            _, max_data_size = self._get_data_size()
        data_stack.sort(key=lambda x: x[5][2], reverse=True)       # Chief.py:51: best max-episode-return first

        feed_dict = {}
        l_mul = data_stack[0][4]
        feed_dict[self.ppo.l_mul] = l_mul
        for i in range(self.NUM_WORKERS):                          # Chief.py:58-63: tower i <- i-th best rollout
            feed_dict[self.workers[i].ppo.s] = data_stack[i][0]
            feed_dict[self.workers[i].ppo.act] = data_stack[i][1]
            feed_dict[self.workers[i].ppo.adv] = data_stack[i][2]
            feed_dict[self.workers[i].ppo.str] = data_stack[i][3]
            feed_dict[self.workers[i].ppo.l_mul] = l_mul
        [self.sess.run(self.ppo.train, feed_dict=feed_dict) for _ in range(self.UPDATE_STEPS)]
        self._logs_writer(data_stack)

    def _update_local_workers_weight(self):
        for worker in self.workers:                                # Chief.py:67-70
            worker.ppo.pipara.assign(self.ppo.pipara)

    def _get_data_size(self):
        min_data_size = sys.maxsize
        max_data_size = -1
        for key, value in self.MEMORY_DICT.items():
            min_data_size = min(min_data_size, len(value))
            max_data_size = max(max_data_size, len(value))
        return min_data_size, max_data_size

    def _logs_writer(self, data_stack):
        logs = [item[5] for item in data_stack]
        localsteps = logs[0][8]
        if localsteps >= self.EPOCH_MAX:                           # Chief.py:86 hard-codes 500
            self.COORD.request_stop()

    def act(self, s):
        s = np.array([s])
        action = self.sess.run(self.ppo.ca, {self.ppo.s: s})
        return action[0][0]
