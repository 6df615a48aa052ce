# -*- coding: utf-8 -*-
"""The reference's `Chief` (/root/reference/Chief.py) on the sm_100a engine: collect the workers' rollouts, train the
chief parameters on them UPDATE_STEPS times, hand the new weights back.

Public surface kept: `Chief(scope, parameter_dict, SESS, MEMORY_DICT, COORD, workers)`, `check`, `_train`,
`_update_local_workers_weight`, `_get_data_size`, `_logs_writer`, `act`.  One Chief object lives on every rank -- there is
no chief rank: every rank applies the same all-reduced update."""
import numpy as np

from RL import envs
from RL.PPO import PPO
from dppo_b200.dist import all_ranks_true

# positions inside one pushed batch (Worker.py:94-136): states, actions, advantages, returns, lr multiplier, logs
_S, _A, _ADV, _RET, _LMUL, _LOGS = range(6)
# positions inside the logs list (Worker.py:123-133)
_LOG_MAX_RETURN, _LOG_EPOCH = 2, 8


class Chief(object):
    def __init__(self, scope, parameter_dict, SESS, MEMORY_DICT, COORD, workers):
        self.sess, self.COORD, self.MEMORY_DICT, self.workers = SESS, COORD, MEMORY_DICT, workers
        for key in ('UPDATE_STEPS', 'NUM_WORKERS', 'EPOCH_MAX'):
            setattr(self, key, parameter_dict[key])
        self.ppo = PPO(scope, parameter_dict, envs.make(parameter_dict['GAME'], 1), workers)

    # ---------------------------------------------------------------------------------------------------------
    def check(self, PUSH_EVENT, UPDATE_EVENT, once=False):
        """Chief.py:19-28: on every UPDATE_EVENT, if each worker has pushed at least one batch, stop the workers from
        pushing, train, redistribute the weights and let them continue.  `once=True` handles a single event (one process
        per GPU drives the loop synchronously instead of parking the chief in a thread)."""
        while not self.COORD.should_stop():
            UPDATE_EVENT.wait()
            fewest, _ = self._get_data_size()
            # one Chief per rank: _train contains the cross-rank gradient exchange, so every rank must take the same
            # branch -- train only when EVERY rank has a batch from each of its workers (collective AND)
            if all_ranks_true(fewest > 0):
                PUSH_EVENT.clear()
                try:
                    self._train()
                    self._update_local_workers_weight()
                finally:
                    PUSH_EVENT.set()
            UPDATE_EVENT.clear()
            if once:
                return

    def _drain(self):
        """Every queued batch of every worker, round-robin over the deques (Chief.py:33-48)."""
        batches = []
        while any(self.MEMORY_DICT.values()):
            batches.extend(list(q.popleft()) for q in self.MEMORY_DICT.values() if q)
        return batches

    def _train(self):
        ranked = sorted(self._drain(), key=lambda b: b[_LOGS][_LOG_MAX_RETURN], reverse=True)   # Chief.py:51
        l_mul = ranked[0][_LMUL]
        feed = {self.ppo.l_mul: l_mul}
        for worker, batch in zip(self.workers[:self.NUM_WORKERS], ranked):     # tower i <- i-th best rollout, Chief.py:58-63
            tower = worker.ppo
            feed.update({tower.s: batch[_S], tower.act: batch[_A], tower.adv: batch[_ADV], tower.str: batch[_RET],
                         tower.l_mul: l_mul})
        # Chief.py:64 runs `train` UPDATE_STEPS times on this one feed; here the k steps are ONE fused engine call
        # (one upload of the towers' batches, k x [forward, loss, backward, tower mean, Adam] on the device)
        feed[self.ppo.update_steps] = self.UPDATE_STEPS
        self.sess.run(self.ppo.train, feed_dict=feed)
        self._logs_writer(ranked)

    def _update_local_workers_weight(self):
        """Chief.py:67-70: every worker's pi <- the chief's pi."""
        for w in self.workers:
            w.ppo.pipara.assign(self.ppo.pipara)

    def _get_data_size(self):
        """(fewest, most) batches queued by any worker (Chief.py:72-78)."""
        sizes = [len(q) for q in self.MEMORY_DICT.values()]
        return (min(sizes), max(sizes)) if sizes else (0, -1)

    def _logs_writer(self, data_stack):
        """Stop condition of Chief.py:80-87; the reference hard-codes 500 epochs there, EPOCH_MAX here."""
        if data_stack[0][_LOGS][_LOG_EPOCH] >= self.EPOCH_MAX:
            self.COORD.request_stop()

    def act(self, s):
        """Chief.py:89-92: the action for ONE state."""
        return self.sess.run(self.ppo.ca, {self.ppo.s: np.asarray(s)[None]})[0][0]
