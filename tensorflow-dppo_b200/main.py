# -*- coding: utf-8 -*-
"""Entry point with the reference's parameter_dict (/root/reference/main.py:12-29), one process
per GPU:   python main.py                      (1 GPU)
           python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 main.py
Every rank builds NUM_WORKERS worker replicas (each stepping NUM_ENVS environments) and a Chief;
the tower mean of PPO.py:55-65 spans the ranks through dppo_allreduce_avg (NCCL over NVLink)."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from collections import deque                      # noqa: E402
from RL.Chief import Chief                         # noqa: E402
from RL.Worker import Worker                       # noqa: E402
from RL.session import Session, Coordinator, FileWriter   # noqa: E402


def build_parameter_dict(**overrides):
    parameter_dict = {
        'GAME': 'Pendulum-v0',                     # reference default: 'CartPole-v0' (main.py:13; also runs: Discrete -> CategoricalPd)
        'LEARNING_RATE': 2e-5,
        'ENTCOEFF': 0.01,
        'VCOEFF': 0.5,
        'CLIP_PARAM': 0.2,
        'GAMMA': 0.99,
        'LAM': 0.95,
        'SCHEDULE': 'linear',
        'MAX_AC_EXP_RATE': 0.4,
        'MIN_AC_EXP_RATE': 0.15,
        'AC_EXP_PERCENTAGE': 1,
        'UPDATE_STEPS': 4,
        'MAX_EPOCH_STEPS': 200,
        'EPOCH_MAX': 500,
        'NUM_WORKERS': 2,                          # reference: multiprocessing.cpu_count() threads
        'LOG_FILE_PATH': '/tmp/dppo_log',
        # new keys (defaults reproduce the reference's behaviour)
        'NUM_ENVS': 2,                             # environments per worker replica (reference: 1)
        'HIDDEN': (16,),                           # trunk widths (Model.py:12)
    }
    parameter_dict.update(overrides)
    return parameter_dict


def run(parameter_dict, max_updates=None):
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    parameter_dict.setdefault('DEVICE_INDEX', int(os.environ.get('LOCAL_RANK', 0)))
    SESS = Session()
    COORD = Coordinator()
    PUSH_EVENT, UPDATE_EVENT = threading.Event(), threading.Event()
    MEMORY_DICT = {}
    workers = []
    for i in range(parameter_dict['NUM_WORKERS']):
        i_name = 'Worker_N%i' % i
        MEMORY_DICT[i_name] = deque()
        workers.append(Worker(i_name, parameter_dict, SESS, MEMORY_DICT, COORD))
    chief = Chief('Chief', parameter_dict, SESS, MEMORY_DICT, COORD, workers)
    log_writer = FileWriter(parameter_dict['LOG_FILE_PATH'])
    if world > 1:
        import torch.distributed as dist
        from dppo_b200.dist import init_engine_comm
        if not dist.is_initialized():
            dist.init_process_group('nccl')
        init_engine_comm(chief.ppo.backend, rank, world)
        dist.broadcast(chief.ppo.pipara.flat, 0)                      # identical init on every rank
    for worker in workers:                                            # main.py:48-50
        worker.ppo.pipara.assign(chief.ppo.pipara)
        worker.ppo.oldpipara.assign(chief.ppo.pipara)
    start_time = time.time()
    PUSH_EVENT.set()
    updates = 0
    while not COORD.should_stop():
        for worker in workers:
            worker.work(PUSH_EVENT, UPDATE_EVENT, log_writer, iterations=1)
        chief.check(PUSH_EVENT, UPDATE_EVENT, once=True)
        updates += 1
        if max_updates is not None and updates >= max_updates:
            break
    if rank == 0:
        print("TRAINING FINISHED.")
        print("Train time elapsed:", time.time() - start_time, "seconds")
    return chief, workers


if __name__ == "__main__":
    run(build_parameter_dict(), max_updates=int(os.environ.get('DPPO_MAX_UPDATES', 0)) or None)
