#######################################################################
# This file restates an interface of ShangtongZhang/DeepRL, whose     #
# licence asks that the following declaration stay at the top:        #
#                                                                     #
# Copyright (C) 2017 Shangtong Zhang(zhangshangtong.cpp@gmail.com)    #
# Permission given to modify the code as long as you keep this        #
# declaration at the top                                              #
#######################################################################
"""Agent base class and actor with the reference's interface (``deep_rl/agent/BaseAgent.py``:
``BaseAgent``:15 -- save / load / eval_step / eval_episodes / record_online_return / switch_task;
``BaseActor``:108 -- ``step()`` returns ``sgd_update_frequency`` transitions).

The reference runs the asynchronous actor in a forked subprocess that shares the network through
``share_memory()`` (BaseAgent.py:142-155, DQN_agent.py:58).  A CUDA context does not survive ``fork`` and the
parameters live in HBM, so ``async_actor=True`` is served by a daemon THREAD with the same 2-deep transition
cache and the same ``config.lock`` around the actor's forward pass and the optimizer step.
"""
import pickle
import queue
import threading

import numpy as np
import torch

from ..utils import close_obj, get_logger, mkdir


class BaseAgent:
    def __init__(self, config):
        self.config = config
        self.logger = get_logger(tag=config.tag, log_level=config.log_level)
        self.task_ind = 0

    def close(self):
        close_obj(self.task)

    def save(self, filename):
        """BaseAgent.py:24-27: ``<f>.model`` = network state_dict, ``<f>.stats`` = pickled normalizer stats."""
        torch.save(self.network.state_dict(), "%s.model" % filename)
        with open("%s.stats" % filename, "wb") as f:
            pickle.dump(self.config.state_normalizer.state_dict(), f)

    def load(self, filename):
        sd = torch.load("%s.model" % filename, map_location=lambda storage, loc: storage)
        self.network.load_state_dict(sd)
        with open("%s.stats" % filename, "rb") as f:
            self.config.state_normalizer.load_state_dict(pickle.load(f))

    def eval_step(self, state):
        raise NotImplementedError

    def eval_episode(self):
        env = self.config.eval_env
        state = env.reset()
        while True:
            action = self.eval_step(state)
            state, reward, done, info = env.step(action)
            ret = info[0]["episodic_return"]
            if ret is not None:
                return ret

    def eval_episodes(self):
        rets = [np.sum(self.eval_episode()) for _ in range(self.config.eval_episodes)]
        self.logger.info("steps %d, episodic_return_test %.2f(%.2f)" % (
            self.total_steps, np.mean(rets), np.std(rets) / np.sqrt(len(rets))))
        self.logger.add_scalar("episodic_return_test", np.mean(rets), self.total_steps)
        return {"episodic_return_test": np.mean(rets)}

    def record_online_return(self, info, offset=0):
        if isinstance(info, dict):
            ret = info["episodic_return"]
            if ret is not None:
                self.logger.add_scalar("episodic_return_train", ret, self.total_steps + offset)
                self.logger.info("steps %d, episodic_return_train %s" % (self.total_steps + offset, ret))
        elif isinstance(info, tuple):
            for i, item in enumerate(info):
                self.record_online_return(item, i)
        else:
            raise NotImplementedError

    def switch_task(self):
        config = self.config
        if not config.tasks:
            return
        segs = np.linspace(0, config.max_steps, len(config.tasks) + 1)
        if self.total_steps > segs[self.task_ind + 1]:
            self.task_ind += 1
            self.task = config.tasks[self.task_ind]
            self.states = config.state_normalizer(self.task.reset())

    def record_step(self, state):
        raise NotImplementedError


class BaseActor:
    STEP, RESET, EXIT, SPECS, NETWORK, CACHE = range(6)

    def __init__(self, config):
        self.config = config
        self._state = None
        self._task = None
        self._network = None
        self._total_steps = 0
        self._cache_len = 2
        self._thread = None
        self._queue = None
        self._stop = threading.Event()
        if not config.async_actor:
            self._set_up()
            self._task = config.task_fn()

    # -- reference-shaped surface
    def start(self):
        if self.config.async_actor and self._thread is None:
            self._queue = queue.Queue(maxsize=self._cache_len)
            self._thread = threading.Thread(target=self.run, daemon=True)

    def _sample(self):
        out = []
        for _ in range(self.config.sgd_update_frequency):
            t = self._transition()
            if t is not None:
                out.append(t)
        return out

    def run(self):
        self._set_up()
        self._task = self.config.task_fn()
        while not self._stop.is_set():
            item = self._sample()
            while not self._stop.is_set():
                try:
                    self._queue.put(item, timeout=0.1)
                    break
                except queue.Full:
                    continue

    def step(self):
        if not self.config.async_actor:
            return self._sample()
        if not self._thread.is_alive():
            self._thread.start()
        return self._queue.get()

    def close(self):
        self._stop.set()
        if self._thread is not None and self._thread.is_alive():
            try:
                self._queue.get_nowait()
            except queue.Empty:
                pass
            self._thread.join(timeout=5)
        close_obj(self._task)

    def set_network(self, net):
        self._network = net

    def _transition(self):
        raise NotImplementedError

    def _set_up(self):
        pass
