#######################################################################
# This file restates an interface of ShangtongZhang/DeepRL, whose     #
# licence asks that the following declaration stay at the top:        #
#                                                                     #
# Copyright (C) 2017 Shangtong Zhang(zhangshangtong.cpp@gmail.com)    #
# Permission given to modify the code as long as you keep this        #
# declaration at the top                                              #
#######################################################################
"""Driver loop and small host helpers (reference: ``deep_rl/utils/misc.py:19-84``)."""
import datetime
import time
from pathlib import Path

import numpy as np


def run_steps(agent):
    """misc.py:19-35 -- order per iteration: save, log, eval, stop-check, step, switch_task."""
    config = agent.config
    name = type(agent).__name__
    t0 = time.time()
    while True:
        steps = agent.total_steps
        if config.save_interval and steps % config.save_interval == 0:
            agent.save("data/%s-%s-%d" % (name, config.tag, steps))
        if config.log_interval and steps % config.log_interval == 0:
            agent.logger.info("steps %d, %.2f steps/s" % (steps, config.log_interval / (time.time() - t0)))
            t0 = time.time()
        if config.eval_interval and steps % config.eval_interval == 0:
            agent.eval_episodes()
        if config.max_steps and steps >= config.max_steps:
            agent.close()
            return
        agent.step()
        agent.switch_task()


def get_time_str():
    return datetime.datetime.now().strftime("%y%m%d-%H%M%S")


def mkdir(path):
    Path(path).mkdir(parents=True, exist_ok=True)


def close_obj(obj):
    if hasattr(obj, "close"):
        obj.close()


def random_sample(indices, batch_size):
    """misc.py:55-62: one ``np.random.permutation``; full rows first, then the ragged tail."""
    perm = np.asarray(np.random.permutation(indices))
    n_full = len(perm) // batch_size
    for row in perm[:n_full * batch_size].reshape(n_full, batch_size) if n_full else ():
        yield row
    if len(perm) % batch_size:
        yield perm[n_full * batch_size:]


def is_plain_type(x):
    return isinstance(x, (str, int, float, bool))


def generate_tag(params):
    """misc.py:72-84: builds ``params['tag']`` from the sorted kwargs unless one is given."""
    if "tag" in params:
        return
    game = params.pop("game")
    run = params.pop("run", 0)
    parts = ["%s_%s" % (k, v if is_plain_type(v) else v.__name__) for k, v in sorted(params.items())]
    params.update(tag="%s-%s-run-%d" % (game, "-".join(parts), run), game=game, run=run)
