#######################################################################
# This file restates an interface of ShangtongZhang/DeepRL, whose     #
# licence asks that the following declaration stay at the top:        #
#                                                                     #
# Copyright (C) 2017 Shangtong Zhang(zhangshangtong.cpp@gmail.com)    #
# Permission given to modify the code as long as you keep this        #
# declaration at the top                                              #
#######################################################################
"""Stateful scalar schedules (reference: ``deep_rl/utils/schedule.py:7-31``).  A
``LinearSchedule`` advances on EVERY call (schedule.py:28-31), so e.g. the PER beta moves once
per gradient update (DQN_agent.py:125)."""


class ConstantSchedule:
    def __init__(self, val):
        self.val = val

    def __call__(self, steps=1):
        return self.val


class LinearSchedule:
    def __init__(self, start, end=None, steps=None):
        if end is None:
            end, steps = start, 1
        self.inc = (end - start) / float(steps)
        self.current, self.end = start, end
        self.bound = min if end > start else max

    def __call__(self, steps=1):
        now = self.current
        self.current = self.bound(now + self.inc * steps, self.end)
        return now
