#######################################################################
# This file restates an interface of ShangtongZhang/DeepRL, whose     #
# licence asks that the following declaration stay at the top:        #
#                                                                     #
# Copyright (C) 2017 Shangtong Zhang(zhangshangtong.cpp@gmail.com)    #
# Permission given to modify the code as long as you keep this        #
# declaration at the top                                              #
#######################################################################
"""Thin logging facade with the reference's surface (``deep_rl/utils/logger.py:17-70``):
``get_logger(tag, log_level)`` -> object with ``info/debug/warning/add_scalar/add_histogram``.
TensorBoard is created lazily and only if importable; ``tag=None`` writes no files."""
import logging

import numpy as np
import torch

from .misc import get_time_str, mkdir

logging.basicConfig(format="%(asctime)s - %(name)s - %(levelname)s: %(message)s")


def get_logger(tag="default", log_level=0):
    logger = logging.getLogger()
    logger.setLevel(logging.INFO)
    if tag is not None:
        mkdir("./log")
        fh = logging.FileHandler("./log/%s-%s.txt" % (tag, get_time_str()))
        fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s: %(message)s"))
        fh.setLevel(logging.INFO)
        logger.addHandler(fh)
    log_dir = None if tag is None else "./tf_log/logger-%s-%s" % (tag, get_time_str())
    return Logger(logger, log_dir, log_level)


class Logger:
    def __init__(self, vanilla_logger, log_dir, log_level=0):
        self.log_level = log_level
        self.writer = None
        self.log_dir = log_dir
        self.all_steps = {}
        if vanilla_logger is not None:
            self.info, self.debug, self.warning = vanilla_logger.info, vanilla_logger.debug, vanilla_logger.warning

    def lazy_init_writer(self):
        if self.writer is None and self.log_dir is not None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(self.log_dir)
            except Exception:                                            # tensorboard absent: scalars are dropped
                self.log_dir = None

    @staticmethod
    def to_numpy(v):
        return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v

    def get_step(self, tag):
        step = self.all_steps.get(tag, 0)
        self.all_steps[tag] = step + 1
        return step

    def add_scalar(self, tag, value, step=None, log_level=0):
        if log_level > self.log_level:
            return
        self.lazy_init_writer()
        if self.writer is None:
            return
        value = self.to_numpy(value)
        if step is None:
            step = self.get_step(tag)
        self.writer.add_scalar(tag, float(np.asarray(value).reshape(-1)[0]), step)

    def add_histogram(self, tag, values, step=None, log_level=0):
        if log_level > self.log_level:
            return
        self.lazy_init_writer()
        if self.writer is None:
            return
        if step is None:
            step = self.get_step(tag)
        self.writer.add_histogram(tag, self.to_numpy(values), step)
