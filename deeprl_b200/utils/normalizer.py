"""State / reward normalizers with the reference's interface (``deep_rl/utils/normalizer.py``).

``MeanStdNormalizer`` keeps the running moments of baselines' ``RunningMeanStd`` (un-vendored
third-party dependency of the reference, baselines@8e56dd; call sites normalizer.py:36-43):
float64 mean/var/count, Chan pairwise merge of each batch's moments, count initialised to 1e-4.
"""
import numpy as np
import torch


class BaseNormalizer:
    def __init__(self, read_only=False):
        self.read_only = read_only

    def set_read_only(self):
        self.read_only = True

    def unset_read_only(self):
        self.read_only = False

    def state_dict(self):
        return None

    def load_state_dict(self, _):
        return


class RunningMoments:
    """Streaming mean / population variance over axis 0 (float64)."""

    def __init__(self, shape=(), epsilon=1e-4):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, x):
        x = np.asarray(x)
        b_mean, b_var, b_n = x.mean(axis=0), x.var(axis=0), x.shape[0]
        tot = self.count + b_n
        d = b_mean - self.mean
        m2 = self.var * self.count + b_var * b_n + np.square(d) * self.count * b_n / tot
        self.mean = self.mean + d * b_n / tot
        self.var = m2 / tot
        self.count = tot


class MeanStdNormalizer(BaseNormalizer):
    def __init__(self, read_only=False, clip=10.0, epsilon=1e-8):
        super().__init__(read_only)
        self.rms = None
        self.clip = clip
        self.epsilon = epsilon

    def __call__(self, x):
        x = np.asarray(x)
        if self.rms is None:
            self.rms = RunningMoments(shape=(1,) + x.shape[1:])
        if not self.read_only:
            self.rms.update(x)
        z = (x - self.rms.mean) / np.sqrt(self.rms.var + self.epsilon)
        return np.clip(z, -self.clip, self.clip)

    def state_dict(self):
        return {"mean": self.rms.mean, "var": self.rms.var}

    def load_state_dict(self, saved):
        self.rms.mean, self.rms.var = saved["mean"], saved["var"]


class RescaleNormalizer(BaseNormalizer):
    def __init__(self, coef=1.0):
        super().__init__()
        self.coef = coef

    def __call__(self, x):
        if not isinstance(x, torch.Tensor):
            x = np.asarray(x)
        return self.coef * x


class ImageNormalizer(RescaleNormalizer):
    def __init__(self):
        super().__init__(1.0 / 255)


class SignNormalizer(BaseNormalizer):
    def __call__(self, x):
        return np.sign(x)
