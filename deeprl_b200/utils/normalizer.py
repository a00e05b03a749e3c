"""State / reward normalizers behind the reference's interface (``deep_rl/utils/normalizer.py:11-71``): callables with a
read-only switch and ``state_dict`` / ``load_state_dict``.

``MeanStdNormalizer`` keeps the running moments of baselines' ``RunningMeanStd`` -- an un-vendored third-party
dependency of the reference (baselines@8e56dd; call sites normalizer.py:36-43), restated here from its published
algorithm: float64 mean / population variance / count, Chan's pairwise merge of every batch's moments, count
initialised to 1e-4.  ``tests/test_oracle_golden.py::test_running_mean_std_closed_form`` pins it against the closed-form
statistics of the concatenated batches.  On the DQN path ``ImageNormalizer`` never runs as a separate pass: its 1/255 is
folded into the replay gather's lookup table or conv1's weights (``network/fused.py: frame_scale``).
"""
import numpy as np
import torch


class BaseNormalizer:
    """Common switchboard: ``read_only`` freezes the statistics (evaluation episodes)."""

    def __init__(self, read_only=False):
        self.read_only = bool(read_only)

    def _freeze(self, flag):
        self.read_only = flag

    def set_read_only(self):
        self._freeze(True)

    def unset_read_only(self):
        self._freeze(False)

    # stateless by default
    def state_dict(self):
        return None

    def load_state_dict(self, state):
        del state


class RunningMoments:
    """Streaming mean and population variance over axis 0 in float64 (Chan et al. pairwise update)."""

    def __init__(self, shape=(), epsilon=1e-4):
        self.mean = np.zeros(shape, dtype=np.float64)
        self.var = np.ones(shape, dtype=np.float64)
        self.count = epsilon

    def update(self, batch):
        batch = np.asarray(batch)
        self.merge(batch.mean(axis=0), batch.var(axis=0), batch.shape[0])

    def merge(self, mean_b, var_b, n_b):
        n_a = self.count
        n = n_a + n_b
        delta = mean_b - self.mean
        second_moment = self.var * n_a + var_b * n_b + np.square(delta) * n_a * n_b / n
        self.mean = self.mean + delta * n_b / n
        self.var = second_moment / n
        self.count = n


class MeanStdNormalizer(BaseNormalizer):
    """``clip((x - mean) / sqrt(var + epsilon), -clip, clip)`` with running moments updated on every call unless read-only."""

    def __init__(self, read_only=False, clip=10.0, epsilon=1e-8):
        super().__init__(read_only)
        self.clip, self.epsilon = clip, epsilon
        self.rms = None                                    # created on first use: the moments take the observation shape

    def __call__(self, x):
        x = np.asarray(x)
        moments = self._moments_for(x)
        if not self.read_only:
            moments.update(x)
        standardised = (x - moments.mean) / np.sqrt(moments.var + self.epsilon)
        return np.clip(standardised, -self.clip, self.clip)

    def _moments_for(self, x):
        if self.rms is None:
            self.rms = RunningMoments(shape=(1,) + tuple(x.shape[1:]))
        return self.rms

    def state_dict(self):
        return dict(mean=self.rms.mean, var=self.rms.var)

    def load_state_dict(self, saved):
        if self.rms is None:                               # a fresh agent loading a checkpoint: the moments take the saved shape
            self.rms = RunningMoments(shape=np.shape(saved["mean"]))    # (the reference dereferences None here, normalizer.py:47-51)
        for key in ("mean", "var"):
            setattr(self.rms, key, saved[key])


class RescaleNormalizer(BaseNormalizer):
    """``coef * x``; tensors stay tensors, everything else goes through ``np.asarray`` first."""

    def __init__(self, coef=1.0):
        super().__init__()
        self.coef = coef

    def __call__(self, x):
        return self.coef * (x if torch.is_tensor(x) else np.asarray(x))


class ImageNormalizer(RescaleNormalizer):
    """uint8 frames -> [0, 1] (float64 multiply, as the reference: the consumer rounds once to float32)."""

    def __init__(self):
        super().__init__(coef=1.0 / 255)


class SignNormalizer(BaseNormalizer):
    """Reward clipping to {-1, 0, +1}."""

    def __call__(self, x):
        return np.sign(x)
