"""TEST INFRASTRUCTURE -- restatement of ``baselines.common.running_mean_std.RunningMeanStd``.

The reference depends on openai/baselines pinned at commit 8e56dd (reference Dockerfile:48,
setup.py:4); that dependency is not vendored under /root/reference and is not installable
here.  Its call sites on the hot path are ``deep_rl/utils/normalizer.py:36-43``
(``MeanStdNormalizer.__call__``: ``RunningMeanStd(shape=(1,)+x.shape[1:])`` then ``update(x)``).

Published algorithm (Chan, Golub & LeVeque pairwise update of mean / M2, all float64):
state (mean=0, var=1, count=1e-4); ``update(x)`` merges the batch moments
(mean over axis 0, *population* variance over axis 0, n = x.shape[0]).
Parity is unpinned by the reference (it has no tests); it is pinned here by the known-answer
test tests/test_oracle_golden.py::test_running_mean_std_closed_form.
"""
import numpy as np


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, dtype=np.float64)
        self.var = np.ones(shape, dtype=np.float64)
        self.count = epsilon

    def update(self, x):
        x = np.asarray(x)
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        n_a, n_b = self.count, batch_count
        n = n_a + n_b
        delta = batch_mean - self.mean
        mean = self.mean + delta * n_b / n
        m2 = self.var * n_a + batch_var * n_b + np.square(delta) * n_a * n_b / n
        self.mean, self.var, self.count = mean, m2 / n, n
