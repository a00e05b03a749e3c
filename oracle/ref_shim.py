"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference (ShangtongZhang/DeepRL)
from /root/reference inside THIS container so that golden vectors can be generated from it.

Nothing here is shipped or imported by the product (``deeprl_b200``).  ``/root/reference``
does not exist on the GPU box, so only ``tests/golden/make_golden.py`` (run here, by hand)
uses this module; the fixtures it writes are what travels.

How the reference is made importable on Python 3.12 / torch 2.11 without copying its sources:

* ``replay.py:205,210`` use ``async`` as an identifier (a keyword since 3.7).  A meta-path
  finder compiles the reference modules from their on-disk text and, for that one file,
  renames the token ``async`` -> ``async_`` in memory.
* ``misc.py:14`` imports ``collections.Sequence`` (moved to ``collections.abc`` in 3.10):
  aliased before import.
* ``gym``, ``baselines`` and ``skimage`` are absent: minimal stub modules provide the handful
  of names ``envs.py:8-16``, ``normalizer.py:8`` and ``BaseAgent.py:12`` import.
  ``RunningMeanStd`` is restated from baselines@8e56dd's published algorithm
  (Chan et al. parallel variance merge) -- see oracle/running_mean_std.py.
"""
from __future__ import annotations

import collections
import collections.abc
import importlib.abc
import importlib.util
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("B2RL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deep_rl"))


class _Space:
    pass


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=None):
        import numpy as np
        if shape is None:
            low = np.asarray(low)
            shape = low.shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float32), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float32), self.shape).copy()
        self.dtype = dtype


class _Discrete(_Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class _Wrapper:
    def __init__(self, env):
        self.env = env
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)

    def __getattr__(self, name):
        return getattr(self.env, name)


class _VecEnv:
    """Shape of baselines.common.vec_env.VecEnv that envs.py:126-149 relies on."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


def _install_stubs():
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence  # misc.py:14

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "gym" not in sys.modules:
        envs = mod("gym.envs")
        box = mod("gym.spaces.box", Box=_Box)
        disc = mod("gym.spaces.discrete", Discrete=_Discrete)
        spaces = mod("gym.spaces", box=box, discrete=disc, Box=_Box, Discrete=_Discrete)

        def _no_make(*a, **k):
            raise RuntimeError("gym.make is not available: inject a synthetic Task via config.task_fn")

        mod("gym", Wrapper=_Wrapper, ObservationWrapper=_Wrapper, envs=envs, spaces=spaces, make=_no_make)

    if "baselines" not in sys.modules:
        from oracle.running_mean_std import RunningMeanStd

        def _absent(*a, **k):
            raise RuntimeError("baselines atari wrappers are not available in this container")

        class _FrameStack(_Wrapper):
            def __init__(self, env, k):
                _Wrapper.__init__(self, env)
                self.k = k
                self.frames = collections.deque([], maxlen=k)

        aw = mod("baselines.common.atari_wrappers", make_atari=_absent, wrap_deepmind=_absent,
                 FrameStack=_FrameStack)
        sve = mod("baselines.common.vec_env.subproc_vec_env", SubprocVecEnv=_absent, VecEnv=_VecEnv)
        ve = mod("baselines.common.vec_env", subproc_vec_env=sve, VecEnv=_VecEnv)
        rms = mod("baselines.common.running_mean_std", RunningMeanStd=RunningMeanStd)
        common = mod("baselines.common", atari_wrappers=aw, vec_env=ve, running_mean_std=rms)
        mod("baselines", common=common)

    if "skimage" not in sys.modules:
        io = mod("skimage.io", imsave=lambda *a, **k: None)
        mod("skimage", io=io)


class _PatchedLoader(importlib.abc.Loader):
    def __init__(self, path, is_pkg):
        self.path = path
        self.is_pkg = is_pkg

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(self.path, "r", encoding="utf-8") as f:
            src = f.read()
        if self.path.endswith(os.path.join("component", "replay.py")):
            # the ONLY edit: keyword-safe spelling of the kwarg at replay.py:205 and its use at :210
            src = re.sub(r"\basync\b", "async_", src)
        code = compile(src, self.path, "exec")
        exec(code, module.__dict__)


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != "deep_rl" and not fullname.startswith("deep_rl."):
            return None
        rel = fullname.split(".")
        base = os.path.join(REFERENCE_ROOT, *rel)
        if os.path.isdir(base):
            file = os.path.join(base, "__init__.py")
            spec = importlib.util.spec_from_loader(fullname, _PatchedLoader(file, True), origin=file, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        file = base + ".py"
        if os.path.isfile(file):
            return importlib.util.spec_from_loader(fullname, _PatchedLoader(file, False), origin=file)
        return None


_installed = False


def import_reference():
    """Return the imported reference package ``deep_rl`` (from /root/reference, in memory)."""
    global _installed
    if not reference_available():
        raise RuntimeError("reference tree %s is not present (it exists only in the build container)" % REFERENCE_ROOT)
    if not _installed:
        import warnings
        warnings.filterwarnings("ignore", category=SyntaxWarning)
        warnings.filterwarnings("ignore", category=DeprecationWarning)
        _install_stubs()
        sys.meta_path.insert(0, _RefFinder())
        _installed = True
    import deep_rl  # noqa: F401  (resolved by _RefFinder)
    return deep_rl
