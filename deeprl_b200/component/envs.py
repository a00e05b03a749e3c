"""Vectorised rollout collector: ``Task`` + ``DummyVecEnv`` + ``LazyFrames`` with the reference's
interface (``deep_rl/component/envs.py:92-189``), over environments that live on host CPU.

Environment back-ends, resolved by ``make_env``:

* built-in ids that need no third-party package: ``CartPole-v0`` / ``CartPole-v1`` (classic
  cart-pole dynamics, Euler integration, tau = 0.02 s) and the synthetic shapes BASELINE.json is
  quoted on -- ``SyntheticAtari-v0`` (84x84 uint8 frames, frame-stack 4 delivered as
  ``LazyFrames``, 4 actions; ``SyntheticAtari-A<k>-v0`` for k actions) and
  ``SyntheticCheetah-v0`` (17-dim float32 observation, 6-dim Box action in [-1, 1]);
* anything else is handed to ``gym`` / ``gymnasium`` if one is importable (neither is in this
  image); Atari / dm_control glue of the reference (envs.py:27-55) is out of scope.

``DummyVecEnv.step_wait`` keeps the reference contract (envs.py:136-144): envs are stepped
serially, auto-reset on ``done``, observations come back as a TUPLE (not stacked), rewards and
dones as numpy arrays, infos as a tuple of dicts carrying ``episodic_return``.
"""
import math
import re

import numpy as np

from ..utils.misc import mkdir
from ..utils.torch_utils import random_seed


# ------------------------------------------------------------------------------------ spaces
class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.shape = tuple(shape)
        self.dtype = dtype
        self.low = np.full(self.shape, low, dtype=np.float32) if np.isscalar(low) else np.asarray(low, np.float32)
        self.high = np.full(self.shape, high, dtype=np.float32) if np.isscalar(high) else np.asarray(high, np.float32)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


def _is_discrete(space):
    return hasattr(space, "n")


# ------------------------------------------------------------------------------------ frames
class LazyFrames:
    """Frame-stack view that stores each frame once (envs.py:92-113): converts to one array only
    when asked (``np.asarray(obs)``, ``obs[i]``, ``len(obs)``)."""

    def __init__(self, frames):
        self._frames = frames

    def __array__(self, dtype=None, copy=None):
        out = np.concatenate(self._frames, axis=0)
        return out if dtype is None else out.astype(dtype)

    def __len__(self):
        return len(self.__array__())

    def __getitem__(self, i):
        return self.__array__()[i]


# ------------------------------------------------------------------------------------ built-in envs
class _Env:
    observation_space = None
    action_space = None

    def seed(self, seed=None):
        self.rng = np.random.RandomState(seed)

    def close(self):
        pass


class CartPoleEnv(_Env):
    """Cart-pole balancing (Barto, Sutton & Anderson 1983) with the usual constants: gravity 9.8,
    cart 1.0 kg, pole 0.1 kg / half-length 0.5 m, force 10 N, Euler step 0.02 s; episode ends at
    |x| > 2.4, |theta| > 12 deg or after ``max_steps`` (200 for -v0, 500 for -v1); reward 1/step."""

    def __init__(self, max_steps=200):
        hi = np.array([4.8, np.finfo(np.float32).max, 24 * math.pi / 180, np.finfo(np.float32).max], np.float32)
        self.observation_space = Box(-hi, hi)
        self.action_space = Discrete(2)
        self.max_steps = max_steps
        self.seed(None)

    def reset(self):
        self.x = self.rng.uniform(-0.05, 0.05, size=4)
        self.t = 0
        return self.x.astype(np.float32)

    def step(self, action):
        x, xd, th, thd = self.x
        f = 10.0 if int(action) == 1 else -10.0
        c, s = math.cos(th), math.sin(th)
        tmp = (f + 0.05 * thd * thd * s) / 1.1
        tha = (9.8 * s - c * tmp) / (0.5 * (4.0 / 3.0 - 0.1 * c * c / 1.1))
        xa = tmp - 0.05 * tha * c / 1.1
        self.x = np.array([x + 0.02 * xd, xd + 0.02 * xa, th + 0.02 * thd, thd + 0.02 * tha])
        self.t += 1
        done = bool(abs(self.x[0]) > 2.4 or abs(self.x[2]) > 12 * math.pi / 180 or self.t >= self.max_steps)
        return self.x.astype(np.float32), 1.0, done, {}


class SyntheticAtariEnv(_Env):
    """Atari-SHAPED stream: uint8 84x84 frames, reward in {-1,0,+1} with P=(.05,.9,.05),
    episode end ~ Bernoulli(1/1000).  Observation = ``LazyFrames`` of the last 4 (1,84,84) frames,
    exactly what the reference's FrameStack delivers (envs.py:116-123)."""

    def __init__(self, num_actions=4, p_done=1e-3):
        self.observation_space = Box(0, 255, (4, 84, 84), np.uint8)
        self.action_space = Discrete(num_actions)
        self.p_done = p_done
        self.seed(None)

    def _frame(self):
        return self.rng.randint(0, 256, size=(1, 84, 84), dtype=np.uint8)

    def reset(self):
        first = self._frame()
        self.frames = [first] * 4
        return LazyFrames(list(self.frames))

    def step(self, action):
        self.frames = self.frames[1:] + [self._frame()]
        u = self.rng.rand()
        reward = -1.0 if u < 0.05 else (1.0 if u > 0.95 else 0.0)
        done = bool(self.rng.rand() < self.p_done)
        return LazyFrames(list(self.frames)), reward, done, {}


class SyntheticCheetahEnv(_Env):
    """HalfCheetah-SHAPED stream: 17-dim N(0,1) float32 observation, 6-dim action in [-1,1],
    reward ~ N(0,1) - 0.1*|a|^2, episode end ~ Bernoulli(1/1000)."""

    def __init__(self, obs_dim=17, act_dim=6, p_done=1e-3):
        self.observation_space = Box(-np.inf, np.inf, (obs_dim,))
        self.action_space = Box(-1.0, 1.0, (act_dim,))
        self.p_done = p_done
        self.seed(None)

    def _obs(self):
        return self.rng.randn(*self.observation_space.shape).astype(np.float32)

    def reset(self):
        return self._obs()

    def step(self, action):
        a = np.asarray(action, dtype=np.float64)
        reward = float(self.rng.randn() - 0.1 * np.square(a).sum())
        done = bool(self.rng.rand() < self.p_done)
        return self._obs(), reward, done, {}


class OriginalReturnWrapper:
    """Adds ``info['episodic_return']`` (None until the episode ends) -- envs.py:58-74."""

    def __init__(self, env):
        self.env = env
        self.observation_space, self.action_space = env.observation_space, env.action_space
        self.total_rewards = 0

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.total_rewards += reward
        info = dict(info)
        info["episodic_return"] = self.total_rewards if done else None
        if done:
            self.total_rewards = 0
        return obs, reward, done, info

    def reset(self):
        return self.env.reset()

    def __getattr__(self, name):
        return getattr(self.env, name)


def _builtin(env_id):
    if env_id in ("CartPole-v0", "CartPole-v1"):
        return CartPoleEnv(200 if env_id.endswith("v0") else 500)
    m = re.fullmatch(r"SyntheticAtari(?:-A(\d+))?-v0", env_id)
    if m:
        return SyntheticAtariEnv(int(m.group(1) or 4))
    if env_id == "SyntheticCheetah-v0":
        return SyntheticCheetahEnv()
    return None


def make_env(env_id, seed, rank, episode_life=True):
    """Thunk factory (envs.py:27-55): seeds the process RNGs, builds the env, seeds it with
    ``seed + rank`` and wraps it so infos carry the un-normalised episodic return."""

    def _thunk():
        random_seed(seed)
        env = _builtin(env_id)
        if env is None:
            try:
                import gym
            except ImportError:
                try:
                    import gymnasium as gym
                except ImportError:
                    raise RuntimeError("unknown built-in env id %r and neither gym nor gymnasium is installed" % env_id)
            env = gym.make(env_id)
        env.seed(seed + rank)
        return OriginalReturnWrapper(env)

    return _thunk


# ------------------------------------------------------------------------------------ vector env + Task
class DummyVecEnv:
    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.observation_space = self.envs[0].observation_space
        self.action_space = self.envs[0].action_space
        self.actions = None

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        obs, rew, done, info = [], [], [], []
        for env, a in zip(self.envs, self.actions):
            o, r, d, i = env.step(a)
            if d:
                o = env.reset()
            obs.append(o), rew.append(r), done.append(d), info.append(i)
        return tuple(obs), np.asarray(rew), np.asarray(done), tuple(info)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def reset(self):
        return [env.reset() for env in self.envs]

    def close(self):
        return


class Task:
    """envs.py:153-189.  ``single_process=False`` selected baselines' SubprocVecEnv in the
    reference; that third-party worker pool is out of scope, so it is refused loudly."""

    def __init__(self, name, num_envs=1, single_process=True, log_dir=None, episode_life=True, seed=None):
        if seed is None:
            seed = np.random.randint(int(1e9))
        if log_dir is not None:
            mkdir(log_dir)
        if not single_process:
            raise NotImplementedError("SubprocVecEnv (baselines) is not part of this build; use single_process=True")
        self.env = DummyVecEnv([make_env(name, seed, i, episode_life) for i in range(num_envs)])
        self.name = name
        self.observation_space = self.env.observation_space
        self.state_dim = int(np.prod(self.observation_space.shape))
        self.action_space = self.env.action_space
        if _is_discrete(self.action_space):
            self.action_dim = self.action_space.n
        else:
            self.action_dim = self.action_space.shape[0]

    def reset(self):
        return self.env.reset()

    def step(self, actions):
        if not _is_discrete(self.action_space):
            actions = np.clip(actions, self.action_space.low, self.action_space.high)
        return self.env.step(actions)

    def close(self):
        self.env.close()
