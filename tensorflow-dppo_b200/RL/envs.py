"""Minimal environment layer.  gym is not installed in this image and the simulator is outside
the accelerated path (Worker rollout loop is out of scope); the Worker only needs
``observation_space.shape``, ``action_space.shape`` and ``reset/step`` (Worker.py:10,35,50).
`make` returns a vectorised env of N independent copies so that one Worker replica owns N
rollouts (the [T,N] buffer layout of the engine)."""
import numpy as np


class Box:
    """gym.spaces.Box surface used by make_pdtype (distributions.py:233-235)."""
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = low, high, tuple(shape)


class Discrete:
    """gym.spaces.Discrete surface used by make_pdtype (distributions.py:236-237) and Worker.act (Worker.py:150)."""
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class CartPoleVec:
    """N classic cart-poles (the committed default game, main.py:13 'CartPole-v0'): obs = [x, xdot, th, thdot],
    two discrete actions (push left / right), reward 1 per step, episode ends when the pole falls, the cart leaves
    the track, or after 200 steps."""
    gravity, masscart, masspole, length, force_mag, tau, horizon = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02, 200
    th_limit, x_limit = 12 * 2 * np.pi / 360, 2.4

    def __init__(self, num_envs=1, seed=0):
        self.n = num_envs
        self.rng = np.random.default_rng(seed)
        self.observation_space = Box(-np.inf, np.inf, (4,))
        self.action_space = Discrete(2)
        self.state = np.zeros((self.n, 4)); self.t = np.zeros(self.n, np.int64)

    def reset(self, mask=None):
        m = np.ones(self.n, bool) if mask is None else mask
        self.state[m] = self.rng.uniform(-0.05, 0.05, (int(m.sum()), 4))
        self.t[m] = 0
        return self.state.astype(np.float32)

    def step(self, action):
        a = np.asarray(action).reshape(self.n).astype(np.int64)
        x, xd, th, thd = self.state.T
        force = np.where(a == 1, self.force_mag, -self.force_mag)
        total_mass = self.masscart + self.masspole; pml = self.masspole * self.length
        temp = (force + pml * thd ** 2 * np.sin(th)) / total_mass
        thacc = (self.gravity * np.sin(th) - np.cos(th) * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * np.cos(th) ** 2 / total_mass))
        xacc = temp - pml * thacc * np.cos(th) / total_mass
        self.state = np.stack([x + self.tau * xd, xd + self.tau * xacc, th + self.tau * thd, thd + self.tau * thacc], 1)
        self.t += 1
        done = (np.abs(self.state[:, 0]) > self.x_limit) | (np.abs(self.state[:, 2]) > self.th_limit) | (self.t >= self.horizon)
        return self.state.astype(np.float32), np.ones(self.n, np.float32), done, {}


class PendulumVec:
    """N classic torque-limited pendulums (obs = [cos th, sin th, thdot], 1-D action, 200-step episodes)."""
    max_speed, max_torque, dt, g, m, l, horizon = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0, 200

    def __init__(self, num_envs=1, seed=0):
        self.n = num_envs
        self.rng = np.random.default_rng(seed)
        self.observation_space = Box(-8.0, 8.0, (3,))
        self.action_space = Box(-2.0, 2.0, (1,))
        self.th = np.zeros(self.n); self.thdot = np.zeros(self.n); self.t = np.zeros(self.n, np.int64)

    def _obs(self):
        return np.stack([np.cos(self.th), np.sin(self.th), self.thdot], 1).astype(np.float32)

    def reset(self, mask=None):
        m = np.ones(self.n, bool) if mask is None else mask
        k = int(m.sum())
        self.th[m] = self.rng.uniform(-np.pi, np.pi, k)
        self.thdot[m] = self.rng.uniform(-1, 1, k)
        self.t[m] = 0
        return self._obs()

    def step(self, action):
        u = np.clip(np.asarray(action, np.float64).reshape(self.n), -self.max_torque, self.max_torque)
        ang = ((self.th + np.pi) % (2 * np.pi)) - np.pi
        cost = ang ** 2 + 0.1 * self.thdot ** 2 + 0.001 * u ** 2
        self.thdot = np.clip(self.thdot + (3 * self.g / (2 * self.l) * np.sin(self.th)
                                           + 3.0 / (self.m * self.l ** 2) * u) * self.dt,
                             -self.max_speed, self.max_speed)
        self.th = self.th + self.thdot * self.dt
        self.t += 1
        done = self.t >= self.horizon
        return self._obs(), (-cost).astype(np.float32), done, {}


class SyntheticVec:
    """Shape-only env for Humanoid/HalfCheetah-sized plumbing runs: random-walk observations."""
    def __init__(self, obs_dim, act_dim, num_envs=1, seed=0, horizon=200):
        self.n, self.horizon = num_envs, horizon
        self.rng = np.random.default_rng(seed)
        self.observation_space = Box(-np.inf, np.inf, (obs_dim,))
        self.action_space = Box(-1.0, 1.0, (act_dim,))
        self.x = np.zeros((num_envs, obs_dim), np.float32); self.t = np.zeros(num_envs, np.int64)

    def reset(self, mask=None):
        m = np.ones(self.n, bool) if mask is None else mask
        self.x[m] = self.rng.standard_normal((int(m.sum()), self.x.shape[1])).astype(np.float32)
        self.t[m] = 0
        return self.x.copy()

    def step(self, action):
        a = np.asarray(action, np.float32).reshape(self.n, -1)
        self.x = (0.95 * self.x + 0.1 * self.rng.standard_normal(self.x.shape)).astype(np.float32)
        self.x[:, : a.shape[1]] += 0.05 * a
        self.t += 1
        r = -np.square(self.x[:, : a.shape[1]]).mean(1)
        return self.x.copy(), r.astype(np.float32), self.t >= self.horizon, {}


_SHAPES = {"HalfCheetah": (17, 6), "Humanoid": (376, 17)}


def make(game, num_envs=1, seed=0):
    """gym.make stand-in (Worker.py:10, Chief.py:10)."""
    if game.startswith("Pendulum"):
        return PendulumVec(num_envs, seed)
    if game.startswith("CartPole"):
        return CartPoleVec(num_envs, seed)
    for k, (o, a) in _SHAPES.items():
        if game.startswith(k):
            return SyntheticVec(o, a, num_envs, seed)
    raise NotImplementedError(f"no simulator for {game!r} in this image (gym is not installed)")
