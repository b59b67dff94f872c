"""gym-style shims with the constructor kwargs and reset()/step()/seed() surface of
envs/*_contact_graph_env.py (plate_contact_graph_env.py:48-72, :176-250, :253-632, :657-659) that the
optimiser and scripts rely on.  step() replays the action history through the batched kernel with K=1, so a
single-sample caller sees the same numbers the MPPI batch does."""
from __future__ import annotations

import types

import numpy as np

from . import engine as EN
from . import envspec as ES


class _ContactGraphEnv:
    ENV_NAME = None

    def __init__(self, render=True, init_noise=True, control_skip=50, active_finger_tips=[0, 1], num_interp_f=5,
                 opt_time=False, last_fins=None, init_obj_pose=None, task=None, train=True, steps=3,
                 observe_last_action=False, path=None, dex_path=None, sc_path=None, opt_dict=None, mode="decoder",
                 sigma=0.5, device="cuda", showImage=False, has_distance_field=False, max_forces=80,
                 add_physics=False, validate=False, **_ignored):
        if list(active_finger_tips) != [0, 1] or num_interp_f != ES.NUM_INTERP_F or int(control_skip) != ES.CONTROL_SKIP:
            raise ValueError("the GPU path is built for active_finger_tips=[0,1], num_interp_f=7, control_skip=50 "
                             "(model_optimize.py:75-81)")
        if last_fins is not None or init_obj_pose is not None or opt_time or add_physics or not train:
            raise NotImplementedError("only the train-mode MPPI configuration of model_optimize.py is on the GPU path")
        self.n_steps, self.max_forces, self.validate = steps, max_forces, validate
        self.path = np.asarray(path if path is not None else [0] * (steps + 1), np.int32)
        dev = 0 if device in ("cuda", "cpu") else int(str(device).split(":")[-1])
        self.engine = EN.RolloutEngine(self.ENV_NAME, device=dev, validate=validate)
        self.action_dim = ES.ACTION_DIM
        self.action_space = types.SimpleNamespace(shape=(self.action_dim,), low=-np.ones(self.action_dim),
                                                  high=np.ones(self.action_dim))
        self._actions = []

    def seed(self, seed=None):
        return [seed]

    def reset(self):
        self._actions = []
        return self._obs(np.r_[self.engine.spec.obj_pos, self.engine.spec.obj_quat])

    def _obs(self, pose):
        # get_extended_observation plate_env:634-655: 8 corners of a 0.4 x 0.4 x 0.1 box at the object pose
        dim = np.array([0.4, 0.4, 0.1]) / 2
        sg = np.array([[1, 1, 1], [-1, 1, 1], [1, -1, 1], [1, 1, -1], [-1, -1, 1], [-1, 1, -1], [1, -1, -1], [-1, -1, -1]])
        x, y, z, w = pose[3:]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        return list(((sg * dim) @ R.T + pose[:3]).ravel())

    def step(self, a, rollout_verify=False):
        self._actions.append(np.asarray(a, np.float64))
        n = len(self._actions)
        if n > self.n_steps:
            raise RuntimeError("episode is over: call reset()")
        u = np.stack(self._actions)
        # costs accumulate per step: reward of this step = -(J_n - J_{n-1})
        out = self.engine.physics(u, None, self.path, self.max_forces, K=1)
        logit = self.engine.cost(out["pose"][0, n - 1:n], out["fins"][0, n - 1:n])
        reward = 100.0 * float(logit.item())
        done = n == self.n_steps
        metric = 0.0
        if done:
            full = self.engine.rollout(u, None, self.path, self.max_forces, K=1, want_metric=True)
            metric = float(full[1].item())
            # the last step's reward is evaluated after the settle substeps (plate_env:556-612)
            J = float(full[0].item())
            prev = 0.0
            for j in range(n - 1):
                lj = self.engine.cost(out["pose"][0, j:j + 1], out["fins"][0, j:j + 1])
                prev += -100.0 * float(lj.item())
            reward = -(J - prev)
        return self._obs(out["pose"][0, n - 1].cpu().numpy()), reward, done, {"metric": metric}


def _make(name, env_name):
    return type(name, (_ContactGraphEnv,), {"ENV_NAME": env_name})


PlateBulletEnv = _make("PlateBulletEnv", "plate")
BookShelfBulletEnv = _make("BookShelfBulletEnv", "bookshelf")
FoodboxBulletEnv = _make("FoodboxBulletEnv", "foodbox")
LaptopBulletEnv = _make("LaptopBulletEnv", "laptop")
KeyboardBulletEnv = _make("KeyboardBulletEnv", "keyboard")
CardboardBulletEnv = _make("CardboardBulletEnv", "cardboard")
WaterbottleBulletEnv = _make("WaterbottleBulletEnv", "waterbottle")

# model_optimize.py:22-34
envs_dict = {"bookshelf": BookShelfBulletEnv, "foodbox": FoodboxBulletEnv, "wallbox": LaptopBulletEnv,
             "plate": PlateBulletEnv, "waterbottle": WaterbottleBulletEnv, "cardboard": CardboardBulletEnv,
             "keyboard": KeyboardBulletEnv}
