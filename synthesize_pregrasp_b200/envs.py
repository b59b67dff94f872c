"""gym-style shims with the constructor kwargs and reset()/step()/seed() surface of
envs/*_contact_graph_env.py (plate_contact_graph_env.py:48-72, :176-250, :253-632, :657-659) that the
optimiser and scripts rely on.  step() is ONE pg_env_step_single call: the episode's state (rigid-body world, contact
manifolds, decode carry) stays on the GPU between steps (PgCarry) and the step runs the very code the MPPI batch runs per
rollout, so a single-sample caller sees the same numbers the batch does."""
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
        self.episode = EN.Episode(self.engine)
        self.step_cnt = 0

    def seed(self, seed=None):
        return [seed]

    def reset(self):
        self.step_cnt = 0
        return self._obs(self.episode.reset())

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
        if self.step_cnt >= self.n_steps:
            raise RuntimeError("episode is over: call reset()")
        done = self.step_cnt == self.n_steps - 1                 # plate_env:556: the last step settles and reports the metric
        reward, metric, pose, _ = self.episode.step(a, self.path[self.step_cnt], self.max_forces, last_step=done)
        self.step_cnt += 1
        return self._obs(pose), reward, done, {"metric": metric if done else 0.0}


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
