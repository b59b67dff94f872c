"""Host-side mirror of the reference's optimisation driver, model_optimize.py:36-123: for every contact-state path run
`--runs` MPPI optimisations of `--iters` iterations with 12 x 30 (12 x 15 with --validate) samples, keep the best run, save
its reward logs and controls.

Same argument names and the same output files as the reference (`data/rewards/optimal_<exp>_<F>.npy`, `run_<exp>_<F>.npy`,
`data/traj/<exp>_<F>.npy`, `data/log/<exp>.txt`).  Two schedules:

  * default: as the reference, one `StochTrajOptimizer` per path, its runs one after the other (model_optimize.py:74-98);
  * `--batched`: all paths x runs advance in lock-step as ONE multi-problem launch per MPPI iteration
    (`MultiStochTrajOptimizer` / `pg_rollout_multi`) -- the reference's early exit after the first validated path does not
    apply (validation is not on this path), so every path is optimised.

Not mirrored (out of scope): the `model_test.py` replay that follows each path (:110-113) and the
kinematic `validate_path` screening (:114-120).  Without --validate the reference as committed has no path list (its `paths`
is only defined under `args.validate`, :63-64); here that case runs the single dummy path `[0, 0, ...]` the recordings use
(model_test.py:86-87).

    python -m synthesize_pregrasp_b200.model_optimize --env plate --max_force 20 --has_distance_field --exp_name plate
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from .envs import envs_dict
from .stoch_traj_opt import MultiStochTrajOptimizer, StochTrajOptimizer

MAX_FORCE_DEFAULT = 80.0           # model/param.py:28 MAX_FORCE
SEED = 12367134                    # model_optimize.py:77


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--exp_name", type=str, default="u_opt_0-10_tp4")
    p.add_argument("--env", type=str, default="foodbox", choices=sorted(envs_dict))
    p.add_argument("--iters", type=int, default=20)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--runs", type=int, default=3)
    p.add_argument("--max_force", type=float, default=-1)
    p.add_argument("--has_distance_field", action="store_true", default=False)
    p.add_argument("--validate", action="store_true", default=False)
    # accepted for command-line compatibility; the score network and its weights are fixed in the env constants
    p.add_argument("--mode", type=str, default="only_score")
    p.add_argument("--name_score", type=str, default=None)
    p.add_argument("--name_epoch", type=str, default="2980")
    p.add_argument("--render", action="store_true", default=False)
    # additions
    p.add_argument("--batched", action="store_true", help="all paths x runs as one multi-problem launch per iteration")
    p.add_argument("--max_paths", type=int, default=10, help="paths taken from paths_id.npy (the reference file holds 10)")
    p.add_argument("--out_dir", type=str, default="data")
    p.add_argument("--no_record", action="store_true", help="skip the model_test.py replay / recording that follows every path (model_optimize.py:107-110)")
    return p


def path_list(env_name, steps, validate, max_paths):
    """model_optimize.py:63-64 (`paths_id.npy` under --validate); the dummy path otherwise."""
    from .envspec import make_spec
    spec = make_spec({"wallbox": "laptop"}.get(env_name, env_name))
    if validate:
        if spec.paths_id is None:
            raise ValueError(f"no paths_id.npy for env {env_name}")
        return [list(map(int, p)) for p in np.asarray(spec.paths_id)[:max_paths]]
    return [[0] * (steps + 1)]


def optimiser_kwargs(args, max_force):
    """The keyword set model_optimize.py:76-83 hands to StochTrajOptimizer (the ones that reach the env)."""
    return dict(sigma=0.8, TimeSteps=args.steps, Iterations=args.iters, Num_processes=12,
                Traj_per_process=15 if args.validate else 30, active_finger_tips=[0, 1], num_interp_f=7, opt_time=False,
                mode=args.mode, last_fins=None, init_obj_pose=None, steps=args.steps,
                has_distance_field=args.has_distance_field, max_forces=max_force, validate=args.validate)


def _save(out_dir, exp_name, max_force, Jopt_log, r_log, uopt):
    for sub in ("rewards", "traj", "log"):
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    np.save(os.path.join(out_dir, "rewards", f"optimal_{exp_name}_{max_force}.npy"), Jopt_log)
    if r_log is not None:
        np.save(os.path.join(out_dir, "rewards", f"run_{exp_name}_{max_force}.npy"), r_log)
    np.save(os.path.join(out_dir, "traj", f"{exp_name}_{max_force}.npy"), uopt)


def run(args):
    """-> list of dicts (one per path): path, mean / best reward over the runs, best controls, its logs."""
    max_force = MAX_FORCE_DEFAULT if args.max_force == -1 else args.max_force
    if args.render:
        raise NotImplementedError("rendering is out of scope of the GPU rollout path")
    paths = path_list(args.env, args.steps, args.validate, args.max_paths)
    kw = optimiser_kwargs(args, max_force)
    results = []
    os.makedirs(os.path.join(args.out_dir, "log"), exist_ok=True)
    with open(os.path.join(args.out_dir, "log", f"{args.exp_name}.txt"), "w") as f:
        if args.batched:
            # problem q = path (q // runs), repetition (q % runs).  The reference re-runs optimize() on ONE optimiser object
            # whose workers reseed with their pid, i.e. its repetitions differ only by unreproducible noise; here repetition
            # j of every path gets seed SEED + j.
            Q_paths = [p for p in paths for _ in range(args.runs)]
            seeds = [SEED + j for _ in paths for j in range(args.runs)]
            multi = MultiStochTrajOptimizer(envs_dict[args.env], Q_paths, seeds, verbose=1, **kw)
            uopt, Jopt, logs = multi.optimize()
            for i, path in enumerate(paths):
                sl = slice(i * args.runs, (i + 1) * args.runs)
                best = i * args.runs + int(np.argmax(Jopt[sl]))
                results.append(dict(path=path, mean=float(np.mean(Jopt[sl])), best=float(Jopt[best]), uopt=uopt[best],
                                    Jopt_log=logs[best], r_log=None))
        else:
            for i, path in enumerate(paths):
                best = dict(path=path, mean=0.0, best=-np.inf, uopt=None, Jopt_log=[], r_log=[])
                for j in range(args.runs):
                    opt = StochTrajOptimizer(env=envs_dict[args.env], seed=SEED + j, render=False, initial_guess=None,
                                             verbose=1, path=path, **kw)
                    _uopt, _Jopt, _Jopt_log, _r_log = opt.optimize()
                    best["mean"] += _Jopt / args.runs
                    if _Jopt > best["best"]:
                        best.update(best=float(_Jopt), uopt=_uopt, Jopt_log=_Jopt_log, r_log=_r_log)
                results.append(best)
        for i, r in enumerate(results):
            # The reference saves these three files after EACH path and then replays / records the result with model_test.py
            # (model_optimize.py:97-110), so at any time they hold the path optimised last; under --validate it stops at the first
            # path whose grasps pass validate_path (:112-117).  That screening (Drake IK) is not part of this
            # package, so ALL paths are optimised here: the un-suffixed files end up holding the LAST path -- what the reference
            # leaves behind when no path passes -- and every path additionally keeps its own copy (`..._path<i>`) together with
            # its model_test recording, so that a later `model_test --validate --path_id i` replays controls against the path
            # they were optimised for.
            _save(args.out_dir, args.exp_name, max_force, r["Jopt_log"], r["r_log"], r["uopt"])
            if len(results) > 1 or args.validate:
                _save(args.out_dir, f"{args.exp_name}_path{i}", max_force, r["Jopt_log"], r["r_log"], r["uopt"])
            if not args.no_record:
                from . import model_test as MT
                tips, poses, _, _ = MT.record(args.env, r["uopt"], max_force, r["path"], args.validate)
                for sub in ("tip_data", "object_poses"):
                    os.makedirs(os.path.join(args.out_dir, sub), exist_ok=True)
                for name in ([f"{args.exp_name}_{max_force}"] + ([f"{args.exp_name}_path{i}_{max_force}"] if len(results) > 1 or args.validate else [])):
                    np.save(os.path.join(args.out_dir, "tip_data", f"{name}_tip_poses.npy"), tips)
                    np.save(os.path.join(args.out_dir, "object_poses", f"{name}_object_poses.npy"), poses)
            line = f"Max force: {max_force} mean rewards: {r['mean']}, best rewards:{r['best']}\n"
            print(line, end="")
            f.writelines(line)
    return results


def main(argv=None):
    run(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
