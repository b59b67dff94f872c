"""CPU restatement of the reference's MPPI rollout path -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package may import, link or execute anything under oracle/.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs use it, as the checker or the CPU baseline, never as the thing shipped.

Parity status of each piece (see DESIGN.md "Oracle"):
  decode.py    pinned: reproduces the reference's recorded tip positions (data/tip_data) bit for bit (box envs; plate
               <= 1 float32 ulp) through the single-precision pybullet.multiplyTransforms restated in envsim.py
  scorenet.py  pinned: equals the reference's neurals/network.py LargeScoreFunction run on CPU
  cost.py      pinned against brute-force definitions (Open3D itself is not installable here)
  mppi.py      pinned: same numbers as stoch_traj_opt.py:195-204 run in this container
  dynfeas.py   simple test pinned against utils/dyn_feasibility_check.py imported with stub modules
               and the recorded positives data/dyn_feasible/*.npy;  QP: PARITY UNPINNED (no cvxpy,
               no recorded objectives) -- cross-checked against an independent scipy solve only
  physics/     Bullet restatement (C++ fp64): pinned ONLY by the recorded (u -> pose[100,7])
               trajectories under data/object_poses; achieved tolerances are listed in DESIGN.md
"""
