"""Which first fingertip contact would PyBullet have had to generate for the plate recording's row 1?  The oracle's contact of the
first servo substep (one EPA result: fingertip box vs plate hull) is nudged (oracle knobs cc_nudge0..6: point, distance, normal) and
a least-squares fit asks which nudge reproduces the recorded row.  Result ("seed"): a shift of the contact point by
7e-9 and of the distance by 5e-9 reproduces row 1 to 3e-15 -- the level at which the converged EPA result depends on the path
of the expansion (the support points of margin-rounded shapes carry their query direction).  CPU only.
  python tools/oracle_contact_inverse.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import oracle_spec, recorded
from oracle.envsim import EnvSim
from synthesize_pregrasp_b200.envspec import make_spec
exp="plate_20.0"; STEP=3; ROW=1
g=recorded(exp); osd=oracle_spec(make_spec(g["env"]))
def f(x):
    params={f"cc_nudge{i}":float(x[i]) for i in range(7)}; params["dbg_step"]=STEP
    sim=EnvSim(osd,g["max_force"],params=params); sim.reset()
    P=sim.step(g["u"][0],[0,0,0])["poses"]
    return P[ROW]-g["poses"][ROW]
x=np.zeros(7); r0=f(x); print("r0",r0)
idx=[0,1,2,3,4,5,6]
h=np.array([1e-7]*4+[1e-7]*3)
J=np.zeros((7,7))
for i in idx:
    xp=x.copy(); xp[i]+=h[i]; J[:,i]=(f(xp)-r0)/h[i]
np.set_printoptions(precision=3,linewidth=200)
print("J"); print(J)
for cols,name in (([0,1,2],"point"),([3],"depth"),([4,5,6],"normal"),([0,1,2,3],"point+depth"),([0,1,2,3,4,5,6],"all")):
    dx,res,rk,sv=np.linalg.lstsq(J[:,cols],-r0,rcond=1e-10)
    xx=np.zeros(7); xx[cols]=dx
    r1=f(xx)
    print(name,"dx",dx,"| residual before %.2e after %.2e"%(np.abs(r0).max(),np.abs(r1).max()), "rank",rk)
