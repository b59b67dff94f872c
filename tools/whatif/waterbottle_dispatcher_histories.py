"""Which dispatcher arrays are consistent with the manifold orders the waterbottle_20.0 recording needs?  (CPU, seconds.)
Phases found with the oracle's explicit-order knobs: steps 5-19 fingertip < floor < wall2; 20-32 fingertip < wall2 < floor; 33-38
wall2 < fingertip < floor; 39-42 floor < wall3 < fingertip.  Events: outer idle pair O leaves at step 20, the empty wall2-fingertip
manifold at 33, both inner idle pairs X, Y at 39.  Of the 9! arrays 194 survive; with the idle trio in the first three slots as the
other recordings have it, exactly one: [X, O, Y, E, wall2, fingertip, wall3, floor, L] (E: an empty manifold released at 33, L: an
empty one created after the floor manifold), Y released before X.  The oracle builds [X, O, Y, E, wall2, fingertip, L, wall3, floor]."""
# enumerate dispatcher-array histories of the waterbottle recording consistent with the order of the contact-carrying manifolds
import itertools
from bt_quicksort import sort, remove
key=lambda m: 1 if m in ('O','X','Y') else 0            # idle island id larger than the object's
names=['O','X','Y','26','21','61','41','31','01']
def rel(s, *ms):        # relative order of ms in sorted list s
    return [m for m in s if m in ms]
sols=[]
for perm in itertools.permutations(names):
    # phase A (steps 5-19): 61 < 01 < 21
    sA=sort(perm,key)
    if rel(sA,'61','01','21')!=['61','01','21']: continue
    a=remove(perm,'O')                                   # step 20: outer idle pair leaves
    sB=sort(a,key)
    if rel(sB,'61','21','01')!=['61','21','01']: continue
    b=remove(a,'26')                                     # step 33
    sC=sort(b,key)
    if rel(sC,'21','61','01')!=['21','61','01']: continue
    for first,second in (('X','Y'),('Y','X')):
        c=remove(remove(b,first),second)
        sD=sort(c,key)
        if rel(sD,'01','31','61')==['01','31','61']:
            sols.append((perm,first,tuple(sD)))
print(len(sols))
for s in sols[:60]: print(s)
