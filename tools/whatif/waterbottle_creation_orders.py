"""Second enumeration for waterbottle_20.0 (see waterbottle_dispatcher_histories.py): the idle trio fixed as the other recordings
have it ([inner X, outer O, inner Y]), all 720 creation orders of the six manifolds of world step 3, the empty (wall4, fingertip)
manifold released at step 4 / during one of the phases / never, optionally one extra empty manifold.  Without an extra manifold exactly
one history survives: creation order (wall2-tip, wall2-obj, tip-obj, wall3-obj, wall4-tip, wall4-obj) -- wall by wall, oldest first --
with the (wall4, fingertip) manifold released AFTER the floor manifold was created (steps 6-19) and the inner idle pair Y released
before X.  The oracle creates (wall2-tip, wall4-tip, tip-obj, wall4-obj, wall3-obj, wall2-obj) and releases wall4-tip at step 4."""
import itertools
from bt_quicksort import sort, remove
key=lambda m: 1 if m in ('O','X','Y') else 0
def rel(s,*ms): return [m for m in s if m in ms]
new=['26','46','61','41','31','21']
sols=[]
idle=('X','O','Y')
for perm in itertools.permutations(new):
  for t46 in ('s4','A','B','C','never'):
    for extra in (None,'E_A','E_B','E_C'):          # an extra empty manifold created during phase A / B / C
        a=list(idle)+list(perm)
        if t46=='s4': a=remove(a,'46')
        a=a+['01']
        if extra=='E_A': a=a+['E']
        if t46=='A': a=remove(a,'46')
        if rel(sort(a,key),'61','01','21')!=['61','01','21']: continue
        b=remove(a,'O')
        if extra=='E_B': b=b+['E']
        if t46=='B': b=remove(b,'46')
        if rel(sort(b,key),'61','21','01')!=['61','21','01']: continue
        c=remove(b,'26')
        if extra=='E_C': c=c+['E']
        if t46=='C': c=remove(c,'46')
        if rel(sort(c,key),'21','61','01')!=['21','61','01']: continue
        for f,s in (('X','Y'),('Y','X')):
            d=remove(remove(c,f),s)
            if rel(sort(d,key),'01','31','61')==['01','31','61']: sols.append((perm,t46,extra,f))
print(len(sols))
from collections import Counter
print(Counter((s[1],s[2],s[3]) for s in sols))
for s in sols[:30]: print(s)
