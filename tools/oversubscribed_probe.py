import os, sys, time
sys.path.insert(0,'/root/repo'); 
import numpy as np
import redmax_b200 as rb
sg = rb.chain_scene(32, ground=True, h=5e-4, nsteps=24); sg.init()
B=3001
q0,qd0 = rb.synthetic_inputs(sg,B,seed=20260007)
kw=dict(scheme=2, iterMaxFactor=2, nsteps=24)
os.environ['RMX_GROUP']='1'
ref = sg.rollout(q0,qd0,**kw)
for G in ('1','2','5'):
    for sc in ('2','3'):
        os.environ['RMX_GROUP']=G; os.environ['RMX_DEBUG_SLOTS_SCALE']=sc
        t=time.time(); out = sg.rollout(q0,qd0,**kw); dt=time.time()-t
        same = all(np.array_equal(out[k],ref[k]) for k in ('q','qdot','status','iters'))
        print('G=%s slots x%s: %.2f s, bitwise equal to the plain run: %s, status bits seen: %s' % (G, sc, dt, same, sorted(set(out['status'].tolist()))), flush=True)
        os.environ.pop('RMX_DEBUG_SLOTS_SCALE')
