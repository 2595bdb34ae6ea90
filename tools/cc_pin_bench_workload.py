import sys
sys.path.insert(0,'/root/repo/oracle/ref'); sys.path.insert(0,'/root/repo')
import cc_hybrid as H
import numpy as np, bench, oracle_py
from ntrtsim_b200 import _capi as capi
import os; os.environ["TGB_CC_BENCH_SPEED"] = "4"
model, rods, cab, ctrl, params, _ = bench.build_workload("contact_cable", 12, bench.SEED)
for e in range(12):
    hyb=H.HybridWorld(model, rods[e]); ora=oracle_py.OracleWorld(model,capi.WorldCfg); ora.set_rod_state(rods[e])
    worst=0; same=0; checks=0; nmax=0; nref=0; first=None
    for k in range(30):
        try:
            hyb.step(100)
        except RuntimeError as ex:
            print("env",e,"reference threw at chunk",k,ex); break
        ora.step(H.DT,100)
        err=np.abs(hyb.rod_state()-ora.get_rod_state()).max(); worst=max(worst,err)
        ah=hyb.anchors(); no=int(ora.get_contact_cable_obs()[0,4]); ao=ora.get_contact_cable_anchors(0)
        ok = ah.shape[0]==no and (no==0 or (np.array_equal(ah[:,0],ao[:no,0]) and np.abs(ah[:,1:]-ao[:no,1:]).max()<1e-5))
        same+=int(ok); checks+=1; nmax=max(nmax,no); nref=max(nref,ah.shape[0])
        if not ok and first is None: first=(k,ah.shape[0],no)
    print("env %2d worst %.2e equal %d/%d most anchors oracle %d reference %d first mismatch %s"%(e,worst,same,checks,nmax,nref,first))
    hyb.close()
