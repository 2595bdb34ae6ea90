import sys
sys.path.insert(0,'/root/repo/oracle/ref'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import cc_hybrid as H
import numpy as np, oracle_py, bench
from ntrtsim_b200 import _capi as capi
import os; os.environ["TGB_CC_BENCH_SPEED"] = "4"
model, rods, cab, ctrl, params, _ = bench.build_workload("contact_cable", 4, bench.SEED)
model.contact_cables[0].target_velocity = 2.0
for e in range(4):
    hyb=H.HybridWorld(model, rods[e]); ora=oracle_py.OracleWorld(model,capi.WorldCfg); ora.set_rod_state(rods[e])
    worst=0; same=0; checks=0; nmax=0
    for k in range(2500):
        target=np.array([8.0-0.0008*k])
        hyb.step(1, control=target)
        ora.world_step(H.DT); ora.set_contact_cable_control_input(target, H.DT); ora.models_step(H.DT)
        if k%100==99:
            err=np.abs(hyb.rod_state()-ora.get_rod_state()).max(); worst=max(worst,err)
            ah=hyb.anchors(); no=int(ora.get_contact_cable_obs()[0,4]); checks+=1; same+=int(ah.shape[0]==no); nmax=max(nmax,no)
    print("env %d: worst %.2e anchors equal %d/%d most %d rest ref %.6f oracle %.6f"%(e,worst,same,checks,nmax,hyb.cable()[0],ora.get_contact_cable_obs()[0,0]))
    hyb.close()
