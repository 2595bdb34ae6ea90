import sys, os, tempfile
sys.path.insert(0,'/root/repo/oracle/ref'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import cc_hybrid as H
import numpy as np, oracle_py
import ntrtsim_b200 as nt
from ntrtsim_b200 import _capi as capi
import test_contact_cables as t
from helpers import random_poses
text = t.PRISM_WITH_CONTACT_CABLES.replace("class: tgBasicActuatorInfo", "class: tgBasicContactCableInfo")
d = tempfile.mkdtemp(); path = os.path.join(d, "p.yaml"); open(path, "w").write(text)
m = nt.examples.build_yaml(path)
print("bodies", m.n_bodies, "plain", m.n_cables, "contact", m.n_contact_cables)
states = random_poses(m, 4, seed=21, translate=1.0, lift=(3.0, 6.0))
for e in range(4):
    hyb = H.HybridWorld(m, states[e]); ora = oracle_py.OracleWorld(m, capi.WorldCfg); ora.set_rod_state(states[e])
    worst = 0; nmax = 0; same = 0; checks = 0
    for k in range(12):
        hyb.step(50); ora.step(H.DT, 50)
        err = np.abs(hyb.rod_state() - ora.get_rod_state()).max(); worst = max(worst, err)
        for c in range(m.n_contact_cables):
            ah = hyb.anchors(c); no = int(ora.get_contact_cable_obs()[c, 4]); checks += 1; same += int(ah.shape[0] == no); nmax = max(nmax, no)
    print("env %d: 600 steps, worst |hybrid - oracle| %.2e, anchor counts equal %d/%d, most anchors %d, tension[0] ref %.3f oracle %.3f" % (e, worst, same, checks, nmax, hyb.cable(0)[2], ora.get_contact_cable_obs()[0, 2]))
    hyb.close()
