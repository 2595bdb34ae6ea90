"""Multi-rank parity check of the sharded population evaluation (BASELINE config 4): run under torchrun on N GPUs,
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/population_multirank_check.py
every rank evaluates its shard of P TetraSpine members on its own GPU (f64), the scores are all-gathered over NCCL, and rank 0
compares the gathered scores with the CPU oracle's (test infrastructure) for all P members and with a single-GPU evaluation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import ntrtsim_b200 as nt
from ntrtsim_b200 import _capi as capi
from ntrtsim_b200.neuroevo import evaluate_population

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P, steps = 61, 60   # a ragged split over the ranks
g = np.load(os.path.join(ROOT, "tests", "golden", "tetraspine_cpg.npz"))
rng = np.random.default_rng(3456)
node = np.clip(g["node_actions"][None] + rng.normal(0, 0.03, size=(P, 2)), 0, 1)
edge = np.clip(g["edge_actions"][None] + rng.normal(0, 0.03, size=(P, 57, 2)), 0, 1)
rows = nt.tetraspine_param_rows(node, edge)
model, ctl = nt.build_example("TetraSpine"), nt.tetraspine_cpg()
scores, status = evaluate_population(model, ctl, rows, steps, precision="f64", device=local, rank=rank, world_size=world)
if rank == 0:
    import oracle_py
    rods = np.tile(model.initial_rod_state(), (P, 1, 1))
    cab = np.tile(model.initial_cable_state(), (P, 1, 1))
    _, _, so, _ = oracle_py.rollout_batch(model, capi.WorldCfg, rods, cab, model.dt, steps, capi.TGB_CTRL_CPG_IMPEDANCE, rows, cpg_desc=ctl.desc,
                                          want_scores=True)
    single, _ = evaluate_population(model, ctl, rows, steps, precision="f64", device=local)
    err = float(np.abs(scores - so).max())
    print("population evaluation on %d ranks, %d members, %d steps, f64: max |score - oracle| = %.3e; identical to the single-GPU evaluation: %s"
          % (world, P, steps, err, bool(np.array_equal(scores, single))))
    assert scores.shape == (P, 2) and err < 1e-6 and np.array_equal(scores, single)
    print("OK")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
