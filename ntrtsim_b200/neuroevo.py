"""Population evaluation for the learning examples: the evaluate(parameters) -> scores[2] contract of
NeuroAdapter / AnnealAdapter (src/learning/Adapters/NeuroAdapter.cpp:72-133), batched.

Every population member is one environment.  Members shard across the GPUs of a box by contiguous ranges (one process per
GPU, no exchange during an episode); the per-rollout scores are exchanged once per episode with an all-gather
(torch.distributed: NCCL between GPUs, gloo in the CPU tests).  Population bookkeeping (ordering, mutation, .nnw files)
stays on the host exactly as in src/learning/NeuroEvolution/NeuroEvolution.cpp:138-249."""
import numpy as np


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced partition of the env axis: rank r gets [begin, end)."""
    base, extra = divmod(int(n_items), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allgather_scores(local_scores, n_items, group=None, device=None):
    """All-gather of the per-rollout scores [n_local, 2] -> [n_items, 2] (ragged shards are padded to the largest)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(local_scores, dtype=np.float64).reshape(-1, 2)
    world = dist.get_world_size(group)
    longest = max(shard_bounds(n_items, world, r)[1] - shard_bounds(n_items, world, r)[0] for r in range(world))
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    pad = torch.zeros((longest, 2), dtype=torch.float64, device=dev)
    loc = torch.as_tensor(np.asarray(local_scores, dtype=np.float64).reshape(-1, 2), device=dev)
    pad[: loc.shape[0]] = loc
    out = torch.empty((world * longest, 2), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.cpu().numpy().reshape(world, longest, 2)
    parts = []
    for r in range(world):
        b, e = shard_bounds(n_items, world, r)
        parts.append(out[r, : e - b])
    return np.concatenate(parts, axis=0)


def evaluate_population(model, controller, param_rows, n_steps, dt=None, precision="f32", device=0, rank=0, world_size=1, group=None,
                        steps_per_launch=None):
    """Run one episode of `n_steps` for every row of `param_rows` ([P, 2 + 2 * n_pairs], see tetraspine_param_rows) and return
    (scores [P, 2], status [P_local]).  scores[:, 0] = XZ distance moved by the scored segment, or -1 when the episode was
    bogus (height window left); scores[:, 1] = 0 (history logging is off in the shipped configuration)."""
    from .world import BatchedWorld
    param_rows = np.ascontiguousarray(param_rows, dtype=np.float64)
    P = param_rows.shape[0]
    b, e = shard_bounds(P, world_size, rank)
    dt = dt or model.dt
    local = np.zeros((0, 2))
    status = np.zeros(0, dtype=np.int32)
    if e > b:
        w = BatchedWorld(model, e - b, precision=precision, device=device)
        try:
            w.set_cpg_controller(controller, param_rows[b:e])
            chunk = steps_per_launch or n_steps
            done = 0
            while done < n_steps:
                k = min(chunk, n_steps - done)
                w.step(dt, k)
                done += k
            local = w.scores()
            status = w.status()
        finally:
            w.close()
    return allgather_scores(local, P, group=group), status
