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
    if world_size == 1:   # the whole population was evaluated here, whatever process group the caller may have initialised
        return np.asarray(local, dtype=np.float64).reshape(-1, 2), status
    return allgather_scores(local, P, group=group), status


# ---- population bookkeeping on the host (src/learning/NeuroEvolution) -----------------------------------------------------------
class NeuroEvoPopulation(object):
    """One pool of members with stateless parameters in [0, 1] (numberOfStates = 0, the configuration of every shipped spine:
    NeuroEvoMember.cpp:40-52), ordered and renewed like src/learning/NeuroEvolution/NeuroEvoPopulation.cpp:
      * update_scores: score = multiscore[0]; past scores are appended, the running maximum decays towards lower scores with
        leniencyCoef (NeuroEvolution.cpp:255-275);
      * order: by average of the past scores or by the (lenient) maximum, best first (NeuroEvoPopulation.cpp:141-166; the reference
        sums the past scores with an int accumulator -- std::accumulate(..., 0) -- which truncates every partial sum: reproduced);
      * mutate: the best numMutate members are copied over the worst numMutate and mutated (NeuroEvoPopulation.cpp:62-77);
      * NeuroEvoMember::mutate (NeuroEvoMember.cpp:57-88): with probability 1/2 nothing; else every parameter with probability
        1/2 gets N(0, 0.03) added and is clipped to [0, 1].
    The reference draws from std::ranlux48_base through the standard library's distributions; here a numpy Generator: the same
    distributions, not the same stream."""

    def __init__(self, population_size, n_params, rng, compare_average_scores=False, clear_scores_between_generations=False,
                 leniency_coef=0.5):
        self.rng = rng
        self.params = rng.uniform(size=(population_size, n_params))   # rand() / RAND_MAX per parameter
        self.past_scores = [[] for _ in range(population_size)]
        self.max_score = np.full(population_size, -1000.0)
        self.max_score1 = np.zeros(population_size)
        self.max_score2 = np.zeros(population_size)
        self.average_score = np.zeros(population_size)
        self.compare_average_scores = bool(compare_average_scores)
        self.clear_scores = bool(clear_scores_between_generations)
        self.leniency = float(leniency_coef)

    def __len__(self):
        return self.params.shape[0]

    def update_scores(self, member, multiscore):
        score = 1.0 * multiscore[0] - 0.0 * multiscore[1]
        self.past_scores[member].append(score)
        prev = self.max_score[member]
        if prev > score:
            self.max_score[member] = self.leniency * prev + (1.0 - self.leniency) * score
        else:
            self.max_score[member] = score
            self.max_score1[member], self.max_score2[member] = multiscore[0], multiscore[1]

    def order(self):
        for i, ps in enumerate(self.past_scores):
            if ps:
                acc = 0
                for s in ps:
                    acc = int(acc + s)   # std::accumulate with an int initial value
                self.average_score[i] = acc / float(len(ps))
            else:
                self.average_score[i] = -100000.0
            if self.clear_scores:
                self.past_scores[i] = []
        key = self.average_score if self.compare_average_scores else self.max_score
        order = np.argsort(-key, kind="stable")
        self.params = self.params[order]
        self.past_scores = [self.past_scores[i] for i in order]
        for name in ("max_score", "max_score1", "max_score2", "average_score"):
            setattr(self, name, getattr(self, name)[order])
        return order

    def mutate_member(self, i):
        if self.rng.uniform() > 0.5:
            return
        for k in range(self.params.shape[1]):
            if self.rng.uniform() > 0.5:
                continue
            self.params[i, k] = min(1.0, max(0.0, self.params[i, k] + self.rng.normal(0.0, 3.0 / 100.0)))

    def mutate(self, num_mutate):
        if num_mutate > len(self) // 2:
            raise ValueError("Trying to mutate more than half of the population")
        for i in range(num_mutate):
            dst = len(self) - 1 - i
            self.params[dst] = self.params[i]     # stateless copyFrom keeps the target's scores (NeuroEvoMember.cpp:90-102)
            self.mutate_member(dst)

    def save_best(self, path):
        """NeuroEvoMember::saveToFile: comma-separated, no trailing comma."""
        with open(path, "w") as f:
            f.write(",".join("%g" % v for v in self.params[0]))


def load_nnw(path, n_params, defaults=None):
    """NeuroEvoMember::loadFromFile (NeuroEvoMember.cpp:150-190): values are read up to each comma, and the value after the last
    comma is dropped (getline hits end of file first) -- so a file of n values sets the first n - 1 parameters."""
    out = np.array(defaults if defaults is not None else np.zeros(n_params), dtype=np.float64).copy()
    text = open(path).read()
    parts = text.split(",")
    for i, v in enumerate(parts[:-1]):
        if i < n_params:
            out[i] = float(v) if v.strip() else 0.0   # atof
    return out


class NeuroEvolution(object):
    """The generation loop of src/learning/NeuroEvolution/NeuroEvolution.cpp:138-249 with the episodes of a generation evaluated in
    one batched call instead of one process per episode: `evaluate(param_rows [P, n_params]) -> scores [P, 2]` is the device side
    (evaluate_population for the spines).  As in the reference with coevolution off, the first generation tests every member and
    later ones only the members renewed by mutation (the last numberOfElementsToMutate)."""

    def __init__(self, n_params, population_size, number_of_elements_to_mutate, seed=0, **population_kw):
        self.rng = np.random.default_rng(seed)
        self.population = NeuroEvoPopulation(population_size, n_params, self.rng, **population_kw)
        self.num_mutate = int(number_of_elements_to_mutate)
        self.generation = 0
        self.first = 0   # first member to test in the coming generation
        self.log = []    # (episodes so far, mean score 1, mean score 2, leader's max score) per generation, the evolution log's columns

    def step(self, evaluate):
        pop = self.population
        tested = list(range(self.first, len(pop)))
        scores = np.asarray(evaluate(pop.params[tested]), dtype=np.float64).reshape(len(tested), 2)
        for m, s in zip(tested, scores):
            pop.update_scores(m, s)
        self.generation += 1
        pop.order()
        self.log.append((self.generation * len(tested), float(scores[:, 0].mean()), float(scores[:, 1].mean()), float(pop.max_score[0])))
        pop.mutate(self.num_mutate)
        self.first = len(pop) - self.num_mutate
        return scores


def read_configuration(path):
    """The key = value files of the learning examples (Config.ini, nodeConfig.ini, edgeConfig.ini), read like
    src/learning/Configuration/configuration.cpp:77-118: blank lines and lines starting with '#' or ';' are skipped, keys and values
    are trimmed, later keys overwrite earlier ones; a missing file yields an empty configuration (the reference only warns)."""
    data = {}
    try:
        lines = open(path).read().split("\n")
    except IOError:
        print("Warning! Config.ini file not found!")
        return data
    for s in lines:
        t = s.lstrip(" \f\t\v")
        if not t or t[0] in "#;":
            continue
        key, _, value = t.partition("=")
        key = key.rstrip(" \f\t\v")
        if not key:
            continue
        data[key] = value.strip(" \f\n\r\t\v")
    return data


def neuroevolution_from_configuration(path, n_params, seed=None):
    """A NeuroEvolution set up from an .ini of the learning examples: populationSize, numberOfElementsToMutate, leniencyCoef,
    compareAverageScores, clearScoresBetweenGenerations, startSeed (NeuroEvolution.cpp:70-110, NeuroEvoPopulation.cpp:40-52); missing
    integer keys read as 0 like configuration::getintvalue."""
    c = read_configuration(path)

    def geti(k):
        try:
            return int(float(c.get(k, 0)))
        except ValueError:
            return 0
    return NeuroEvolution(n_params, geti("populationSize"), geti("numberOfElementsToMutate"),
                          seed=geti("startSeed") if seed is None else seed, compare_average_scores=geti("compareAverageScores") != 0,
                          clear_scores_between_generations=geti("clearScoresBetweenGenerations") != 0,
                          leniency_coef=float(c.get("leniencyCoef", 0.0) or 0.0))
