"""world_size-2 gloo run of the sharded fitness path (host logic of the N>1 bench)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

WORKER = r'''
import os, sys, math
sys.path.insert(0, os.environ["B200BT_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ai_crypto_trader_b200 import synth
from ai_crypto_trader_b200.dist import ShardedFitness, shard_bounds
from ai_crypto_trader_b200.genetic_algorithm import GeneticAlgorithm
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
calls = []
def local_eval(pop):
    calls.append(len(pop))
    return np.array([p["rsi_period"] * 1.5 + p["take_profit"] - 0.01 * p["ema_long"] for p in pop])
fit = ShardedFitness(local_eval)
for n in (1, 2, 7, 10):
    pop = synth.random_population(n, seed=n)
    got = fit.batch(pop)
    want = local_eval(pop).tolist()
    assert got == want, (n, got, want)
calls.clear()
ga = GeneticAlgorithm(synth.param_ranges(), fit, population_size=11, generations=3, random_seed=5)
best = ga.run()
assert calls == [shard_bounds(11, world, rank)[1] - shard_bounds(11, world, rank)[0]] * 4, calls
# identical trajectory on every rank
blob = torch.tensor([ga.best_fitness, float(sum(ga.fitness_scores))], dtype=torch.float64)
out = [torch.zeros_like(blob) for _ in range(world)]
dist.all_gather(out, blob)
assert all(torch.equal(o, out[0]) for o in out)
# ranks that disagree on the population, or a rank whose local evaluation throws, fail TOGETHER (no silent mixing of
# unrelated fitness values, no rank left waiting in the collective)
from ai_crypto_trader_b200.dist import PopulationMismatch, broadcast_seed
try:
    fit.batch(synth.random_population(6, seed=100 + rank))
    raise SystemExit("different populations were not detected")
except PopulationMismatch:
    pass
def flaky(pop):
    if rank == 1:
        raise ValueError("boom")
    return np.zeros(len(pop))
try:
    ShardedFitness(flaky).batch(synth.random_population(6, seed=1))
    raise SystemExit("a failing rank was not reported")
except RuntimeError as e:
    assert "rank(s) [1]" in str(e), str(e)
assert fit.batch(synth.random_population(5, seed=3)) == local_eval(synth.random_population(5, seed=3)).tolist()   # still usable
seeds = [broadcast_seed(None), broadcast_seed(77 + rank)]
blob = torch.tensor(seeds, dtype=torch.int64)
out = [torch.zeros_like(blob) for _ in range(world)]
dist.all_gather(out, blob)
assert all(torch.equal(o, out[0]) for o in out) and seeds[1] == 77
# Monte-Carlo shards: contiguous path ranges, one gather, every rank ends with the full arrays in path order
from ai_crypto_trader_b200.dist import gather_paths
for n in (1, 2, 9, 64):
    lo, hi, _ = shard_bounds(n, world, rank)
    full = torch.arange(n, dtype=torch.float32) * 0.5 + 100.0
    f, d = gather_paths(full[lo:hi].clone(), -full[lo:hi].clone(), n)
    assert torch.equal(f, full) and torch.equal(d, -full), (n, rank)
if rank == 0:
    print("GLOO_OK", ga.best_fitness)
dist.destroy_process_group()
'''


def test_sharded_fitness_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, B200BT_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "GLOO_OK" in res.stdout
