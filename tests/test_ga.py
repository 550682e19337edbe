"""GeneticAlgorithm drop-in vs the trajectory produced by the reference's own class."""
import json
import math

import pytest

from conftest import GOLDEN
from ai_crypto_trader_b200 import synth
from ai_crypto_trader_b200.genetic_algorithm import GeneticAlgorithm


def toy_fitness(ind):
    ranges = synth.param_ranges()
    s = 0.0
    for i, (k, (lo, hi)) in enumerate(ranges.items()):
        x = (ind[k] - lo) / (hi - lo)
        s += math.sin(3.0 * x + 0.37 * i) * (1.0 + 0.1 * i)
    return s


@pytest.fixture(scope="module")
def ga_golden():
    return json.loads((GOLDEN / "ga_run.json").read_text())["runs"]


@pytest.mark.parametrize("batched", [False, True])
def test_seeded_run_reproduces_reference_trajectory(ga_golden, batched):
    for ref in ga_golden:
        calls = []

        def batch(pop):
            calls.append(len(pop))
            return [toy_fitness(p) for p in pop]

        ga = GeneticAlgorithm(synth.param_ranges(), toy_fitness, population_size=ref["pop"], generations=ref["generations"],
                              mutation_rate=0.2, crossover_rate=0.8, elitism_pct=0.1, random_seed=ref["seed"],
                              batch_fitness_function=batch if batched else None)
        best = ga.run(seeded_individuals=ref["seeded"])
        assert best == ref["best"]
        assert ga.best_fitness == ref["best_fitness"]
        assert ga.population == ref["final_population"]
        assert ga.fitness_scores == ref["final_fitness"]
        hist = ga.get_generation_history()
        assert len(hist) == len(ref["history"]) == ref["generations"] + 1
        for h, r in zip(hist, ref["history"]):
            for k in ("generation", "best_fitness", "avg_fitness", "min_fitness", "best_individual", "diversity"):
                assert h[k] == r[k], (k, h["generation"])
        assert ga.get_population_diversity() == ref["diversity"]
        assert ga.get_best_individual() == ref["best"]
        if batched:
            assert calls == [ref["pop"]] * (ref["generations"] + 1)   # one call per generation


def test_batch_attribute_on_fitness_function_is_used():
    seen = []

    def f(ind):
        raise AssertionError("serial path must not be used when .batch exists")

    f.batch = lambda pop: (seen.append(len(pop)) or [float(p["rsi_period"]) for p in pop])
    ga = GeneticAlgorithm(synth.param_ranges(), f, population_size=12, generations=2, random_seed=1)
    ga.run()
    assert seen == [12, 12, 12]
    assert ga.best_fitness == max(ga.best_fitness, *ga.fitness_scores)


def test_param_ranges_match_reference_search_space():
    r = synth.param_ranges()
    assert list(r)[:3] == ["rsi_period", "rsi_overbought", "rsi_oversold"] and len(r) == 18
    assert r["stop_loss"] == (1, 5) and synth.param_ranges(True)["take_profit"] == (2, 20)
