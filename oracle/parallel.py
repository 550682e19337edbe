"""The C oracle (oracle/sim_oracle.c) over whole populations on all host cores.  TEST INFRASTRUCTURE.

Jobs are (symbol, rsi period): a worker regenerates the synthetic symbol (ai_crypto_trader_b200.synth is
deterministic), computes that period's RSI row with the float64 pandas restatement (oracle/indicators_ref.py)
and runs every individual of the population that uses the period through `oracle_lanes`
(= _simulate_trades + calculate_metrics + _calculate_strategy_score, strategy_evaluation.py:746-878,:32-228,:579-633).
Workers are spawned (never forked: the parent usually holds a CUDA context).
"""
from __future__ import annotations

import functools
import os
from concurrent.futures import ProcessPoolExecutor
from typing import Dict, List, Optional, Sequence

import numpy as np


@functools.lru_cache(maxsize=2)
def _close(symbol: int, n_bars: int, seed_base: int) -> np.ndarray:
    from ai_crypto_trader_b200 import synth
    return synth.synth_symbol(symbol, n_bars, seed_base)["close"]


def _job(args):
    symbol, n_bars, seed_base, period, rows, plist, minute0, bar_minutes, goals, close_override = args
    from oracle import indicators_ref, sim_oracle
    close = close_override if close_override is not None else _close(symbol, n_bars, seed_base)
    rsi = rows if isinstance(rows, np.ndarray) else indicators_ref.rsi_bank(close, [period])[0]
    cfg = sim_oracle.config_of(minute0, bar_minutes, goals)
    return sim_oracle.lanes(close, rsi, plist, cfg)


def default_workers() -> int:
    return max(1, min(os.cpu_count() or 1, 64))


def population_stats(population: List[Dict], symbols: Sequence[int], n_bars: int, minute0: int, bar_minutes: int = 1,
                     goals: Optional[Dict] = None, seed_base: int = 1234, workers: Optional[int] = None,
                     lanes_of: Optional[Sequence[int]] = None) -> np.ndarray:
    """sim_oracle.STATS_DTYPE array [len(population)][len(symbols)] for the synthetic market
    (symbol numbers as in synth.synth_ohlcv(first_symbol=...)).  `lanes_of`: evaluate only these
    individuals (the other rows stay zero)."""
    from oracle import sim_oracle
    sim_oracle.build()
    pick = list(range(len(population))) if lanes_of is None else [int(i) for i in lanes_of]
    by_period: Dict[int, List[int]] = {}
    for i in pick:
        by_period.setdefault(int(population[i].get("rsi_period", 14)), []).append(i)
    jobs, where = [], []
    for j, s in enumerate(symbols):
        for w, idx in sorted(by_period.items()):
            jobs.append((int(s), int(n_bars), seed_base, w, None, [population[i] for i in idx], int(minute0),
                         int(bar_minutes), goals, None))
            where.append((j, idx))
    out = np.zeros((len(population), len(symbols)), dtype=sim_oracle.STATS_DTYPE)
    import multiprocessing as mp
    with ProcessPoolExecutor(max_workers=workers or default_workers(), mp_context=mp.get_context("spawn")) as ex:
        # consecutive jobs share a symbol: chunks keep a worker on one symbol's cached close row
        for (j, idx), res in zip(where, ex.map(_job, jobs, chunksize=max(1, len(by_period) // 4))):
            out[idx, j] = res
    return out
