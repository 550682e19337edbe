"""Properties of the reference's trade rule (through its pinned restatement oracle/simulate_ref.py) that the time-chunked
sweep relies on.  CPU only.

The speculative scan starts a chunk `warm` bars early from the FLAT state, and the repair pass re-scans a chunk only until
its state equals the recorded trajectory's (csrc/sweep_chunked.cu, chunk_scan_item<REPAIR>).  Both rest on one fact: the
machine's future depends on its past only through (side, entry bar), so two runs that are flat at the same bar produce the
same records from there on."""
from datetime import datetime

import numpy as np
import pytest

from ai_crypto_trader_b200 import synth
from oracle import indicators_ref, simulate_ref

EPOCH = datetime(1970, 1, 1)


def _records(params, price, rsi, start):
    """(bar, side) of every trade record of a run that starts flat at bar `start` (the run's forced close included)."""
    pts = simulate_ref.market_points(price[start:], rsi[start:], "SYM", minute0=start)       # timestamp = bar index in minutes
    recs = simulate_ref.simulate_trades(params, pts)
    return [(int((datetime.fromisoformat(r["timestamp"]) - EPOCH).total_seconds() // 60), r["side"]) for r in recs]


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_two_runs_that_are_flat_at_the_same_bar_coincide_afterwards(seed):
    n = 6000
    close = synth.synth_symbol(seed, n)["close"].astype(np.float32)
    checked = 0
    for params in synth.random_population(6, seed=seed):
        rsi = indicators_ref.rsi_bank(close, [int(params["rsi_period"])])[0]
        full = _records(params, close, rsi, 0)
        held = [(full[i][0], full[i + 1][0]) for i in range(0, len(full) - 1, 2)]       # (entry bar, exit bar) of the full run

        def full_is_flat_after(t):             # the full run holds no position over bar t + 1
            return not any(a <= t < b for a, b in held)

        for start in (700, 1500, 2900):
            late = _records(params, close, rsi, start)
            # the first bar from which BOTH runs are flat: `start` itself, or the bar after an exit of the late run
            sync = start if full_is_flat_after(start - 1) else next(
                (late[i][0] + 1 for i in range(1, len(late), 2) if late[i][0] < n - 1 and full_is_flat_after(late[i][0])), None)
            if sync is None:
                continue                       # they never met inside this series (one long position): nothing to check
            assert [r for r in full if r[0] >= sync] == [r for r in late if r[0] >= sync], (seed, params, start, sync)
            checked += 1
    assert checked >= 6                        # (the property was exercised, not skipped)
