"""CPU restatement of the Monte-Carlo path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Two parts:
  * `risk_statistics` restates the statistics block of
    MonteCarloService.run_monte_carlo_simulation (services/monte_carlo_service.py:305-336)
    with the same NumPy calls; pinned against the reference's own output by
    tests/golden/mc_reference.npz (tests/test_oracle_golden.py).
  * `philox4x32_10`, `gbm_paths`, `bootstrap_paths` restate the GPU generator of
    csrc/montecarlo.cu (Philox4x32-10 as published by Salmon et al., SC'11; Box-Muller
    on 24-bit uniforms; fp64 log-price accumulation).  The reference itself draws from
    NumPy's global MT19937 (:271,:283) -- bit parity with it is impossible by design, so
    agreement with the reference generator is statistical (tests compare moments).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. Counters are uint32 arrays; key words Python ints."""
    c0 = np.asarray(c0, dtype=np.uint64); c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64); c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & MASK
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def u01(x):
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def box_muller(a, b):
    r = np.sqrt(np.abs(np.float32(-2.0) * np.log(u01(a))))   # abs: -0.0 when the uniform rounds to 1
    ang = np.float32(2.0) * u01(b)             # sincospif(2u): sin/cos of pi*2u
    ang64 = ang.astype(np.float64) * np.pi
    return r * np.cos(ang64).astype(np.float32), r * np.sin(ang64).astype(np.float32)


def _increments(mode, n_paths, steps, seed, path_offset, returns=None, block_len=1, log_returns=True):
    gid = np.arange(n_paths, dtype=np.uint64) + np.uint64(path_offset)
    c0 = (gid & MASK).astype(np.uint32)
    c1 = (gid >> np.uint64(32)).astype(np.uint32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    inc = np.zeros((steps, n_paths), dtype=np.float32)
    if mode == 1:
        table = np.asarray(returns, dtype=np.float32)
        if not log_returns:
            table = np.log1p(table).astype(np.float32)
        idx = np.zeros(n_paths, dtype=np.int64)
        left = np.zeros(n_paths, dtype=np.int64)
        R = len(table)
    for t0 in range(0, steps, 4):
        x = philox4x32_10(c0, c1, np.uint32(t0 >> 2), np.uint32(mode), k0, k1)
        if mode == 0:
            z0, z1 = box_muller(x[0], x[1])
            z2, z3 = box_muller(x[2], x[3])
            vals = (z0, z1, z2, z3)
        else:
            vals = []
            for j in range(4):
                fresh = left == 0
                start = ((x[j].astype(np.uint64) * np.uint64(R)) >> np.uint64(32)).astype(np.int64)
                nxt = np.where(idx + 1 == R, 0, idx + 1)
                idx = np.where(fresh, start, nxt)
                left = np.where(fresh, block_len, left) - 1
                vals.append(table[idx])
        for j in range(4):
            if t0 + j < steps:
                inc[t0 + j] = vals[j]
    return inc


def _walk(inc64, s0):
    logS = np.cumsum(inc64, axis=0)
    logS = np.vstack([np.zeros((1, inc64.shape[1])), logS])
    runmax = np.maximum.accumulate(logS, axis=0)
    worst = np.min(logS - runmax, axis=0)
    finals = (s0 * np.exp(logS[-1])).astype(np.float32)
    maxdd = (1.0 - np.exp(worst)).astype(np.float32)
    return finals, maxdd, logS


def gbm_paths(s0, mu, sigma, dt, n_paths, steps, seed, path_offset=0):
    z = _increments(0, n_paths, steps, seed, path_offset)
    drift = (mu - 0.5 * sigma * sigma) * dt
    vol = sigma * np.sqrt(dt)
    return _walk(drift + vol * z.astype(np.float64), s0)


def bootstrap_paths(returns, log_returns, s0, n_paths, steps, seed, path_offset=0, block_len=1):
    inc = _increments(1, n_paths, steps, seed, path_offset, returns, block_len, log_returns)
    return _walk(inc.astype(np.float64), s0)


def risk_statistics(final_prices, max_drawdowns, initial_price, confidence):
    """monte_carlo_service.py:305-336, same NumPy calls, on float64 arrays."""
    final_prices = np.asarray(final_prices, dtype=np.float64)
    percentiles = [1, 5, 10, 25, 50, 75, 90, 95, 99]
    percentile_values = np.percentile(final_prices, percentiles)
    pct_changes = (final_prices / initial_price - 1) * 100
    var_percentile = 100 * (1 - confidence)
    var = np.percentile(pct_changes, var_percentile)
    cvar = np.mean(pct_changes[pct_changes <= var])
    prob_profit = np.mean(final_prices > initial_price)
    md = np.asarray(max_drawdowns, dtype=np.float64)
    return {
        "percentiles": {str(p): {"price": float(v), "pct_change": float((v / initial_price - 1) * 100)}
                        for p, v in zip(percentiles, percentile_values)},
        "expected": {"price": float(np.mean(final_prices)), "pct_change": float(np.mean(pct_changes))},
        "risk_metrics": {"var": float(abs(var)), "cvar": float(abs(cvar)), "prob_profit": float(prob_profit),
                         "prob_loss": float(1 - prob_profit),
                         "max_drawdown": {"mean": float(np.mean(md)), "median": float(np.median(md)),
                                          "max": float(np.max(md))}},
    }


def path_drawdowns(paths):
    """:327-336 for a (days, n) array."""
    paths = np.asarray(paths, dtype=np.float64)
    running_max = np.maximum.accumulate(paths, axis=0)
    return ((running_max - paths) / running_max).max(axis=0)


def drift_and_vol(returns, scenario_params=None):
    """:236-256 (pandas mean / std ddof=1, x252 / sqrt(252), scenario factors)."""
    import pandas as pd
    r = pd.Series(np.asarray(returns, dtype=np.float64)).dropna()
    mu = r.mean() * 252
    sigma = r.std() * np.sqrt(252)
    sp = scenario_params or {}
    return mu * sp.get("drift_factor", 1.0), sigma * sp.get("volatility_factor", 1.0)


def numpy_gbm_reference(s0, mu, sigma, days, n_paths, seed=0):
    """The reference's OWN arithmetic and generator for the GBM mode (monte_carlo_service.py:266-273 paths, :327-336 per-path
    drawdowns, in its vectorised form): a float64 (days, n) array stepped serially in time with NumPy's standard normals.
    Used as the CPU baseline of the Monte-Carlo leg (bench.py) and for distribution-level tests -- its stream is NumPy's,
    not Philox, so it is comparable with the kernels only statistically.  -> (finals, max drawdowns)"""
    rng = np.random.RandomState(seed)
    dt = 1.0 / 252
    paths = np.zeros((days, n_paths))
    paths[0] = s0
    for t in range(1, days):
        z = rng.standard_normal(n_paths)
        paths[t] = paths[t - 1] * np.exp((mu - 0.5 * sigma ** 2) * dt + sigma * np.sqrt(dt) * z)
    return paths[-1], path_drawdowns(paths)
