"""CPU restatement of the portfolio risk numerics.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows services/portfolio_risk_service.py of zd87pl/ai-crypto-trader:
  returns                      :208      df['close'].pct_change()
  calculate_var                :217-246  |np.percentile(returns.dropna(), 100(1-c)) * value|
  calculate_conditional_var    :248-284  |mean(returns[returns <= percentile]) * value|
  calculate_asset_correlation  :286-326  DataFrame(returns).corr()
  calculate_portfolio_var      :328-396  sqrt(w' (v v' o C) w) * total_value, identity C if not positive definite
Pinned by tests/golden/pf_reference.* (outputs of the reference's own methods, tests/golden/make_golden.py pf).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def pct_change(close: np.ndarray) -> np.ndarray:
    c = np.asarray(close, dtype=np.float64)
    out = np.full(c.shape, np.nan)
    out[..., 1:] = c[..., 1:] / c[..., :-1] - 1.0
    return out


def value_at_risk(returns: np.ndarray, confidence_level: float = 0.95, value: float = 1.0) -> float:
    r = np.asarray(returns, dtype=np.float64)
    r = r[~np.isnan(r)]
    if r.size < 2:
        return 0.0
    return float(abs(np.percentile(r, 100 * (1 - confidence_level)) * value))


def conditional_value_at_risk(returns: np.ndarray, confidence_level: float = 0.95, value: float = 1.0) -> float:
    r = np.asarray(returns, dtype=np.float64)
    r = r[~np.isnan(r)]
    if r.size < 2:
        return 0.0
    q = np.percentile(r, 100 * (1 - confidence_level))
    return float(abs(r[r <= q].mean() * value))


def correlation(returns: np.ndarray) -> np.ndarray:
    """Pairwise-complete Pearson correlation of the rows of `returns` ([S][N], NaN = missing)."""
    return pd.DataFrame(np.asarray(returns, dtype=np.float64).T).corr().to_numpy()


def portfolio_var(weights_value: np.ndarray, var_estimates: np.ndarray, corr: np.ndarray, total_value: float) -> float:
    v = np.asarray(weights_value, dtype=np.float64)
    total = v.sum()
    if total == 0:
        return 0.0
    w = v / total
    c = np.asarray(corr, dtype=np.float64)
    if not np.all(np.linalg.eigvals(c) > 0):
        c = np.eye(len(w))
    vm = np.outer(var_estimates, var_estimates) * c
    return float(np.sqrt(w @ vm @ w) * total_value)
