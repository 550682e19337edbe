"""Seeded synthetic 1-minute OHLCV.

The reference ships no market data and its own synthetic generator
(services/strategy_evaluation.py:1197-1296) is daily and unseeded, so the
benchmark and parity inputs are generated here, deterministically:

    per symbol s (seed 1234+s), float64 then cast to fp32:
      close_t = 100*(1+s/10) * exp(cumsum(N(0, 0.001)))
      open_t  = close_{t-1}                      (open_0 = close_0)
      high_t  = max(o,c) * (1 + |N(0, 5e-4)|)
      low_t   = min(o,c) * (1 - |N(0, 5e-4)|)
      volume  = lognormal(mu=5, sigma=1)
    timestamps: 1-minute grid from 2024-01-01T00:00Z.
"""
from __future__ import annotations

import datetime as _dt

import numpy as np

FIELDS = ("open", "high", "low", "close", "volume")
EPOCH_2024_MINUTES = int(_dt.datetime(2024, 1, 1, tzinfo=_dt.timezone.utc).timestamp() // 60)


def synth_symbol(s: int, n_bars: int, seed_base: int = 1234) -> dict:
    """OHLCV for synthetic symbol number `s`: dict field -> float32[n_bars]."""
    rng = np.random.default_rng(seed_base + s)
    close = 100.0 * (1.0 + s / 10.0) * np.exp(np.cumsum(rng.normal(0.0, 0.001, n_bars)))
    open_ = np.empty_like(close)
    open_[0] = close[0]
    open_[1:] = close[:-1]
    hi = np.maximum(open_, close) * (1.0 + np.abs(rng.normal(0.0, 5e-4, n_bars)))
    lo = np.minimum(open_, close) * (1.0 - np.abs(rng.normal(0.0, 5e-4, n_bars)))
    vol = rng.lognormal(5.0, 1.0, n_bars)
    return {
        "open": open_.astype(np.float32),
        "high": hi.astype(np.float32),
        "low": lo.astype(np.float32),
        "close": close.astype(np.float32),
        "volume": vol.astype(np.float32),
    }


def synth_ohlcv(n_symbols: int, n_bars: int, seed_base: int = 1234, first_symbol: int = 0) -> np.ndarray:
    """float32 [5][S][N] in FIELDS order (field-major SoA: one [S][N] matrix per field)."""
    out = np.empty((len(FIELDS), n_symbols, n_bars), dtype=np.float32)
    for i in range(n_symbols):
        d = synth_symbol(first_symbol + i, n_bars, seed_base)
        for f, name in enumerate(FIELDS):
            out[f, i] = d[name]
    return out


def bar_timestamp(t: int, minute0: int = EPOCH_2024_MINUTES, bar_minutes: int = 1) -> str:
    """ISO timestamp of bar t, the form the reference's trade records carry."""
    return (_dt.datetime(1970, 1, 1) + _dt.timedelta(minutes=minute0 + t * bar_minutes)).isoformat()


def param_ranges(leverage_trading: bool = False) -> dict:
    """The 18-gene search space of services/strategy_evolution_service.py:98-117."""
    return {
        "rsi_period": (5, 30),
        "rsi_overbought": (65, 85),
        "rsi_oversold": (15, 35),
        "macd_fast": (8, 20),
        "macd_slow": (20, 40),
        "macd_signal": (5, 15),
        "bollinger_period": (10, 30),
        "bollinger_std": (1.5, 3.0),
        "atr_period": (7, 25),
        "atr_multiplier": (1.0, 4.0),
        "ema_short": (5, 20),
        "ema_long": (20, 100),
        "volume_ma_period": (5, 30),
        "social_sentiment_threshold": (50, 80),
        "social_volume_threshold": (5000, 50000),
        "social_engagement_threshold": (1000, 20000),
        "stop_loss": (1, 5) if not leverage_trading else (0.5, 2.5),
        "take_profit": (1, 10) if not leverage_trading else (2, 20),
    }


def random_population(pop: int, seed: int = 42, leverage_trading: bool = False) -> list:
    """Population drawn the way GeneticAlgorithm.initialize_population does
    (services/genetic_algorithm.py:106-115): stdlib `random`, randint for
    int ranges, uniform for float ranges, parameters in dict order."""
    import random

    rnd = random.Random(seed)
    ranges = param_ranges(leverage_trading)
    out = []
    for _ in range(pop):
        ind = {}
        for name, (lo, hi) in ranges.items():
            if isinstance(lo, int) and isinstance(hi, int):
                ind[name] = rnd.randint(lo, hi)
            else:
                ind[name] = rnd.uniform(lo, hi)
        out.append(ind)
    return out
