"""float64 pandas restatement of the `ta` indicators the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the third-party
package `ta` (bukosabino/ta, unpinned `ta>=0.7.0` in requirements.txt:5) is not
vendored in the reference tree and not installed here, so these functions
restate its published definitions (fillna=False) on top of the same pandas
primitives `ta` itself uses (ewm / rolling).  Call sites restated:
binance_ml_strategy.py:63-182 (TechnicalAnalyzer) and
services/market_monitor_service.py:228-260.

All functions take float64 pandas Series (the fp32 market data widened) and
return float64 Series with NaN where `ta` leaves values undefined;
`handle_nan` applies TechnicalAnalyzer._handle_nan_values
(binance_ml_strategy.py:28-38: ffill, bfill, then 0).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def handle_nan(s: pd.Series) -> pd.Series:
    # fillna(method=...) was removed in pandas 3; same semantics
    return s.ffill().bfill().fillna(0)


def sma(close: pd.Series, w: int) -> pd.Series:
    return close.rolling(window=w, min_periods=w).mean()


def ema(series: pd.Series, w: int) -> pd.Series:
    return series.ewm(span=w, min_periods=w, adjust=False).mean()


def macd(close: pd.Series, fast: int = 12, slow: int = 26, sign: int = 9):
    line = ema(close, fast) - ema(close, slow)
    signal = ema(line, sign)
    return line, signal, line - signal


def rsi(close: pd.Series, w: int = 14) -> pd.Series:
    diff = close.diff(1)
    up = diff.where(diff > 0, 0.0)
    dn = -diff.where(diff < 0, 0.0)
    ema_up = up.ewm(alpha=1 / w, min_periods=w, adjust=False).mean()
    ema_dn = dn.ewm(alpha=1 / w, min_periods=w, adjust=False).mean()
    rs = ema_up / ema_dn
    return pd.Series(np.where(ema_dn == 0, 100, 100 - (100 / (1 + rs))), index=close.index)


def stochastic(high, low, close, w: int = 14, smooth: int = 3):
    smin = low.rolling(w, min_periods=w).min()
    smax = high.rolling(w, min_periods=w).max()
    k = 100 * (close - smin) / (smax - smin)
    d = k.rolling(smooth, min_periods=smooth).mean()
    return k, d


def williams_r(high, low, close, lbp: int = 14) -> pd.Series:
    hh = high.rolling(lbp, min_periods=lbp).max()
    ll = low.rolling(lbp, min_periods=lbp).min()
    return -100 * (hh - close) / (hh - ll)


def bollinger(close: pd.Series, w: int = 20, dev: float = 2):
    mavg = close.rolling(w, min_periods=w).mean()
    mstd = close.rolling(w, min_periods=w).std(ddof=0)
    hband = mavg + dev * mstd
    lband = mavg - dev * mstd
    return hband, mavg, lband


def bollinger_width_position(close, hband, mavg, lband):
    # binance_ml_strategy.py:152-156
    width = (hband - lband) / mavg
    rng = (hband - lband).replace(0, np.nan)
    return width, (close - lband) / rng


def true_range(high, low, close) -> pd.Series:
    prev = close.shift(1)
    return pd.DataFrame({"a": high - low, "b": (high - prev).abs(), "c": (low - prev).abs()}).max(axis=1)


def atr(high, low, close, w: int = 14) -> pd.Series:
    tr = true_range(high, low, close).to_numpy()
    out = np.zeros(len(tr))
    if len(tr) >= w:
        out[w - 1] = tr[0:w].mean()
        for i in range(w, len(tr)):
            out[i] = (out[i - 1] * (w - 1) + tr[i]) / float(w)
    return pd.Series(out, index=close.index)


def ichimoku(high, low, w1: int = 9, w2: int = 26, w3: int = 52):
    conv = 0.5 * (high.rolling(w1, min_periods=w1).max() + low.rolling(w1, min_periods=w1).min())
    base = 0.5 * (high.rolling(w2, min_periods=w2).max() + low.rolling(w2, min_periods=w2).min())
    a = 0.5 * (conv + base)
    b = 0.5 * (high.rolling(w3, min_periods=w3).max() + low.rolling(w3, min_periods=w3).min())
    return a, b


def vwap(high, low, close, volume, w: int = 14) -> pd.Series:
    tp = (high + low + close) / 3.0
    pv = tp * volume
    return pv.rolling(w, min_periods=w).sum() / volume.rolling(w, min_periods=w).sum()


def rsi_bank(close32: np.ndarray, periods, fill: bool = True) -> np.ndarray:
    """[P][N] float32 bank from one fp32 close row: float64 RSI, NaN-filled the
    TechnicalAnalyzer way, rounded once to fp32 -- the contract of b200bt_rsi_bank."""
    c = pd.Series(close32.astype(np.float64))
    rows = []
    for w in periods:
        r = rsi(c, int(w))
        if fill:
            r = handle_nan(r)
        rows.append(r.to_numpy().astype(np.float32))
    return np.stack(rows)


def analyzer_columns(open_, high, low, close, volume) -> dict:
    """All TechnicalAnalyzer columns (binance_ml_strategy.py:40-182) after _handle_nan_values."""
    cols = {}
    cols["sma_20"], cols["sma_50"], cols["sma_200"] = sma(close, 20), sma(close, 50), sma(close, 200)
    cols["ema_12"], cols["ema_26"] = ema(close, 12), ema(close, 26)
    cols["macd"], cols["macd_signal"], cols["macd_diff"] = macd(close)
    cols["ichimoku_a"], cols["ichimoku_b"] = ichimoku(high, low)
    cols["rsi"] = rsi(close)
    cols["stoch_k"], cols["stoch_d"] = stochastic(high, low, close)
    cols["williams_r"] = williams_r(high, low, close)
    cols["bb_high"], cols["bb_mid"], cols["bb_low"] = bollinger(close)
    cols["bb_width"], cols["bb_position"] = bollinger_width_position(close, cols["bb_high"], cols["bb_mid"], cols["bb_low"])
    cols["atr"] = atr(high, low, close)
    cols["vwap"] = vwap(high, low, close, volume)
    return {k: handle_nan(v) for k, v in cols.items()}


def analyzer_scalars(open_, high, low, close, volume) -> dict:
    """get_all_indicators / get_trend / get_volatility (binance_ml_strategy.py:184-249)."""
    c = analyzer_columns(open_, high, low, close, volume)
    last = lambda s: float(s.iloc[-1])
    last_close, sma20, sma50 = last(close), last(c["sma_20"]), last(c["sma_50"])
    strength = ((last_close - sma20) / sma20 * 100 + (last_close - sma50) / sma50 * 100) / 2
    if last_close > sma20 and sma20 > sma50:
        trend = "uptrend"
    elif last_close < sma20 and sma20 < sma50:
        trend = "downtrend"
    else:
        trend = "sideways"
    return {"rsi": last(c["rsi"]), "stoch_k": last(c["stoch_k"]), "stoch_d": last(c["stoch_d"]), "macd": last(c["macd"]),
            "macd_signal": last(c["macd_signal"]), "williams_r": last(c["williams_r"]), "bb_position": last(c["bb_position"]),
            "volatility": last(c["atr"]) / last_close, "trend": trend, "trend_strength": abs(strength)}


def rsi_rows_on_timeframe(close32: np.ndarray, minute0: int, k: int, periods, bar_minutes: int = 1) -> np.ndarray:
    """[P][N] float32: RSI of the clock-aligned k-minute closes (last close of each bucket, like an exchange kline /
    pandas resample), NaN-filled on that clock the TechnicalAnalyzer way, then brought back to the base clock with the value
    of the last COMPLETED k-minute bar (no look-ahead; NaN before the first completed bar).  The multi-timeframe recipe of
    services/market_monitor_service.py:219-301 applied to one derived series -- the contract of PopulationSweep's
    multi-timeframe bank (b200bt_resample -> b200bt_rsi_bank -> b200bt_align)."""
    n = len(close32)
    m = minute0 + np.arange(n, dtype=np.int64) * bar_minutes
    bucket = m // k - minute0 // k
    last = np.zeros(int(bucket[-1]) + 1, dtype=np.int64)
    last[bucket] = np.arange(n)                                   # ascending assignment: the last bar of each bucket wins
    rows = rsi_bank(np.asarray(close32, dtype=np.float32)[last], periods)      # [P][M] float32
    closes_here = ((m + bar_minutes) // k) != (m // k)
    j = np.where(closes_here, bucket, bucket - 1)
    out = rows[:, np.maximum(j, 0)].astype(np.float32)
    out[:, j < 0] = np.nan
    return out
