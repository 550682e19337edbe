"""Pure-Python float64 restatement of the reference's single-symbol backtest loop
(BASELINE configs[0]).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  TradingSignal._calculate_signal / _calculate_strength   binance_ml_strategy.py:489-581
  PositionSizer.calculate_position_size                   binance_ml_strategy.py:251-291
  StrategyTester.prepare_market_data                      backtesting/strategy_tester.py:63-125
  StrategyTester.backtest_strategy / open / close / stats backtesting/strategy_tester.py:156-430
  StrategyTester.should_execute_trade, AITrader.should_take_trade   :371-401, services/ai_trader.py:368-387
with the two OpenAI calls replaced by the deterministic stub documented in DESIGN.md
(decision = technical signal, confidence 1.0; no AI risk opinion).  Pinned against the
reference's own StrategyTester (run with that stub) by tests/golden/bt_reference.*.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import pandas as pd

from . import indicators_ref as R


def trading_signal(u: Dict):
    """(signal, strength) of TradingSignal for one market update dict."""
    pts = 0.0
    rsi, stoch, macd = u["rsi"], u["stoch_k"], u["macd"]
    wr, trend, ts, bb = u["williams_r"], u["trend"], u["trend_strength"], u["bb_position"]
    if rsi < 35: pts += 3.0
    elif rsi < 45: pts += 2.0
    if stoch < 20: pts += 3.0
    elif stoch < 30: pts += 2.0
    if macd > 0 and macd > macd * 1.1: pts += 3.0        # dead branch for macd > 0 (:509), kept
    elif macd > 0: pts += 2.0
    if wr and wr < -80: pts += 3.0                       # truthiness tests kept (:516,:530,:575)
    elif wr and wr < -65: pts += 2.0
    if trend == "uptrend" and ts and ts > 10: pts += 3.0
    elif trend == "uptrend" and ts and ts > 5: pts += 2.0
    if bb and bb < 0.2: pts += 3.0
    elif bb and bb < 0.4: pts += 2.0
    ratio = pts / 6
    signal = "BUY" if ratio >= 0.6 else ("SELL" if ratio <= 0.3 else "NEUTRAL")
    if signal == "NEUTRAL":
        return signal, 0
    s = 0
    s += ((45 - min(rsi, 45)) / 15 if signal == "BUY" else (max(rsi, 55) - 55) / 15) * 30
    s += ((30 - min(stoch, 30)) / 30 if signal == "BUY" else (max(stoch, 70) - 70) / 30) * 20
    s += min(abs(macd), 1) * 20
    s += min(u["avg_volume"] / 100000, 1) * 15
    if ts:
        t = min(ts / 20, 1)
        if (signal == "BUY" and trend == "uptrend") or (signal == "SELL" and trend == "downtrend"):
            s += t * 15
    return signal, min(max(s, 0), 100)


def position_size(total_capital, volatility, volume, max_risk_per_trade=0.15) -> Dict:
    if volatility > 0.02: pct, sl = 0.25, 0.02
    elif volatility > 0.01: pct, sl = 0.20, 0.015
    else: pct, sl = 0.15, 0.01
    size = total_capital * pct * min(volume / 50000, 1)
    size = min(size, (total_capital * max_risk_per_trade) / sl)
    size = min(size, total_capital * 0.20)
    size = max(size, total_capital * 0.10)
    size = max(size, 40)
    return {"position_size": size, "stop_loss_pct": sl, "take_profit_pct": sl * 2.0}


def market_constants(df: pd.DataFrame) -> Dict:
    """The whole-frame constants prepare_market_data puts into every bar's dict (:68-97)."""
    f64 = lambda c: df[c].astype(np.float64)
    sc = R.analyzer_scalars(f64("open"), f64("high"), f64("low"), f64("close"), f64("volume"))
    sc["avg_volume"] = float(df["volume"].mean() * df["close"].mean())          # :74 (pandas means of the frame's dtype)
    return sc


def backtest(df: pd.DataFrame, initial_balance: float = 10000.0, ai_confidence_threshold: float = 0.7) -> Dict:
    u = market_constants(df)
    signal, strength = trading_signal(u)
    # should_take_trade + should_execute_trade with the stub (confidence 1.0, decision = signal)
    can_enter = (1.0 >= ai_confidence_threshold) and strength >= 70 and signal == "BUY"
    close = [float(x) for x in df["close"]]
    times = [t.isoformat() for t in df.index]
    balance, max_equity = initial_balance, initial_balance
    st = {"initial_balance": initial_balance, "total_trades": 0, "winning_trades": 0, "losing_trades": 0,
          "total_profit": 0.0, "total_loss": 0.0, "max_drawdown": 0.0, "max_drawdown_pct": 0.0, "win_rate": 0.0,
          "profit_factor": 0.0, "sharpe_ratio": 0.0, "trades": [],
          "equity_curve": [{"timestamp": times[0], "equity": initial_balance}], "drawdown_curve": []}
    pos = None

    def close_position(price, ts, reason):
        nonlocal balance, pos
        pnl = (price - pos["entry_price"]) * pos["quantity"]
        pnl_pct = ((price - pos["entry_price"]) / pos["entry_price"]) * 100
        balance += pnl
        pos["trade"].update(exit_price=price, exit_time=ts, pnl=pnl, pnl_pct=pnl_pct, exit_reason=reason)
        st["total_trades"] += 1
        if pnl > 0:
            st["winning_trades"] += 1; st["total_profit"] += pnl
        else:
            st["losing_trades"] += 1; st["total_loss"] -= pnl
        pos = None

    for i, price in enumerate(close):
        if i < 10:
            continue
        if pos is not None:
            pnl_pct = ((price - pos["entry_price"]) / pos["entry_price"]) * 100
            if pnl_pct <= -pos["stop_loss_pct"]:
                close_position(price, times[i], "Stop Loss")
            elif pnl_pct >= pos["take_profit_pct"]:
                close_position(price, times[i], "Take Profit")
        if pos is not None:
            continue
        if can_enter:
            pp = position_size(balance, u["volatility"], u["avg_volume"])
            trade = {"entry_price": price, "entry_time": times[i], "quantity": pp["position_size"] / price,
                     "position_size": pp["position_size"], "stop_loss_pct": pp["stop_loss_pct"],
                     "take_profit_pct": pp["take_profit_pct"], "exit_price": None, "exit_time": None, "pnl": None,
                     "pnl_pct": None, "exit_reason": None}
            st["trades"].append(trade)
            pos = {"entry_price": price, "quantity": pp["position_size"] / price, "stop_loss_pct": pp["stop_loss_pct"],
                   "take_profit_pct": pp["take_profit_pct"], "trade": trade}
        st["equity_curve"].append({"timestamp": times[i], "equity": balance})
        if balance > max_equity:
            max_equity = balance
        dd = max_equity - balance
        dd_pct = (dd / max_equity) * 100
        st["drawdown_curve"].append({"timestamp": times[i], "drawdown": dd, "drawdown_pct": dd_pct})
        if dd > st["max_drawdown"]:
            st["max_drawdown"], st["max_drawdown_pct"] = dd, dd_pct
    if pos is not None:
        close_position(close[-1], times[-1], "End of Test")
    st["final_balance"] = balance
    if st["total_trades"] > 0:
        st["win_rate"] = (st["winning_trades"] / st["total_trades"]) * 100
    if st["total_loss"] > 0:
        st["profit_factor"] = st["total_profit"] / st["total_loss"]
    rets, prev = [], initial_balance
    for p in st["equity_curve"]:
        rets.append((p["equity"] - prev) / prev if prev > 0 else 0)
        prev = p["equity"]
    if len(rets) > 1:
        sd = np.std(rets)
        if sd > 0:
            st["sharpe_ratio"] = (np.mean(rets) / sd) * np.sqrt(252)
    st["constants"] = dict(u, signal=signal, strength=strength, can_enter=can_enter)
    return st
