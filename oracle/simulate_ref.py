"""Pure-Python float64 restatement of the reference's parameterised backtest rule.

Follows services/strategy_evaluation.py:
  simulate_trades   <- StrategyEvaluationSystem._simulate_trades        :746-878
  calculate_metrics <- StrategyPerformanceMetrics.calculate_metrics     :32-228
  strategy_score    <- StrategyEvaluationSystem._calculate_strategy_score :579-633

Written as a restatement (own structure), validated record-for-record against
the reference's own functions by tests/test_oracle_golden.py.
Test infrastructure only -- see oracle/__init__.py.
"""
from __future__ import annotations

import math
from datetime import datetime
from typing import Dict, List, Sequence

import numpy as np

FEE_RATE = 0.001  # :798, :811  (0.1 % per side)


def position_size_of(params: Dict) -> float:
    # :761-764
    initial_capital = 10000
    pct = min(params.get("max_position_size", 5), 20) / 100
    return initial_capital * pct


def simulate_trades(params: Dict, market_data: Sequence[Dict]) -> List[Dict]:
    """Trade records (one per entry, one per exit) for one parameter set.

    market_data: sequence of dicts with keys timestamp, symbol, price, rsi
    (the reference's data-point schema, :777-781).
    """
    size = position_size_of(params)
    oversold = params.get("rsi_oversold", 30)       # :772
    overbought = params.get("rsi_overbought", 70)   # :771
    tp = params.get("take_profit", 3) / 100         # :773
    sl = params.get("stop_loss", 2) / 100           # :774

    records: List[Dict] = []
    side = 0          # 0 flat, +1 long, -1 short
    entry = 0.0

    def emit(point, action, price, qty, pnl):
        records.append({"timestamp": point["timestamp"], "symbol": point["symbol"], "side": action,
                        "price": price, "quantity": qty, "fees": size * FEE_RATE, "pnl": pnl})

    def close_pnl(price):
        qty = size / entry
        move = (price - entry) if side > 0 else (entry - price)
        return qty, qty * move - size * 0.002        # :819, :836

    for i, raw in enumerate(market_data):
        point = {"timestamp": raw.get("timestamp", f"2023-01-{i + 1:02d}T00:00:00Z"),
                 "symbol": raw.get("symbol", "BTCUSDT")}
        price = raw.get("price", 50000)
        rsi = raw.get("rsi", 50)
        if side == 0:
            # entry: long on oversold has priority over short on overbought (:785-813)
            if rsi < oversold:
                side, entry = 1, price
                emit(point, "buy", price, size / price, -size * FEE_RATE)
            elif rsi > overbought:
                side, entry = -1, price
                emit(point, "sell", price, size / price, -size * FEE_RATE)
            continue
        # exit (:815-847): take-profit, stop-loss, or RSI reversal
        gain = ((price - entry) if side > 0 else (entry - price)) / entry
        reversal = (rsi > overbought) if side > 0 else (rsi < oversold)
        if gain >= tp or gain <= -sl or reversal:
            qty, pnl = close_pnl(price)
            emit(point, "sell" if side > 0 else "buy", price, qty, pnl)
            side = 0

    if side != 0 and len(market_data) > 0:
        # forced close on the last bar (:849-876)
        last = market_data[-1]
        point = {"timestamp": last.get("timestamp", "2023-01-31T00:00:00Z"),
                 "symbol": last.get("symbol", "BTCUSDT")}
        price = last.get("price", 50000)
        qty, pnl = close_pnl(price)
        emit(point, "sell" if side > 0 else "buy", price, qty, pnl)
    return records


_EMPTY = dict(total_trades=0, win_rate=0.0, profit_factor=0.0, sharpe_ratio=0.0, max_drawdown=0.0,
              average_profit=0.0, average_loss=0.0, largest_profit=0.0, largest_loss=0.0,
              total_profit=0.0, total_loss=0.0, net_profit=0.0, return_pct=0.0, avg_trade_duration=0,
              risk_reward_ratio=0.0)


def calculate_metrics(trades: List[Dict], initial_capital: float = 10000.0) -> Dict:
    """Metrics over trade RECORDS (:32-228).  Returns the scalar fields plus
    daily_returns / equity_curve; the per-symbol and monthly dicts of the
    reference are omitted (they do not feed the score)."""
    if not trades:
        return dict(_EMPTY, daily_returns={}, equity_curve=[initial_capital])
    recs = sorted(trades, key=lambda r: r.get("timestamp", ""))   # stable, :65
    pnls = [r.get("pnl", 0) for r in recs]
    n = len(recs)
    if n < 2:
        # degenerate single-record case (:68-92)
        p = pnls[0]
        return dict(_EMPTY, total_trades=1, win_rate=1.0 if p > 0 else 0.0,
                    average_profit=p if p > 0 else 0.0, average_loss=p if p < 0 else 0.0,
                    largest_profit=p if p > 0 else 0.0, largest_loss=p if p < 0 else 0.0,
                    total_profit=p if p > 0 else 0.0, total_loss=p if p < 0 else 0.0,
                    net_profit=p, return_pct=(p / initial_capital) * 100,
                    daily_returns={}, equity_curve=[initial_capital])
    wins = [p for p in pnls if p > 0]
    losses = [p for p in pnls if p < 0]
    total_profit = sum(wins)
    total_loss = sum(losses)
    net = total_profit + total_loss
    profit_factor = abs(total_profit / total_loss) if total_loss != 0 else float("inf")   # :113
    avg_p = total_profit / len(wins) if wins else 0
    avg_l = total_loss / len(losses) if losses else 0

    equity = [initial_capital]
    peak = initial_capital
    worst_dd = 0
    seen_dd = False
    daily: Dict[str, float] = {}
    for r, p in zip(recs, pnls):
        cur = equity[-1] + p
        equity.append(cur)
        if cur > peak:                       # :139-144
            peak = cur
        else:
            dd = (peak - cur) / peak
            worst_dd = dd if not seen_dd else max(worst_dd, dd)
            seen_dd = True
        ts = r.get("timestamp", "")
        if ts:
            day = datetime.fromisoformat(ts.replace("Z", "+00:00")).strftime("%Y-%m-%d")
            daily[day] = daily[day] + p if day in daily else p      # :153-157
    max_dd = worst_dd if seen_dd else 0

    durations = []
    for i in range(0, n - 1, 2):             # :167-176 pairs (0,1), (2,3), ...
        t0 = datetime.fromisoformat(recs[i].get("timestamp", "").replace("Z", "+00:00"))
        t1 = datetime.fromisoformat(recs[i + 1].get("timestamp", "").replace("Z", "+00:00"))
        durations.append((t1 - t0).total_seconds() / 60)
    avg_dur = sum(durations) / len(durations) if durations else 0

    vals = list(daily.values())
    if len(vals) > 1:                        # :181-188
        sd = np.std(vals)
        sharpe = (np.mean(vals) / sd) * np.sqrt(252) if sd > 0 else 0
    else:
        sharpe = 0
    return dict(total_trades=n, win_rate=len(wins) / n, profit_factor=profit_factor, sharpe_ratio=sharpe,
                max_drawdown=max_dd, average_profit=avg_p, average_loss=avg_l,
                largest_profit=max(wins) if wins else 0, largest_loss=min(losses) if losses else 0,
                total_profit=total_profit, total_loss=total_loss, net_profit=net,
                return_pct=(net / initial_capital) * 100, avg_trade_duration=avg_dur,
                risk_reward_ratio=abs(avg_p / avg_l) if avg_l != 0 else float("inf"),
                daily_returns=daily, equity_curve=equity)


ADVANCED_KEYS = ("calmar_ratio", "sortino_ratio", "recovery_factor", "expectancy", "profit_per_day")


def calculate_advanced_metrics(metrics: Dict) -> Dict:
    """calculate_advanced_metrics (:231-319) on the dict calculate_metrics returns.  That dict carries neither
    'trades' nor 'initial_capital', so the reference never produces the streak fields and always divides the
    recovery factor by 10 000."""
    out = dict(metrics)
    dd = metrics.get("max_drawdown", 0)
    out["calmar_ratio"] = (metrics.get("return_pct", 0) / 100) / dd if dd > 0 else float("inf")            # :244-250
    daily = list(metrics.get("daily_returns", {}).values())
    if daily:                                                                                               # :253-263
        neg = [r for r in daily if r < 0]
        downside = np.std(neg) if neg else 0
        out["sortino_ratio"] = (np.mean(daily) / downside) * np.sqrt(252) if downside > 0 else float("inf")
    else:
        out["sortino_ratio"] = 0
    out["recovery_factor"] = (metrics.get("net_profit", 0) / (dd * metrics.get("initial_capital", 10000))
                              if dd > 0 else float("inf"))                                                  # :291-296
    if metrics.get("total_trades", 0) > 0:                                                                  # :298-306
        wr = metrics.get("win_rate", 0)
        out["expectancy"] = wr * metrics.get("average_profit", 0) - (1 - wr) * abs(metrics.get("average_loss", 0))
    else:
        out["expectancy"] = 0
    out["profit_per_day"] = np.mean(daily) if daily else 0                                                  # :308-313
    return out


def strategy_score(metrics: Dict, goals: Dict) -> float:
    """:579-633.  `trades_per_day` is never produced by calculate_metrics, so the
    min_trades_per_day constraint never penalises; mirrored."""
    score = metrics.get(goals.get("primary", "sharpe_ratio"), 0)
    for name in goals.get("secondary", []):
        if name == "max_drawdown":
            score *= (1 - metrics.get("max_drawdown", 0))
        elif name == "win_rate":
            score *= (1 + metrics.get("win_rate", 0))
        elif name == "profit_factor":
            score *= (metrics.get("profit_factor", 1) / 2)
        elif name == "expectancy":
            score *= (1 + min(metrics.get("expectancy", 0) / 100, 1))
    floor = goals.get("constraints", {}).get("min_trades_per_day", 0)
    per_day = metrics.get("trades_per_day", floor)
    if per_day < floor:
        score *= per_day / floor
    return score


def market_points(price: np.ndarray, rsi: np.ndarray, symbol: str, minute0: int, bar_minutes: int = 1) -> List[Dict]:
    """fp32 series -> the reference's data-point dicts (values as Python floats)."""
    from datetime import timedelta
    t0 = datetime(1970, 1, 1) + timedelta(minutes=int(minute0))
    step = timedelta(minutes=int(bar_minutes))
    return [{"timestamp": (t0 + i * step).isoformat(), "symbol": symbol, "price": float(price[i]), "rsi": float(rsi[i])}
            for i in range(len(price))]


def lane_fitness(params: Dict, price: np.ndarray, rsi: np.ndarray, symbol: str, minute0: int,
                 goals: Dict, bar_minutes: int = 1) -> float:
    recs = simulate_trades(params, market_points(price, rsi, symbol, minute0, bar_minutes))
    return float(strategy_score(calculate_metrics(recs), goals))
