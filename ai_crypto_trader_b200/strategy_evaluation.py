"""StrategyEvaluationSystem / StrategyPerformanceMetrics call surface on the GPU engine.

Reference: services/strategy_evaluation.py.
  StrategyEvaluationSystem._simulate_trades(strategy_id, parameters, market_data)   :746-878
  StrategyPerformanceMetrics.calculate_metrics(trades, initial_capital)             :32-228
  StrategyEvaluationSystem._calculate_strategy_score(metrics)                       :579-633

`_simulate_trades` keeps the reference's signature and returns the reference's trade
records (one dict per entry and per exit, keys timestamp/symbol/side/price/quantity/
fees/pnl); the bar-by-bar state machine runs in the sweep kernel (one lane), the host
only formats the O(#trades) records.  `evaluate_population` is the batched form the GA
uses.  Market data contract: price and rsi are fp32 (values are rounded to fp32 on
upload; DESIGN.md "Data contract").
"""
from __future__ import annotations

import json
import logging
from datetime import datetime
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .sweep import DEFAULT_GOALS, MarketData, PopulationSweep

logger = logging.getLogger("b200bt.strategy_evaluation")


class StrategyPerformanceMetrics:
    """Host-side metrics over an arbitrary list of trade records (e.g. live trades handed in
    by a caller).  The GA path never calls this: there the same quantities are reduced inside
    the sweep kernel.  Restates calculate_metrics (:32-228) for the scalar fields."""

    @staticmethod
    def calculate_metrics(trades: List[Dict], initial_capital: float = 10000.0) -> Dict:
        base = {"total_trades": 0, "win_rate": 0.0, "profit_factor": 0.0, "sharpe_ratio": 0.0, "max_drawdown": 0.0,
                "average_profit": 0.0, "average_loss": 0.0, "largest_profit": 0.0, "largest_loss": 0.0,
                "total_profit": 0.0, "total_loss": 0.0, "net_profit": 0.0, "return_pct": 0.0,
                "avg_trade_duration": 0, "risk_reward_ratio": 0.0, "profitable_symbols": {},
                "unprofitable_symbols": {}, "monthly_returns": {}, "daily_returns": {}}
        if not trades:
            return base
        recs = sorted(trades, key=lambda t: t.get("timestamp", ""))
        pnl = [t.get("pnl", 0) for t in recs]
        if len(recs) < 2:
            p = pnl[0]
            base.update(total_trades=1, win_rate=1.0 if p > 0 else 0.0, net_profit=p,
                        return_pct=(p / initial_capital) * 100)
            for key, cond in (("average_profit", p > 0), ("largest_profit", p > 0), ("total_profit", p > 0),
                              ("average_loss", p < 0), ("largest_loss", p < 0), ("total_loss", p < 0)):
                base[key] = p if cond else 0.0
            return base
        wins = [p for p in pnl if p > 0]
        losses = [p for p in pnl if p < 0]
        tp, tl = sum(wins), sum(losses)
        equity, peak, dds = [initial_capital], initial_capital, []
        daily, monthly, by_symbol = {}, {}, {}
        for t, p in zip(recs, pnl):
            cur = equity[-1] + p
            equity.append(cur)
            if cur > peak:
                peak = cur
            else:
                dds.append((peak - cur) / peak)
            ts = t.get("timestamp", "")
            if ts:
                try:
                    d = datetime.fromisoformat(ts.replace("Z", "+00:00"))
                    for key, bucket in ((d.strftime("%Y-%m-%d"), daily), (d.strftime("%Y-%m"), monthly)):
                        bucket[key] = bucket[key] + p if key in bucket else p
                except ValueError:
                    pass
            sym = t.get("symbol", "UNKNOWN")
            by_symbol[sym] = by_symbol.get(sym, 0) + p
        durs = []
        for i in range(0, len(recs) - 1, 2):
            try:
                a = datetime.fromisoformat(recs[i].get("timestamp", "").replace("Z", "+00:00"))
                b = datetime.fromisoformat(recs[i + 1].get("timestamp", "").replace("Z", "+00:00"))
                durs.append((b - a).total_seconds() / 60)
            except ValueError:
                pass
        vals = list(daily.values())
        sharpe = 0
        if len(vals) > 1:
            sd = np.std(vals)
            sharpe = (np.mean(vals) / sd) * np.sqrt(252) if sd > 0 else 0
        avg_p = tp / len(wins) if wins else 0
        avg_l = tl / len(losses) if losses else 0
        net = tp + tl
        return {
            "total_trades": len(recs), "win_rate": len(wins) / len(recs),
            "profit_factor": abs(tp / tl) if tl != 0 else float("inf"), "sharpe_ratio": sharpe,
            "max_drawdown": max(dds) if dds else 0, "average_profit": avg_p, "average_loss": avg_l,
            "largest_profit": max(wins) if wins else 0, "largest_loss": min(losses) if losses else 0,
            "total_profit": tp, "total_loss": tl, "net_profit": net, "return_pct": (net / initial_capital) * 100,
            "avg_trade_duration": sum(durs) / len(durs) if durs else 0,
            "risk_reward_ratio": abs(avg_p / avg_l) if avg_l != 0 else float("inf"),
            "profitable_symbols": {s: v for s, v in by_symbol.items() if v > 0},
            "unprofitable_symbols": {s: v for s, v in by_symbol.items() if v <= 0},
            "monthly_returns": monthly, "daily_returns": daily, "equity_curve": equity,
        }


    @staticmethod
    def calculate_advanced_metrics(metrics: Dict) -> Dict:
        """calculate_advanced_metrics (:231-319): Calmar, Sortino, recovery factor, expectancy, profit per day
        from a calculate_metrics dict; the streak fields appear only when the caller added a 'trades' list (:267-288)."""
        import numpy as np
        out = dict(metrics)
        dd = metrics.get("max_drawdown", 0)
        out["calmar_ratio"] = (metrics.get("return_pct", 0) / 100) / dd if dd > 0 else float("inf")
        daily = list(metrics.get("daily_returns", {}).values())
        if daily:
            neg = [r for r in daily if r < 0]
            downside = np.std(neg) if neg else 0
            out["sortino_ratio"] = (np.mean(daily) / downside) * np.sqrt(252) if downside > 0 else float("inf")
        else:
            out["sortino_ratio"] = 0
        trades = metrics.get("trades", []) if metrics.get("total_trades", 0) > 0 else []
        if trades:
            win = loss = best_win = best_loss = 0
            for t in sorted(trades, key=lambda x: x.get("timestamp", "")):
                if t.get("pnl", 0) > 0:
                    win, loss = win + 1, 0
                    best_win = max(best_win, win)
                else:
                    loss, win = loss + 1, 0
                    best_loss = max(best_loss, loss)
            out["max_consecutive_wins"], out["max_consecutive_losses"] = best_win, best_loss
        out["recovery_factor"] = (metrics.get("net_profit", 0) / (dd * metrics.get("initial_capital", 10000))
                                  if dd > 0 else float("inf"))
        if metrics.get("total_trades", 0) > 0:
            wr = metrics.get("win_rate", 0)
            out["expectancy"] = wr * metrics.get("average_profit", 0) - (1 - wr) * abs(metrics.get("average_loss", 0))
        else:
            out["expectancy"] = 0
        out["profit_per_day"] = np.mean(daily) if daily else 0
        return out


class StrategyEvaluationSystem:
    def __init__(self, config_path: Optional[str] = "config.json", config: Optional[Dict] = None):
        if config is None:
            try:
                with open(config_path, "r") as f:
                    config = json.load(f)
            except (OSError, TypeError):
                config = {}
        self.config = config
        self.optimization_goals = self.config.get("evolution", {}).get("optimization_goals", {}) or dict(DEFAULT_GOALS)
        self.performance_metrics = self.config.get("evolution", {}).get("performance_metrics", {})

    # -- reference-compatible single-strategy call -----------------------------------
    def _simulate_trades(self, strategy_id: str, parameters: Dict, market_data: List[Dict]) -> List[Dict]:
        if not market_data:
            return []
        n = len(market_data)
        price = np.array([d.get("price", 50000) for d in market_data], dtype=np.float32)
        rsi = np.array([d.get("rsi", 50) for d in market_data], dtype=np.float32)
        ohlcv = np.zeros((5, 1, n), dtype=np.float32)
        ohlcv[3, 0] = price
        market = MarketData(ohlcv)
        period = int(parameters.get("rsi_period", 14))
        sweep = PopulationSweep.from_bank(market, torch.from_numpy(rsi).to(market.device).view(1, 1, n), [period],
                                          self.optimization_goals, event_cap=n + 1)
        sweep.evaluate([dict(parameters)])
        n_rec = int(sweep.lane_stats()["n_records"][0, 0])
        ev = sweep.events()[0, 0, :n_rec]
        size = 10000 * (min(parameters.get("max_position_size", 5), 20) / 100)     # :761-764
        out, entry = [], 0.0
        for w in ev.tolist():
            bar, is_exit, sell = w & _lib.EVENT_BAR_MASK, bool(w & _lib.EVENT_EXIT), bool(w & _lib.EVENT_SELL)
            d = market_data[bar]
            px = float(price[bar])
            if not is_exit:
                entry = px
                qty, pnl = size / px, -size * 0.001
            else:
                qty = size / entry
                pnl = qty * ((px - entry) if sell else (entry - px)) - size * 0.002
            out.append({"timestamp": d.get("timestamp", f"2023-01-{bar + 1:02d}T00:00:00Z"),
                        "symbol": d.get("symbol", "BTCUSDT"), "side": "sell" if sell else "buy", "price": px,
                        "quantity": qty, "fees": size * 0.001, "pnl": pnl})
        return out

    # -- k-fold cross-validation (:635-744) -----------------------------------------------------------
    CV_METRICS = ("sharpe_ratio", "win_rate", "max_drawdown", "profit_factor", "return_pct")

    def cross_validate_population(self, population: List[Dict], close: torch.Tensor, bank: torch.Tensor, periods,
                                  minute0: int, bar_minutes: int = 1, k_folds: int = 5,
                                  initial_capital: float = 10000.0) -> Dict[str, np.ndarray]:
        """The numeric core of cross_validate_strategy for a whole population at once.

        close [S][N] fp32 and bank [S][P][N] (RSI rows for `periods`) are device tensors.  The bars are cut into
        k contiguous folds (:659-666); for each fold every individual is simulated on the fold ("test") and on
        the other folds glued together ("train", :674-682 -- the state machine runs straight across the seam,
        as the reference's does, and the records keep their own calendar days).  Returns arrays [k][pop][S] for
        the lane metrics `CV_METRICS` and `score`, prefixed train_ / test_."""
        S, N = close.shape
        if N < k_folds:
            k_folds = N
        fold = N // k_folds
        bounds = [(i * fold, (i + 1) * fold if i < k_folds - 1 else N) for i in range(k_folds)]
        out = {f"{side}_{m}": np.zeros((k_folds, len(population), S)) for side in ("train", "test")
               for m in self.CV_METRICS + ("score", "n_records")}

        def run(side, f, c, bk, m0, gap_bar=0, gap_minutes=0):
            market = MarketData.from_close(c.contiguous(), minute0=m0, bar_minutes=bar_minutes)
            sweep = PopulationSweep.from_bank(market, bk.contiguous(), periods, self.optimization_goals,
                                              initial_capital=initial_capital, mode="auto")
            sweep.set_gap(gap_bar, gap_minutes)
            sweep.evaluate(population)
            st = sweep.lane_stats()
            for m in ("sharpe_ratio", "win_rate", "max_drawdown", "profit_factor", "score", "n_records"):
                out[f"{side}_{m}"][f] = st[m]
            out[f"{side}_return_pct"][f] = st["net_profit"] / initial_capital * 100

        for f, (a, b) in enumerate(bounds):
            run("test", f, close[:, a:b], bank[:, :, a:b], minute0 + a * bar_minutes)
            if a == 0 and b == N:
                continue                                               # k = 1: nothing to train on (:672 gives [])
            if a == 0:
                run("train", f, close[:, b:], bank[:, :, b:], minute0 + b * bar_minutes)
            elif b == N:
                run("train", f, close[:, :a], bank[:, :, :a], minute0)
            else:
                run("train", f, torch.cat([close[:, :a], close[:, b:]], dim=1), torch.cat([bank[:, :, :a], bank[:, :, b:]], dim=2),
                    minute0, gap_bar=a, gap_minutes=(b - a) * bar_minutes)
        out["fold_bounds"] = np.array(bounds)
        return out

    @staticmethod
    def _summarize_market_conditions(market_data: List[Dict]) -> Dict:
        """:880-935 (host; a fold's worth of points)."""
        if not market_data:
            return {"trend": "unknown", "volatility": 0, "volume": 0, "period_start": None, "period_end": None}
        prices = [d.get("price", 0) for d in market_data if d.get("price", 0) > 0]
        volumes = [d.get("volume", 0) for d in market_data if d.get("volume", 0) > 0]
        trend, volatility = "unknown", 0
        if len(prices) >= 2:
            change = (prices[-1] - prices[0]) / prices[0] if prices[0] > 0 else 0
            trend = "uptrend" if change > 0.05 else "downtrend" if change < -0.05 else "ranging"
            p = np.asarray(prices, dtype=np.float64)
            volatility = np.std((p[1:] - p[:-1]) / p[:-1])
        return {"trend": trend, "volatility": float(volatility), "volume": float(np.mean(volumes) if volumes else 0),
                "period_start": market_data[0].get("timestamp"), "period_end": market_data[-1].get("timestamp")}

    @staticmethod
    def _calculate_cv_summary(fold_results: List[Dict], test_metric: str, normalize: bool) -> Dict:
        """:937-990."""
        tr = [f["train_score"] for f in fold_results]
        te = [f["test_score"] for f in fold_results]
        summary = {"mean_train_score": float(np.mean(tr)), "std_train_score": float(np.std(tr)),
                   "mean_test_score": float(np.mean(te)), "std_test_score": float(np.std(te)),
                   "train_test_gap": float(np.mean(tr) - np.mean(te)),
                   "relative_overfitting": float((np.mean(tr) - np.mean(te)) / np.mean(tr)) if np.mean(tr) > 0 else 0}
        for side in ("train", "test"):
            for metric in (fold_results[0][f"{side}_metrics"] if fold_results else {}):
                vals = [f[f"{side}_metrics"][metric] for f in fold_results]
                summary[f"mean_{side}_{metric}"] = float(np.mean(vals))
                summary[f"std_{side}_{metric}"] = float(np.std(vals))
        summary["consistency"] = (1.0 - min(1.0, summary["std_test_score"] / summary["mean_test_score"])
                                  if summary["mean_test_score"] > 0 else 0.0)
        return summary

    def cross_validate_strategy(self, strategy_id: str, parameters: Dict, market_data_periods: List[Dict],
                                k_folds: int = 5, test_metric: str = "sharpe_ratio",
                                normalize_results: bool = True) -> Dict:
        """Drop-in for cross_validate_strategy (:635-744) on a flat list of market-data points (keys timestamp,
        symbol, price, rsi[, volume]) on a uniform time grid.  JSON dump and plots (:732-739) are out of scope."""
        pts = market_data_periods
        if len(pts) < k_folds:
            k_folds = len(pts)
        n = len(pts)
        t0 = datetime.fromisoformat(pts[0]["timestamp"].replace("Z", "+00:00"))
        minute0 = int((t0.replace(tzinfo=None) - datetime(1970, 1, 1)).total_seconds() // 60)   # calendar days as written (:151)
        step = 1
        if n > 1:
            t1 = datetime.fromisoformat(pts[1]["timestamp"].replace("Z", "+00:00"))
            step = max(1, int(round((t1 - t0).total_seconds() / 60)))
        dev = torch.device("cuda", torch.cuda.current_device())
        close = torch.from_numpy(np.array([d.get("price", 50000) for d in pts], dtype=np.float32)).to(dev).view(1, n)
        bank = torch.from_numpy(np.array([d.get("rsi", 50) for d in pts], dtype=np.float32)).to(dev).view(1, 1, n)
        period = int(parameters.get("rsi_period", 14))
        cv = self.cross_validate_population([dict(parameters)], close, bank, [period], minute0, step, k_folds)
        fold_results = []
        for f, (a, b) in enumerate(cv["fold_bounds"].tolist()):
            block = {}
            for side in ("train", "test"):
                m = {k: float(cv[f"{side}_{k}"][f, 0, 0]) for k in self.CV_METRICS}
                m["test_metric"] = m.get(test_metric, 0)
                block[f"{side}_metrics"] = m
                block[f"{side}_score"] = float(cv[f"{side}_score"][f, 0, 0])
            fold_results.append({"fold": f + 1, **block, "market_conditions": self._summarize_market_conditions(pts[a:b])})
        return {"strategy_id": strategy_id, "parameters": parameters, "k_folds": len(fold_results),
                "test_metric": test_metric,
                "cv_summary": self._calculate_cv_summary(fold_results, test_metric, normalize_results),
                "fold_results": fold_results, "timestamp": datetime.now().isoformat()}

    def _calculate_strategy_score(self, metrics: Dict) -> float:
        goals = self.optimization_goals
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

    # -- batched form -------------------------------------------------------------
    def evaluate_population(self, population: List[Dict], market: MarketData,
                            sweep: Optional[PopulationSweep] = None) -> np.ndarray:
        """fitness[i] = mean over symbols of score(metrics(simulate(population[i], symbol)))."""
        sweep = sweep or PopulationSweep(market, optimization_goals=self.optimization_goals)
        return sweep.evaluate(population)
