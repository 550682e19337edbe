"""StrategyTester: the reference's single-symbol backtest behind its own call surface, on the GPU.

Reference: backtesting/strategy_tester.py (class StrategyTester).  Same methods, stats schema
(:45-61), trade record keys (:320-333) and quirks.  Where the time goes:
  * the 21 indicator columns of TechnicalAnalyzer are computed by the Family-1 kernels;
  * the technical x AI entry gate is evaluated ONCE per symbol on the host, because every input
    the reference feeds it is a whole-frame constant (:68-71,:106-117);
  * the per-bar loop (:190-300) and forced close (:303-307) run in b200bt_backtest_ref;
  * the host only formats the O(#trades + #equity points) output lists.
The reference calls OpenAI once per flat bar (:244); that network LLM is replaced by
strategy.DeterministicAITrader (or any object with AITrader's four methods whose answers do
not depend on the bar -- the GPU path evaluates the gate once).
"""
from __future__ import annotations

import ctypes as C
import json
import logging
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
import torch

from .. import _lib
from ..indicators import TechnicalAnalyzer
from ..strategy import DeterministicAITrader, PositionSizer, TradingSignal
from ..sweep import MarketData
from .data_manager import FIELDS, HistoricalDataManager

logger = logging.getLogger("b200bt.strategy_tester")

DEFAULT_TRADING_PARAMS = {"max_positions": 5, "ai_confidence_threshold": 0.7, "position_size": 0.4,
                          "stop_loss_pct": 2.0, "take_profit_pct": 4.0, "candle_interval": "1m"}
SOCIAL_DEFAULTS = {  # what SocialDataProvider returns with no social CSVs (social_data_provider.py:17-25,:201-232)
    "social_volume": 0, "social_engagement": 0, "social_contributors": 0, "social_sentiment": 0.5, "twitter_volume": 0,
    "reddit_volume": 0, "news_volume": 0, "news_sentiment": 0.5, "recent_news": [], "social_momentum": 0,
    "social_trend": "neutral", "social_intensity": 0, "social_engagement_rate": 0}
REASONS = {1: "Stop Loss", 2: "Take Profit", 3: "End of Test"}


class StrategyTester:
    def __init__(self, config_path: Optional[str] = "config.json", config: Optional[Dict] = None,
                 data_manager: Optional[HistoricalDataManager] = None, ai_trader=None,
                 results_dir: str = "backtesting/results"):
        if config is None:
            try:
                with open(config_path, "r") as f:
                    config = json.load(f)
            except (OSError, TypeError):
                config = {}
        self.config = config
        self.config.setdefault("trading_params", dict(DEFAULT_TRADING_PARAMS))
        self.trading_params = self.config["trading_params"]
        self.data_manager = data_manager or HistoricalDataManager(config_path)
        self.ai_trader = ai_trader or DeterministicAITrader(self.config)
        self.results_dir = Path(results_dir)
        self.results_dir.mkdir(parents=True, exist_ok=True)
        self.current_balance = 0.0
        self.reset_stats()

    def reset_stats(self):
        self.stats = {"initial_balance": 0.0, "final_balance": 0.0, "total_trades": 0, "winning_trades": 0,
                      "losing_trades": 0, "win_rate": 0.0, "profit_factor": 0.0, "total_profit": 0.0, "total_loss": 0.0,
                      "max_drawdown": 0.0, "max_drawdown_pct": 0.0, "sharpe_ratio": 0.0, "trades": [],
                      "equity_curve": [], "drawdown_curve": []}

    # -- the frame constants (prepare_market_data, :63-125) -----------------------------
    def _frame_constants(self, df: pd.DataFrame, market: MarketData) -> Dict:
        analyzer = TechnicalAnalyzer(market)
        ind = analyzer.get_all_indicators(0)
        avg_volume = df["volume"].mean() * df["close"].mean()                                   # :74
        close = df["close"]
        cur = float(close.iloc[-1])
        chg = lambda k: ((cur - float(close.iloc[-k - 1])) / float(close.iloc[-k - 1])) * 100 if len(df) > k else 0
        return {"avg_volume": avg_volume, "rsi": ind["rsi"], "stoch_k": ind["stoch_k"], "macd": ind["macd"],
                "williams_r": ind["williams_r"], "bb_position": ind["bb_position"], "trend": ind["trend"],
                "trend_strength": ind["trend_strength"], "volatility": ind["volatility"],
                "price_change_1m": chg(1), "price_change_3m": chg(3), "price_change_5m": chg(5), "price_change_15m": chg(15)}

    @staticmethod
    def _market_of(df: pd.DataFrame) -> MarketData:
        ohlcv = np.stack([df[f].to_numpy(dtype=np.float32) for f in FIELDS])[:, None, :]
        minute0 = int(pd.Timestamp(df.index[0]).timestamp() // 60)
        return MarketData(np.ascontiguousarray(ohlcv), minute0=minute0)

    def prepare_market_data(self, df: pd.DataFrame, symbol: str) -> List[Dict]:
        consts = self._frame_constants(df, self._market_of(df))
        out = []
        for idx, price in zip(df.index, df["close"]):
            d = {"symbol": symbol, "current_price": float(price), "timestamp": idx.isoformat()}
            d.update(consts)
            d.update(SOCIAL_DEFAULTS)
            out.append(d)
        return out

    async def analyze_with_ai(self, market_data: Dict) -> Dict:
        md = dict(market_data)
        analysis = await self.ai_trader.analyze_trade_opportunity(md)
        if self.ai_trader.should_take_trade(analysis):
            risk_setup = {"symbol": md["symbol"], "available_capital": self.trading_params.get("position_size_pct", 0.4) * self.current_balance,
                          "volatility": md["volatility"], "current_price": md["current_price"], "trend_strength": md["trend_strength"]}
            return {"trade_analysis": analysis, "risk_analysis": await self.ai_trader.analyze_risk_setup(risk_setup)}
        return {"trade_analysis": analysis, "risk_analysis": None}

    def should_execute_trade(self, technical_signal: TradingSignal, ai_analysis: Dict) -> Dict:
        try:
            if not technical_signal or not ai_analysis:
                return {"execute": False}
            if ai_analysis["trade_analysis"]["confidence"] < self.config["trading_params"]["ai_confidence_threshold"]:
                return {"execute": False}
            if technical_signal.strength < 70:
                return {"execute": False}
            decision = ai_analysis["trade_analysis"]["decision"]
            if technical_signal.signal != decision:
                return {"execute": False}
            return {"execute": True, "decision": decision}
        except Exception as e:
            logger.error("Error in trade decision: %s", e)
            return {"execute": False}

    # -- the backtest ------------------------------------------------------------------
    async def backtest_strategy(self, symbol: str, interval: str, start_date: datetime, end_date: datetime = None,
                                initial_balance: float = 10000.0) -> Dict:
        self.reset_stats()
        self.current_balance = initial_balance
        self.stats["initial_balance"] = initial_balance
        self.stats["equity_curve"].append({"timestamp": start_date.isoformat(), "equity": initial_balance})
        df = self.data_manager.merge_market_and_social_data(symbol, interval, start_date, end_date)
        if df.empty:
            logger.error("No data available for %s from %s to %s", symbol, start_date, end_date)
            return self.stats
        return await self.backtest_frame(df, symbol, initial_balance)

    async def backtest_frame(self, df: pd.DataFrame, symbol: str, initial_balance: float = 10000.0) -> Dict:
        """The body of backtest_strategy for an in-memory OHLCV frame (timestamp index)."""
        if not self.stats["equity_curve"]:
            self.reset_stats()
            self.current_balance = initial_balance
            self.stats["initial_balance"] = initial_balance
            self.stats["equity_curve"].append({"timestamp": df.index[0].isoformat(), "equity": initial_balance})
        market = self._market_of(df)
        consts = self._frame_constants(df, market)
        update = dict(consts, symbol=symbol, current_price=float(df["close"].iloc[-1]), timestamp=df.index[-1].isoformat())
        signal = TradingSignal(symbol=symbol, price=update["current_price"], rsi=update["rsi"], stoch_k=update["stoch_k"],
                               macd=update["macd"], volume=update["avg_volume"], volatility=update["volatility"],
                               williams_r=update["williams_r"], trend=update["trend"],
                               trend_strength=update["trend_strength"], bb_position=update["bb_position"])
        ai = await self.analyze_with_ai(update)
        gate = self.should_execute_trade(signal, ai)
        can_enter = bool(gate["execute"] and gate.get("decision") == "BUY")                         # :249
        tech = PositionSizer.calculate_position_size(initial_balance, update["volatility"], update["avg_volume"])
        if ai["risk_analysis"]:
            raise NotImplementedError("an AI risk opinion changes the sizing per entry; only risk_analysis=None "
                                      "(DeterministicAITrader) is supported on the GPU path")
        pct, stop = PositionSizer.volatility_class(update["volatility"])
        prm = _lib.BtParams(initial_balance=float(initial_balance), position_pct=pct, stop_loss_pct=tech["stop_loss_pct"],
                            take_profit_pct=tech["take_profit_pct"], volume_factor=float(min(update["avg_volume"] / 50000, 1)),
                            max_risk_per_trade=0.15, can_enter=1 if can_enter else 0, skip=10)
        n = market.N
        dev = market.device
        prm_dev = torch.frombuffer(bytearray(bytes(prm)), dtype=torch.uint8).to(dev)
        stats_dev = torch.zeros(16, dtype=torch.float64, device=dev)
        trades_dev = torch.zeros((n + 1, 8), dtype=torch.float64, device=dev)
        equity_dev = torch.zeros((n + 1, 2), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.call("b200bt_backtest_ref", market.close.data_ptr(), _lib.ld(market.close), 1, n, prm_dev.data_ptr(),
                      stats_dev.data_ptr(), trades_dev.data_ptr(), n + 1, equity_dev.data_ptr(), n + 1, _lib.current_stream())
        s = stats_dev.cpu().numpy()
        n_tr, n_eq = int(s[1]), int(s[8])
        tr = trades_dev[:n_tr].cpu().numpy()
        eq = equity_dev[:n_eq].cpu().numpy()
        # host materialisation of the reference's result schema (:320-333, :166-171, :222-236), column-wise
        idx = df.index
        if len(idx) and bool((idx.microsecond == 0).all()) and bool((idx.nanosecond == 0).all()):
            times = np.asarray(idx.strftime("%Y-%m-%dT%H:%M:%S"))        # == Timestamp.isoformat() for whole seconds
        else:
            times = np.array([t.isoformat() for t in idx], dtype=object)
        st = self.stats
        e_bar, x_bar = tr[:, 0].astype(np.int64), tr[:, 1].astype(np.int64)
        exit_price = df["close"].to_numpy().astype(np.float32).astype(np.float64)[x_bar]      # market_price(): the fp32 close
        sl, tp = tech["stop_loss_pct"], tech["take_profit_pct"]
        reasons = [REASONS[int(r)] for r in tr[:, 2].astype(np.int64).tolist()]
        st["trades"].extend(
            {"symbol": symbol, "entry_price": ep, "entry_time": et, "quantity": q, "position_size": ps,
             "stop_loss_pct": sl, "take_profit_pct": tp, "exit_price": xp, "exit_time": xt, "pnl": pnl, "pnl_pct": pp,
             "exit_reason": why}
            for ep, et, q, ps, xp, xt, pnl, pp, why in zip(tr[:, 3].tolist(), times[e_bar].tolist(), tr[:, 4].tolist(),
                                                           tr[:, 5].tolist(), exit_price.tolist(), times[x_bar].tolist(),
                                                           tr[:, 6].tolist(), tr[:, 7].tolist(), reasons))
        bal = eq[:, 1]
        peak = np.maximum.accumulate(np.concatenate([[float(initial_balance)], bal]))[1:]     # running max incl. the start
        dd = peak - bal
        ts = times[eq[:, 0].astype(np.int64)].tolist()
        st["equity_curve"].extend({"timestamp": t, "equity": b} for t, b in zip(ts, bal.tolist()))
        st["drawdown_curve"].extend({"timestamp": t, "drawdown": d, "drawdown_pct": p}
                                    for t, d, p in zip(ts, dd.tolist(), ((dd / peak) * 100).tolist()))
        st["total_trades"], st["winning_trades"], st["losing_trades"] = int(s[1]), int(s[2]), int(s[3])
        st["total_profit"], st["total_loss"] = float(s[4]), float(s[5])
        st["max_drawdown"], st["max_drawdown_pct"] = float(s[6]), float(s[7])
        self.current_balance = float(s[0])
        self.calculate_final_stats()
        self.frame_constants = dict(consts, signal=signal.signal, strength=signal.strength, can_enter=can_enter)
        return st

    def calculate_final_stats(self):
        st = self.stats
        st["final_balance"] = self.current_balance
        if st["total_trades"] > 0:
            st["win_rate"] = (st["winning_trades"] / st["total_trades"]) * 100
        if st["total_loss"] > 0:
            st["profit_factor"] = st["total_profit"] / st["total_loss"]
        rets, prev = [], st["initial_balance"]
        for p in st["equity_curve"]:                       # "daily" returns over the recorded points (:417-423)
            rets.append((p["equity"] - prev) / prev if prev > 0 else 0)
            prev = p["equity"]
        if len(rets) > 1:
            sd = np.std(rets)
            if sd > 0:
                st["sharpe_ratio"] = (np.mean(rets) / sd) * np.sqrt(252)

    def save_results(self, strategy_name: str, symbol: str, interval: str, start_date: datetime, end_date: datetime = None) -> str:
        if end_date is None:
            end_date = datetime.now()
        path = self.results_dir / f"{strategy_name}_{symbol}_{interval}_{start_date.strftime('%Y%m%d')}_{end_date.strftime('%Y%m%d')}.json"
        with open(path, "w") as f:
            json.dump({"strategy": strategy_name, "symbol": symbol, "interval": interval, "start_date": start_date.isoformat(),
                       "end_date": end_date.isoformat(), "stats": self.stats}, f, indent=2, default=float)
        return str(path)

    async def run_multiple_backtests(self, symbols: List[str], intervals: List[str], start_date: datetime,
                                     end_date: datetime = None, initial_balance: float = 10000.0) -> Dict:
        results = {}
        for symbol in symbols:
            per = {}
            for interval in intervals:
                try:
                    r = await self.backtest_strategy(symbol, interval, start_date, end_date, initial_balance)
                    self.save_results("AI_Social_Strategy", symbol, interval, start_date, end_date)
                    per[interval] = r
                except Exception as e:
                    logger.error("Error in backtest for %s on %s: %s", symbol, interval, e)
                    per[interval] = {"error": str(e)}
            results[symbol] = per
        return results


def market_price(df: pd.DataFrame, bar: int) -> float:
    return float(np.float32(df["close"].iloc[bar]))
