"""Scalar decision helpers of the reference's strategy module (host side).

Reference: binance_ml_strategy.py -- TradingSignal (:470-581) and PositionSizer (:251-291);
services/ai_trader.py -- should_take_trade (:368-387), adjust_position_size (:389-418).
In the reference's backtest these consume whole-frame constants, so they run
once per symbol on the host; the per-bar loop that uses their outputs is the device kernel
b200bt_backtest_ref.  The two OpenAI calls of AITrader are network LLM calls (out of scope);
`DeterministicAITrader` is the documented stand-in with AITrader's method names.
"""
from __future__ import annotations

from typing import Dict, Optional


class TradingSignal:
    def __init__(self, symbol, price, rsi, stoch_k, macd, volume, volatility, williams_r=None, trend=None,
                 trend_strength=None, bb_position=None):
        self.symbol, self.price, self.rsi, self.stoch_k, self.macd = symbol, price, rsi, stoch_k, macd
        self.volume, self.volatility, self.williams_r = volume, volatility, williams_r
        self.trend, self.trend_strength, self.bb_position = trend, trend_strength, bb_position
        self.signal = self._calculate_signal()
        self.strength = self._calculate_strength()

    def _calculate_signal(self) -> str:
        votes = 0.0                      # each of six indicators contributes 3 (strong) or 2 (moderate)
        votes += 3.0 if self.rsi < 35 else (2.0 if self.rsi < 45 else 0.0)
        votes += 3.0 if self.stoch_k < 20 else (2.0 if self.stoch_k < 30 else 0.0)
        if self.macd > 0:
            votes += 3.0 if self.macd > self.macd * 1.1 else 2.0     # the 1.1x branch can never fire for macd > 0 (:509)
        if self.williams_r:              # 0.0 / None skip the test, as the reference's truthiness check does (:516)
            votes += 3.0 if self.williams_r < -80 else (2.0 if self.williams_r < -65 else 0.0)
        if self.trend == "uptrend" and self.trend_strength:
            votes += 3.0 if self.trend_strength > 10 else (2.0 if self.trend_strength > 5 else 0.0)
        if self.bb_position:
            votes += 3.0 if self.bb_position < 0.2 else (2.0 if self.bb_position < 0.4 else 0.0)
        ratio = votes / 6
        if ratio >= 0.6:
            return "BUY"
        if ratio <= 0.3:
            return "SELL"
        return "NEUTRAL"

    def _calculate_strength(self):
        if self.signal == "NEUTRAL":
            return 0
        buy = self.signal == "BUY"
        total = 0
        total += ((45 - min(self.rsi, 45)) / 15 if buy else (max(self.rsi, 55) - 55) / 15) * 30
        total += ((30 - min(self.stoch_k, 30)) / 30 if buy else (max(self.stoch_k, 70) - 70) / 30) * 20
        total += min(abs(self.macd), 1) * 20
        total += min(self.volume / 100000, 1) * 15
        if self.trend_strength:
            aligned = (buy and self.trend == "uptrend") or (not buy and self.trend == "downtrend")
            if aligned:
                total += min(self.trend_strength / 20, 1) * 15
        return min(max(total, 0), 100)


class PositionSizer:
    @staticmethod
    def volatility_class(volatility):
        """(position_pct, stop_loss_pct) by volatility (:256-264)."""
        if volatility > 0.02:
            return 0.25, 0.02
        if volatility > 0.01:
            return 0.20, 0.015
        return 0.15, 0.01

    @staticmethod
    def calculate_position_size(total_capital, volatility, volume, max_risk_per_trade=0.15) -> Dict:
        pct, stop = PositionSizer.volatility_class(volatility)
        size = total_capital * pct * min(volume / 50000, 1)
        size = min(size, (total_capital * max_risk_per_trade) / stop)
        size = min(size, total_capital * 0.20)
        size = max(size, total_capital * 0.10)
        size = max(size, 40)
        return {"position_size": size, "stop_loss_pct": stop, "take_profit_pct": stop * 2.0,
                "trailing_stop_activation": stop * 1.5, "trailing_stop_distance": stop * 0.75}


class DeterministicAITrader:
    """Stand-in for services/ai_trader.py:AITrader with its LLM calls replaced:
    analyze_trade_opportunity -> the technical signal with confidence 1.0,
    analyze_risk_setup -> None (no AI risk opinion: the tester keeps the technical sizing,
    strategy_tester.py:259-265).  should_take_trade / adjust_position_size keep the reference logic."""

    def __init__(self, config: Optional[Dict] = None):
        self.config = config or {"trading_params": {"ai_confidence_threshold": 0.7}}

    async def analyze_trade_opportunity(self, market_data: Dict) -> Dict:
        sig = TradingSignal(symbol=market_data["symbol"], price=market_data["current_price"], rsi=market_data["rsi"],
                            stoch_k=market_data["stoch_k"], macd=market_data["macd"], volume=market_data["avg_volume"],
                            volatility=market_data["volatility"], williams_r=market_data["williams_r"],
                            trend=market_data["trend"], trend_strength=market_data["trend_strength"],
                            bb_position=market_data["bb_position"])
        return {"decision": sig.signal, "confidence": 1.0, "reasoning": "deterministic stub"}

    async def analyze_risk_setup(self, risk_setup: Dict):
        return None

    def should_take_trade(self, analysis: Dict) -> bool:
        try:
            if analysis.get("decision") == "ERROR":
                return False
            if analysis.get("confidence", 0) < self.config["trading_params"]["ai_confidence_threshold"]:
                return False
            return analysis.get("decision") == "BUY"
        except Exception:
            return False

    def adjust_position_size(self, ai_position: Dict, technical_position: Dict) -> Dict:
        try:
            ai_size = float(ai_position.get("position_size", 0))
            tech_size = float(technical_position.get("position_size", 0))
            return {"position_size": (ai_size + tech_size) / 2,
                    "stop_loss_pct": max(float(ai_position.get("stop_loss_pct", 0)), float(technical_position.get("stop_loss_pct", 0))),
                    "take_profit_pct": min(float(ai_position.get("take_profit_pct", 0)), float(technical_position.get("take_profit_pct", 0))),
                    "reasoning": f"Combined AI ({ai_size:.2f}) and Technical ({tech_size:.2f}) analysis"}
        except Exception:
            return technical_position
