"""BacktestEngine: the reference's orchestration surface over the GPU StrategyTester.

Reference: backtesting/backtest_engine.py (class BacktestEngine): run_backtest (:64-125),
run_multiple_backtests (:127-178), get_available_data (:180-199), the asyncio task queue
(:217-304).  Error convention kept: never raise, return {'error': str} (:85,:91,:125).
Network fetches (fetch_data_for_backtest) and plots are out of scope."""
from __future__ import annotations

import asyncio
import json
import logging
from datetime import datetime
from typing import Dict, List, Optional

from .data_manager import HistoricalDataManager
from .result_analyzer import ResultAnalyzer
from .strategy_tester import StrategyTester

logger = logging.getLogger("b200bt.backtest_engine")


class BacktestEngine:
    def __init__(self, config_path: Optional[str] = "config.json", config: Optional[Dict] = None,
                 data_dir: Optional[str] = None, results_dir: str = "backtesting/results"):
        if config is None:
            try:
                with open(config_path, "r") as f:
                    config = json.load(f)
            except (OSError, TypeError):
                config = {}
        self.config = config
        self.data_manager = HistoricalDataManager(config_path, data_dir=data_dir)
        self.strategy_tester = StrategyTester(config_path, config=config, data_manager=self.data_manager,
                                              results_dir=results_dir)
        self.result_analyzer = ResultAnalyzer(results_dir)
        self.task_queue: asyncio.Queue = asyncio.Queue()
        self.running_tasks = set()

    async def fetch_data_for_backtest(self, symbol, intervals, start_date, end_date=None, include_social=True) -> Dict:
        """The reference fetches from Binance / LunarCrush here (:33-62): network, out of scope."""
        return {i: {"market_data": False, "social_data": False, "error": "network fetch is out of scope"} for i in intervals}

    async def run_backtest(self, symbol: str, interval: str, start_date: datetime, end_date: datetime = None,
                           initial_balance: float = 10000.0, save_results: bool = True) -> Dict:
        try:
            if end_date is None:
                end_date = datetime.now()
            market_data = self.data_manager.load_market_data(symbol, interval, start_date, end_date)
            if market_data.empty:
                return {"error": "No market data available and data fetch failed"}
            result = await self.strategy_tester.backtest_strategy(symbol, interval, start_date, end_date, initial_balance)
            if save_results:
                result["result_path"] = self.strategy_tester.save_results("AI_Social_Strategy", symbol, interval, start_date, end_date)
            return result
        except Exception as e:
            logger.error("Error running backtest for %s (%s): %s", symbol, interval, e)
            return {"error": str(e)}

    async def run_multiple_backtests(self, symbols: List[str], intervals: List[str], start_date: datetime,
                                     end_date: datetime = None, initial_balance: float = 10000.0) -> Dict:
        results = {}
        for symbol in symbols:
            per = {}
            for interval in intervals:
                try:
                    per[interval] = dict(await self.run_backtest(symbol, interval, start_date, end_date, initial_balance))
                except Exception as e:
                    per[interval] = {"error": str(e)}
            results[symbol] = per
        flat = [{"strategy": "AI_Social_Strategy", "symbol": s, "interval": i, "stats": r}
                for s, per in results.items() for i, r in per.items() if "error" not in r]
        if flat:
            summary = self.result_analyzer.generate_summary_report(flat)
            results["summary"] = {"path": self.result_analyzer.save_summary_report(summary),
                                  "profitable_strategies": summary.get("profitable_strategies", 0),
                                  "total_results": summary.get("total_results", 0)}
        return results

    def get_available_data(self) -> Dict:
        out = {}
        for symbol in self.data_manager.available_symbols():
            info = {"intervals": {}}
            for interval in self.data_manager.available_intervals(symbol):
                a, b = self.data_manager.get_data_range(symbol, interval)
                if a and b:
                    info["intervals"][interval] = {"start_date": a.isoformat(), "end_date": b.isoformat(), "days": (b - a).days}
            out[symbol] = info
        return out

    async def add_backtest_task(self, task_type: str, params: Dict) -> int:
        task_id = len(self.running_tasks) + self.task_queue.qsize() + 1
        await self.task_queue.put({"id": task_id, "type": task_type, "params": params, "status": "queued",
                                   "created_at": datetime.now().isoformat()})
        return task_id

    async def process_task_queue(self, stop_when_empty: bool = False):
        while True:
            if stop_when_empty and self.task_queue.empty():
                return
            task = await self.task_queue.get()
            try:
                task["status"] = "running"
                self.running_tasks.add(task["id"])
                p = task["params"]
                if task["type"] == "run_backtest":
                    result = await self.run_backtest(p.get("symbol"), p.get("interval"), datetime.fromisoformat(p.get("start_date")),
                                                     datetime.fromisoformat(p["end_date"]) if p.get("end_date") else None,
                                                     p.get("initial_balance", 10000.0))
                elif task["type"] == "run_multiple_backtests":
                    result = await self.run_multiple_backtests(p.get("symbols", []), p.get("intervals", []),
                                                               datetime.fromisoformat(p.get("start_date")),
                                                               datetime.fromisoformat(p["end_date"]) if p.get("end_date") else None,
                                                               p.get("initial_balance", 10000.0))
                else:
                    result = {"error": f"Unknown task type: {task['type']}"}
                task.update(status="completed", completed_at=datetime.now().isoformat(), result=result)
            except Exception as e:
                task.update(status="failed", error=str(e), completed_at=datetime.now().isoformat())
            finally:
                self.running_tasks.discard(task["id"])
                self.task_queue.task_done()

    async def run(self):
        processor = asyncio.create_task(self.process_task_queue())
        try:
            while True:
                await asyncio.sleep(1)
        except asyncio.CancelledError:
            pass
        finally:
            processor.cancel()
