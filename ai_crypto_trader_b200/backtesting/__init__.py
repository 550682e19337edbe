"""Drop-in for the reference's `backtesting` package (run_backtest.py:11:
`from backtesting import BacktestEngine, ResultAnalyzer`)."""
from .backtest_engine import BacktestEngine
from .data_manager import HistoricalDataManager
from .result_analyzer import ResultAnalyzer
from .strategy_tester import StrategyTester

__all__ = ["BacktestEngine", "HistoricalDataManager", "ResultAnalyzer", "StrategyTester"]
