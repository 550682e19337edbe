"""Summary report of a batch of backtests (plots are out of scope: matplotlib).

Reference: backtesting/result_analyzer.py:226-328 (generate_summary_report), :413-427."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List


class ResultAnalyzer:
    def __init__(self, results_dir: str = "backtesting/results"):
        self.results_dir = Path(results_dir)
        self.results_dir.mkdir(parents=True, exist_ok=True)
        self.plots_dir = self.results_dir / "plots"

    def generate_summary_report(self, results: List[Dict]) -> Dict:
        if not results:
            return {}
        rows, strategies, symbols, intervals = [], set(), set(), set()
        tot_wr = tot_pf = tot_sh = 0
        total_trades = profitable = 0
        best = worst = None
        best_ret, worst_ret = -float("inf"), float("inf")
        for i, r in enumerate(results):
            st = r.get("stats", {})
            strategies.add(r.get("strategy", "Unknown")); symbols.add(r.get("symbol", "Unknown")); intervals.add(r.get("interval", "Unknown"))
            ib, fb = st.get("initial_balance", 0), st.get("final_balance", 0)
            ret = ((fb / ib) - 1) * 100 if ib > 0 else 0
            profitable += 1 if fb > ib else 0
            if ret > best_ret:
                best_ret, best = ret, i
            if ret < worst_ret:
                worst_ret, worst = ret, i
            total_trades += st.get("total_trades", 0)
            tot_wr += st.get("win_rate", 0); tot_pf += st.get("profit_factor", 0); tot_sh += st.get("sharpe_ratio", 0)
            rows.append({"strategy": r.get("strategy", "Unknown"), "symbol": r.get("symbol", "Unknown"),
                         "interval": r.get("interval", "Unknown"), "trades": st.get("total_trades", 0),
                         "win_rate": st.get("win_rate", 0), "profit_factor": st.get("profit_factor", 0),
                         "sharpe_ratio": st.get("sharpe_ratio", 0), "return_pct": ret,
                         "max_drawdown": st.get("max_drawdown_pct", 0), "file_path": r.get("file_path", "")})
        n = len(results)
        return {"strategies": list(strategies), "symbols": list(symbols), "intervals": list(intervals), "total_results": n,
                "average_win_rate": tot_wr / n, "average_profit_factor": tot_pf / n, "average_sharpe_ratio": tot_sh / n,
                "best_result": results[best] if best is not None else None,
                "worst_result": results[worst] if worst is not None else None,
                "total_trades": total_trades, "profitable_strategies": profitable, "results": rows}

    def save_summary_report(self, summary: Dict, filename: str = "backtest_summary.json") -> str:
        path = self.results_dir / filename
        with open(path, "w") as f:
            json.dump(summary, f, indent=2, default=float)
        return str(path)
