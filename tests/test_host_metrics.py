"""Host-side StrategyPerformanceMetrics.calculate_metrics / score (product package) vs the
reference's numbers for the same trade records."""
import numpy as np

from conftest import unjson
from ai_crypto_trader_b200 import synth
from ai_crypto_trader_b200.strategy_evaluation import StrategyEvaluationSystem, StrategyPerformanceMetrics


def test_metrics_and_score_match_reference(sim_golden):
    meta, arrays = sim_golden
    ses = StrategyEvaluationSystem(config={"evolution": {"optimization_goals": meta["goals"]}})
    for c in meta["cases"]:
        key = c["key"]
        recs = [{"timestamp": synth.bar_timestamp(int(b), meta["minute0"]), "symbol": "X", "side": "sell" if s else "buy",
                 "price": float(p), "quantity": float(q), "fees": 0.5, "pnl": float(x)}
                for b, s, p, q, x in zip(arrays[f"bar_{key}"], arrays[f"sell_{key}"], arrays[f"price_{key}"],
                                         arrays[f"qty_{key}"], arrays[f"pnl_{key}"])]
        m = StrategyPerformanceMetrics.calculate_metrics(recs)
        for name, want in c["metrics"].items():
            assert float(m[name]) == unjson(want), (key, name)
        if recs:
            assert np.array_equal(np.array(m["equity_curve"]), arrays[f"equity_{key}"])
        assert float(ses._calculate_strategy_score(m)) == unjson(c["score"]), key
        adv = StrategyPerformanceMetrics.calculate_advanced_metrics(m)
        for name, want in c["advanced"].items():
            assert float(adv[name]) == unjson(want), (key, name)
        for g, alt in enumerate(meta["alt_goals"]):             # other goal sets, plain / advanced dict
            ses2 = StrategyEvaluationSystem(config={"evolution": {"optimization_goals": alt["goals"]}})
            got, want = float(ses2._calculate_strategy_score(adv if alt["advanced"] else m)), unjson(c["alt_scores"][g])
            assert got == want or (np.isnan(got) and np.isnan(want)), (key, alt)
        assert "max_consecutive_wins" not in adv              # the metrics dict carries no 'trades' (:267)
        streaks = StrategyPerformanceMetrics.calculate_advanced_metrics(dict(m, trades=recs))
        if recs and m["total_trades"] > 0:
            assert streaks["max_consecutive_wins"] <= 1 and streaks["max_consecutive_losses"] >= 1   # entries always lose the fee
    one = StrategyPerformanceMetrics.calculate_metrics([{"timestamp": "2024-01-01T00:00:00", "pnl": 3.0}])
    assert one["total_trades"] == 1 and one["win_rate"] == 1.0 and one["net_profit"] == 3.0
