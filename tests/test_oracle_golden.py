"""The oracle restatements (Python and C) against fixtures produced by executing
the reference's own _simulate_trades / calculate_metrics / _calculate_strategy_score
(tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import unjson
from oracle import sim_oracle, simulate_ref


def _case_inputs(meta, arrays, case):
    sym = case["symbol"]
    period = case["params"].get("rsi_period", 14)
    return arrays[f"close_{sym}"], arrays[f"rsi_{sym}_{period}"]


def test_python_restatement_matches_reference_records(sim_golden):
    meta, arrays = sim_golden
    for case in meta["cases"]:
        close, rsi = _case_inputs(meta, arrays, case)
        pts = simulate_ref.market_points(close, rsi, f"SYN{case['symbol']:03d}USDT", meta["minute0"])
        recs = simulate_ref.simulate_trades(dict(case["params"]), pts)
        key = case["key"]
        assert len(recs) == case["n_records"], key
        ts_to_bar = {p["timestamp"]: i for i, p in enumerate(pts)}
        assert [ts_to_bar[r["timestamp"]] for r in recs] == arrays[f"bar_{key}"].tolist(), key
        assert [r["side"] == "sell" for r in recs] == arrays[f"sell_{key}"].tolist(), key
        # same float64 expressions -> bit-identical
        assert np.array_equal(np.array([r["pnl"] for r in recs]), arrays[f"pnl_{key}"]), key
        assert np.array_equal(np.array([r["quantity"] for r in recs]), arrays[f"qty_{key}"]), key
        assert np.array_equal(np.array([r["price"] for r in recs]), arrays[f"price_{key}"]), key
        m = simulate_ref.calculate_metrics(recs)
        for name, want in case["metrics"].items():
            want = unjson(want)
            got = float(m[name])
            assert got == want or (np.isnan(got) and np.isnan(want)), (key, name, got, want)
        assert np.array_equal(np.array(m["equity_curve"]), arrays[f"equity_{key}"]), key
        adv = simulate_ref.calculate_advanced_metrics(m)
        for name, want in case["advanced"].items():
            want, got = unjson(want), float(adv[name])
            assert got == want or (np.isnan(got) and np.isnan(want)), (key, name, got, want)
        got_score = float(simulate_ref.strategy_score(m, meta["goals"]))
        assert got_score == unjson(case["score"]), key


def test_c_oracle_matches_reference(sim_golden):
    meta, arrays = sim_golden
    cfg = sim_oracle.config_of(meta["minute0"], 1, meta["goals"])
    for case in meta["cases"]:
        close, rsi = _case_inputs(meta, arrays, case)
        key = case["key"]
        st, ev, pnl = sim_oracle.lane(close, rsi, case["params"], cfg, event_cap=8192)
        assert int(st["n_records"]) == case["n_records"], key
        assert (ev & 0x3FFFFFFF).tolist() == arrays[f"bar_{key}"].tolist(), key
        assert ((ev >> 31) == 1).tolist() == arrays[f"sell_{key}"].tolist(), key
        # same operations in the same order, -ffp-contract=off: bit-identical pnl
        assert np.array_equal(pnl, arrays[f"pnl_{key}"]), key
        m = case["metrics"]
        # Python's sum() is compensated (3.12), C accumulates plainly: 1e-12 relative
        for got, want in ((st["total_profit"], m["total_profit"]), (st["total_loss"], m["total_loss"]),
                          (st["net_profit"], m["net_profit"]), (st["win_rate"], m["win_rate"]),
                          (st["max_drawdown"], m["max_drawdown"]), (st["sharpe_ratio"], m["sharpe_ratio"]),
                          (st["largest_profit"], m["largest_profit"]), (st["largest_loss"], m["largest_loss"]),
                          (st["profit_factor"], m["profit_factor"]), (st["score"], case["score"])):
            want = unjson(want)
            assert got == pytest.approx(want, rel=1e-11, abs=1e-12), (key, got, want)
        assert int(st["n_days"]) == len(arrays[f"daily_{key}"]), key
        adv = case["advanced"]
        assert st["sortino_ratio"] == pytest.approx(unjson(adv["sortino_ratio"]), rel=1e-10, abs=1e-12), key
        assert st["mean_daily_pnl"] == pytest.approx(unjson(adv["profit_per_day"]), rel=1e-11, abs=1e-12), key
        assert int(st["n_negative_days"]) == int((arrays[f"daily_{key}"] < 0).sum()) * (case["n_records"] >= 2), key
        if case["n_records"]:
            dur = m["avg_trade_duration"] * (case["n_records"] // 2)
            assert st["sum_duration_bars"] == pytest.approx(dur, rel=1e-12), key


def test_c_oracle_event_cap_and_hash_are_consistent(sim_golden):
    meta, arrays = sim_golden
    cfg = sim_oracle.config_of(meta["minute0"], 1, meta["goals"])
    case = next(c for c in meta["cases"] if c["key"] == "fast_rsi_tight_0")
    close, rsi = _case_inputs(meta, arrays, case)
    st_a, ev_a, _ = sim_oracle.lane(close, rsi, case["params"], cfg, event_cap=16)
    st_b, ev_b, _ = sim_oracle.lane(close, rsi, case["params"], cfg, event_cap=4096)
    assert st_a["trade_hash"] == st_b["trade_hash"] != 0
    assert len(ev_a) == 16 and np.array_equal(ev_a, ev_b[:16])


def test_c_oracle_calendar_gap_matches_python_restatement(sim_golden):
    """A series glued from two pieces (training folds of the cross-validation): the C oracle with gap_bar / gap_minutes
    equals the Python restatement fed the glued points with their own timestamps."""
    meta, arrays = sim_golden
    case = next(c for c in meta["cases"] if c["name"] == "fast_rsi_tight")
    close, rsi = _case_inputs(meta, arrays, case)
    a, b = 1500, 4400                       # cut [a, b) out: the seam jumps ~2 days
    pts = simulate_ref.market_points(close, rsi, "X", meta["minute0"])
    glued = pts[:a] + pts[b:]
    recs = simulate_ref.simulate_trades(dict(case["params"]), glued)
    m = simulate_ref.calculate_metrics(recs)
    cfg = sim_oracle.config_of(meta["minute0"], 1, meta["goals"], gap_bar=a, gap_minutes=b - a)
    st, _, _ = sim_oracle.lane(np.concatenate([close[:a], close[b:]]), np.concatenate([rsi[:a], rsi[b:]]), case["params"], cfg)
    assert int(st["n_records"]) == len(recs)
    assert int(st["n_days"]) == len(m["daily_returns"])
    assert st["sharpe_ratio"] == pytest.approx(float(m["sharpe_ratio"]), rel=1e-10)
    assert st["score"] == pytest.approx(float(simulate_ref.strategy_score(m, meta["goals"])), rel=1e-10)


def _same(got, want):
    return got == pytest.approx(want, rel=1e-10, abs=1e-12) or (np.isnan(got) and np.isnan(want)) or got == want


def test_other_optimisation_goals_match_reference_scores(sim_golden):
    """_calculate_strategy_score under other goal sets, on the plain and on the advanced metrics dict (fixture
    `alt_scores` = the reference's own outputs): the C oracle's score rule and the Python restatement."""
    meta, arrays = sim_golden
    for g, alt in enumerate(meta["alt_goals"]):
        cfg = sim_oracle.config_of(meta["minute0"], 1, alt["goals"], advanced=alt["advanced"])
        for case in meta["cases"]:
            close, rsi = _case_inputs(meta, arrays, case)
            want = unjson(case["alt_scores"][g])
            st, _, _ = sim_oracle.lane(close, rsi, case["params"], cfg)
            assert _same(float(st["score"]), want), (alt, case["key"], float(st["score"]), want)
    case = meta["cases"][1]
    close, rsi = _case_inputs(meta, arrays, case)
    pts = simulate_ref.market_points(close, rsi, "SYN000USDT", meta["minute0"])
    m = simulate_ref.calculate_metrics(simulate_ref.simulate_trades(dict(case["params"]), pts))
    adv = simulate_ref.calculate_advanced_metrics(m)
    for g, alt in enumerate(meta["alt_goals"]):
        got = float(simulate_ref.strategy_score(adv if alt["advanced"] else m, alt["goals"]))
        assert _same(got, unjson(case["alt_scores"][g])), alt
