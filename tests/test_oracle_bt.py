"""configs[0] oracle (oracle/tester_ref.py) vs the reference's own StrategyTester run
(tests/golden/bt_reference.*, produced by tests/golden/make_golden.py bt)."""
import json
import importlib.util
import sys

import numpy as np
import pytest

from conftest import GOLDEN, unjson
from oracle import tester_ref


def load_bt_golden():
    meta = json.loads((GOLDEN / "bt_reference.json").read_text())["cases"]
    arrays = np.load(GOLDEN / "bt_reference.npz")
    spec = importlib.util.spec_from_file_location("make_golden_mod", GOLDEN / "make_golden.py")
    return meta, arrays


def bt_frame(case):
    """Same frames as make_golden.bt_frame (kept in sync by test: the reference numbers only match if they are)."""
    import pandas as pd
    from ai_crypto_trader_b200 import synth
    n = case["n"]
    d = synth.synth_symbol(case["symbol"], n)
    cols = {k: d[k].astype(np.float64).copy() for k in synth.FIELDS}
    if case["crash"]:
        f = np.ones(n)
        f[-60:] = np.linspace(1.0, 0.90, 60)
        for k in ("open", "high", "low", "close"):
            cols[k] *= f
    cols["volume"] *= case["vol_scale"]
    idx = pd.date_range("2024-01-01", periods=n, freq="min")
    return pd.DataFrame({k: v.astype(np.float32).astype(np.float64) for k, v in cols.items()}, index=idx)  # fp32 market data, widened (CSV-loaded frames are float64)


def test_oracle_backtest_matches_reference_strategy_tester():
    meta, arrays = load_bt_golden()
    for case in meta:
        df = bt_frame(case)
        st = tester_ref.backtest(df)
        name = case["name"]
        c = st["constants"]
        assert c["signal"] == case["signal"] and c["strength"] == unjson(case["strength"]), name
        for k, want in case["constants"].items():
            if k.startswith("price_change"):
                continue
            got = c[k]
            assert (got == want) if isinstance(want, str) else got == pytest.approx(unjson(want), rel=1e-13), (name, k)
        bar = {t.isoformat(): i for i, t in enumerate(df.index)}
        tr = st["trades"]
        assert len(tr) == case["n_trades"], name
        assert [bar[t["entry_time"]] for t in tr] == arrays[f"entry_bar_{name}"].tolist()
        assert [bar[t["exit_time"]] for t in tr] == arrays[f"exit_bar_{name}"].tolist()
        code = {"Stop Loss": 1, "Take Profit": 2, "End of Test": 3}
        assert [code[t["exit_reason"]] for t in tr] == arrays[f"reason_{name}"].tolist()
        for k in ("entry_price", "quantity", "position_size", "pnl", "pnl_pct"):
            assert np.array_equal(np.array([t[k] for t in tr], dtype=np.float64), arrays[f"{k}_{name}"]), (name, k)
        assert np.array_equal(np.array([p["equity"] for p in st["equity_curve"]]), arrays[f"eq_{name}"]), name
        assert [bar[p["timestamp"]] for p in st["equity_curve"]] == arrays[f"eq_bar_{name}"].tolist()
        dd = np.array([[p["drawdown"], p["drawdown_pct"]] for p in st["drawdown_curve"]]).reshape(-1, 2)
        assert np.array_equal(dd, arrays[f"dd_{name}"]), name
        for k, want in case["stats"].items():
            assert float(st[k]) == unjson(want), (name, k)
