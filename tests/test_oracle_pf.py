"""The portfolio-risk oracle against the fixture produced by the reference's own methods."""
import json
from pathlib import Path

import numpy as np
import pytest

from ai_crypto_trader_b200 import synth
from oracle import portfolio_ref

GOLD = Path(__file__).parent / "golden"


def test_oracle_matches_reference_fixture():
    ref = json.loads((GOLD / "pf_reference.json").read_text())
    close = synth.synth_ohlcv(ref["S"], ref["N"])[3].astype(np.float64)
    close[1] = close[1] * (close[0] / close[0][0]) ** 0.5
    close[2] = close[2] * (close[0][0] / close[0]) ** 0.25
    close = close.astype(np.float32).astype(np.float64)
    r = portfolio_ref.pct_change(close)
    r[3, ref["holes"]] = np.nan
    for conf, want in ref["var"].items():
        assert [portfolio_ref.value_at_risk(r[i], float(conf), 1000.0) for i in range(ref["S"])] == pytest.approx(want, rel=1e-12)
    for conf, want in ref["cvar"].items():
        assert [portfolio_ref.conditional_value_at_risk(r[i], float(conf), 1000.0) for i in range(ref["S"])] == pytest.approx(want, rel=1e-12)
    corr = portfolio_ref.correlation(r)
    np.testing.assert_allclose(corr, np.array(ref["correlation"]), atol=1e-13)
    pv = portfolio_ref.portfolio_var(np.array(ref["holdings_values"]), np.array(ref["var"]["0.95"]) / 1000.0, corr, 6000.0)
    assert pv == pytest.approx(ref["portfolio_var"], rel=1e-12)
