"""Portfolio risk kernels against the reference fixture and the CPU oracle."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import portfolio_ref

GOLD = Path(__file__).parent / "golden"
pytestmark = pytest.mark.gpu


def _fixture_close():
    from ai_crypto_trader_b200 import synth
    ref = json.loads((GOLD / "pf_reference.json").read_text())
    close = synth.synth_ohlcv(ref["S"], ref["N"])[3].astype(np.float64)
    close[1] = close[1] * (close[0] / close[0][0]) ** 0.5
    close[2] = close[2] * (close[0][0] / close[0]) ** 0.25
    return ref, close.astype(np.float32)


def _service(ref, close):
    import torch
    from ai_crypto_trader_b200.portfolio_risk import PortfolioRiskService, ReturnBank
    bank = ReturnBank.from_close(ref["symbols"], close)
    bank.returns[3, torch.tensor(ref["holes"], device=bank.returns.device)] = float("nan")
    return PortfolioRiskService(bank), bank


def test_returns_bitwise():
    ref, close = _fixture_close()
    _, bank = _service(ref, close)
    got = bank.returns.cpu().numpy()
    want = portfolio_ref.pct_change(close.astype(np.float64)).astype(np.float32)
    want[3, ref["holes"]] = np.nan
    assert np.array_equal(got, want, equal_nan=True)


def test_var_cvar_against_reference():
    ref, close = _fixture_close()
    svc, _ = _service(ref, close)
    for conf, want in ref["var"].items():
        got = [svc.calculate_var(s, float(conf), 1000.0) for s in ref["symbols"]]
        np.testing.assert_allclose(got, want, rtol=2e-6)
    for conf, want in ref["cvar"].items():
        got = [svc.calculate_conditional_var(s, float(conf), 1000.0) for s in ref["symbols"]]
        np.testing.assert_allclose(got, want, rtol=2e-6)
    assert svc.calculate_var(np.array([0.1]), 0.95) == 0.0          # < 2 points (:235)


def test_correlation_and_portfolio_var_against_reference():
    ref, close = _fixture_close()
    svc, _ = _service(ref, close)
    corr = svc.calculate_asset_correlation(ref["symbols"] + ["NOPEUSDC"])
    got = np.array([[corr[a][b] for b in ref["symbols"]] for a in ref["symbols"]])
    np.testing.assert_allclose(got, np.array(ref["correlation"]), atol=5e-7)
    assert corr["NOPEUSDC"][ref["symbols"][0]] == 0.0
    names = [s.replace("USDC", "") for s in ref["symbols"]]
    svc.asset_correlations = {a: {b: got[i, j] for j, b in enumerate(names)} for i, a in enumerate(names)}
    holdings = {"assets": {a: {"value_usdc": v} for a, v in zip(names, ref["holdings_values"])}, "total_value": 6000.0}
    holdings["assets"]["USDC"] = {"value_usdc": 500.0}
    var_est = {a: v / 1000.0 for a, v in zip(names, ref["var"]["0.95"])}
    assert svc.calculate_portfolio_var(holdings, var_est) == pytest.approx(ref["portfolio_var"], rel=1e-6)


@pytest.mark.parametrize("S,N", [(1, 17), (3, 1000), (50, 200_003), (130, 4099)])
def test_correlation_against_oracle(S, N):
    import torch
    from ai_crypto_trader_b200.portfolio_risk import PortfolioRiskService, ReturnBank
    rng = np.random.default_rng(S * 1000 + N)
    base = rng.standard_normal(N)
    x = (0.001 * (rng.standard_normal((S, N)) + np.linspace(-1, 1, S)[:, None] * base)).astype(np.float32)
    x[rng.random((S, N)) < 0.01] = np.nan
    if S > 2:
        x[2] = 0.25          # constant row -> NaN correlations
    bank = ReturnBank([f"a{i}" for i in range(S)], torch.from_numpy(x).cuda())
    got = PortfolioRiskService(bank).correlation_matrix()
    want = portfolio_ref.correlation(x)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(want), atol=1e-9)


def test_tail_stats_large():
    import torch
    from ai_crypto_trader_b200.portfolio_risk import PortfolioRiskService
    rng = np.random.default_rng(5)
    x = rng.standard_normal(3_000_001).astype(np.float32) * 0.02
    x[::1000] = np.nan
    svc = PortfolioRiskService(device="cuda")
    for conf in (0.95, 0.999):
        assert svc.calculate_var(x, conf, 5.0) == pytest.approx(portfolio_ref.value_at_risk(x, conf, 5.0), rel=1e-6)
        assert svc.calculate_conditional_var(x, conf, 5.0) == pytest.approx(
            portfolio_ref.conditional_value_at_risk(x, conf, 5.0), rel=1e-6)
