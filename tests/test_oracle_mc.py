"""Monte-Carlo oracle vs fixtures produced by the reference's own MonteCarloService."""
import json

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mc_ref


@pytest.fixture(scope="module")
def mc_golden():
    return json.loads((GOLDEN / "mc_reference.json").read_text()), np.load(GOLDEN / "mc_reference.npz")


def _close(a, b, tol=1e-12):
    if isinstance(a, dict):
        assert a.keys() == b.keys()
        for k in a:
            _close(a[k], b[k], tol)
    else:
        assert a == pytest.approx(b, rel=tol, abs=tol), (a, b)


def test_statistics_restatement_matches_reference(mc_golden):
    meta, arrays = mc_golden
    for case in meta["cases"]:
        paths = arrays[f"paths_{case['key']}"]
        ref = case["result"]
        assert paths.shape == (case["days"], case["n"])
        got = mc_ref.risk_statistics(paths[-1], mc_ref.path_drawdowns(paths), ref["initial_price"], 0.95)
        for k in ("percentiles", "expected", "risk_metrics"):
            _close(got[k], ref[k])
        sp = {"volatile": {"drift_factor": 1.0, "volatility_factor": 2.0},
              "bear": {"drift_factor": 0.5, "volatility_factor": 1.2}}.get(case["scenario"], {})
        mu, sigma = mc_ref.drift_and_vol(arrays["returns"], sp)
        assert mu == pytest.approx(ref["mu"], rel=1e-14) and sigma == pytest.approx(ref["sigma"], rel=1e-14)


def test_philox_known_answer():
    # Random123 known-answer vectors for Philox4x32-10
    out = mc_ref.philox4x32_10(np.uint32([0]), np.uint32([0]), np.uint32([0]), np.uint32([0]), 0, 0)
    assert [int(o[0]) for o in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = mc_ref.philox4x32_10(np.uint32([0xffffffff]), np.uint32([0xffffffff]), np.uint32([0xffffffff]),
                               np.uint32([0xffffffff]), 0xffffffff, 0xffffffff)
    assert [int(o[0]) for o in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = mc_ref.philox4x32_10(np.uint32([0x243f6a88]), np.uint32([0x85a308d3]), np.uint32([0x13198a2e]),
                               np.uint32([0x03707344]), 0xa4093822, 0x299f31d0)
    assert [int(o[0]) for o in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_oracle_generator_agrees_statistically_with_reference(mc_golden):
    """Different RNG streams, same law: finals of the oracle GBM vs the reference's own
    paths agree within Monte-Carlo error (3 sigma of the standard error of the mean log return)."""
    meta, arrays = mc_golden
    case = meta["cases"][0]
    ref_paths = arrays[f"paths_{case['key']}"]
    r = case["result"]
    finals, maxdd, _ = mc_ref.gbm_paths(100.0, r["mu"], r["sigma"], 1 / 252, 20000, case["days"] - 1, seed=5)
    lr_ref = np.log(ref_paths[-1] / 100.0)
    lr = np.log(finals.astype(np.float64) / 100.0)
    se = np.sqrt(lr.var() / len(lr) + lr_ref.var() / len(lr_ref))
    assert abs(lr.mean() - lr_ref.mean()) < 3.5 * se
    assert lr.std() == pytest.approx(lr_ref.std(), rel=0.12)
    assert maxdd.mean() == pytest.approx(mc_ref.path_drawdowns(ref_paths).mean(), rel=0.12)


def test_bootstrap_oracle_draws_only_sample_values():
    ret = np.array([0.01, -0.02, 0.03], dtype=np.float32)
    finals, maxdd, logS = mc_ref.bootstrap_paths(ret, True, 50.0, 64, 9, seed=11)
    inc = np.diff(logS, axis=0)
    assert np.all(np.isclose(inc[..., None], ret.astype(np.float64), atol=1e-12).any(-1))
    # block bootstrap of length 3 walks the sample circularly
    _, _, logS3 = mc_ref.bootstrap_paths(ret, True, 50.0, 8, 6, seed=3, block_len=3)
    inc3 = np.diff(logS3, axis=0)
    for p in range(8):
        idx = [int(np.argmin(abs(ret - v))) for v in inc3[:, p]]
        assert idx[1] == (idx[0] + 1) % 3 and idx[2] == (idx[0] + 2) % 3 and idx[4] == (idx[3] + 1) % 3


def test_portfolio_stats_host_logic_matches_reference(mc_golden):
    """_calculate_portfolio_stats is host arithmetic in the product too (no kernel involved)."""
    import importlib
    meta, _ = mc_golden
    mod = importlib.import_module("ai_crypto_trader_b200.monte_carlo")
    svc = mod.MonteCarloService()
    got = svc._calculate_portfolio_stats(meta["portfolio"]["holdings"], meta["portfolio"]["simulations"])
    assert got == meta["portfolio"]["stats"]
