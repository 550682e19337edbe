"""GPU indicator kernels vs the float64 pandas restatement of `ta` (oracle/indicators_ref.py)."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

RTOL = 2e-6   # fp64 evaluation rounded once to fp32 vs float64 pandas (tolerance stated per north star: 1e-5)


@pytest.fixture(scope="module")
def gpu(native_lib):
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def _series(ohlcv, s):
    return [pd.Series(ohlcv[f, s].astype(np.float64)) for f in range(5)]


def _cmp(got, want, name, rtol=RTOL, atol=0.0):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, name
    assert np.array_equal(np.isnan(got), np.isnan(want)), (name, np.flatnonzero(np.isnan(got) != np.isnan(want))[:5])
    m = ~np.isnan(want)
    np.testing.assert_allclose(got[m], want[m], rtol=rtol, atol=atol, err_msg=name)


@pytest.mark.parametrize("n_bars", [1, 5, 19, 20, 60, 255, 2048, 2049, 5000, 70001])
def test_all_indicators_no_fill(gpu, n_bars):
    torch = gpu
    from ai_crypto_trader_b200 import indicators as ind, synth
    from oracle import indicators_ref as ref
    S = 2
    ohlcv = synth.synth_ohlcv(S, n_bars, first_symbol=2)
    o, h, l, c, v = [torch.from_numpy(ohlcv[f]).cuda() for f in range(5)]
    ema = ind.ema_bank(c, [3, 12, 26, 100], fill=False).cpu().numpy()
    sma = ind.sma_bank(c, [5, 20, 50, 200], fill=False).cpu().numpy()
    line, sig, diff = [t.cpu().numpy() for t in ind.macd(c, fill=False)]
    bb = [t.cpu().numpy() for t in ind.bollinger(c, fill=False)]
    sk, sd = [t.cpu().numpy() for t in ind.stochastic(h, l, c, fill=False)]
    wr = ind.williams_r(h, l, c, fill=False).cpu().numpy()
    ia, ib = [t.cpu().numpy() for t in ind.ichimoku(h, l, fill=False)]
    atr = ind.atr_bank(h, l, c, [7, 14, 25]).cpu().numpy()
    vw = ind.vwap(h, l, c, v, fill=False).cpu().numpy()
    for s in range(S):
        so, sh, sl, sc, sv = _series(ohlcv, s)
        for i, w in enumerate([3, 12, 26, 100]):
            _cmp(ema[s, i], ref.ema(sc, w), f"ema{w}")
        for i, w in enumerate([5, 20, 50, 200]):
            _cmp(sma[s, i], ref.sma(sc, w), f"sma{w}")
        rl, rs, rd = ref.macd(sc)
        _cmp(line[s], rl, "macd", atol=1e-9)
        _cmp(sig[s], rs, "macd_signal", atol=1e-9)
        _cmp(diff[s], rd, "macd_diff", rtol=1e-4, atol=2e-7)   # difference of two nearly equal fp32-rounded numbers
        rh, rm, rlo = ref.bollinger(sc)
        rw, rp = ref.bollinger_width_position(sc, rh, rm, rlo)
        for got, want, nm, tol in zip(bb, (rh, rm, rlo, rw, rp), ("bb_high", "bb_mid", "bb_low", "bb_width", "bb_pos"),
                                      (RTOL, RTOL, RTOL, 2e-5, 2e-5)):
            _cmp(got[s], want, nm, rtol=tol, atol=1e-7)
        rk, rdd = ref.stochastic(sh, sl, sc)
        _cmp(sk[s], rk, "stoch_k", atol=1e-5)
        _cmp(sd[s], rdd, "stoch_d", atol=1e-5)
        _cmp(wr[s], ref.williams_r(sh, sl, sc), "williams", atol=1e-5)
        ra, rb = ref.ichimoku(sh, sl)
        _cmp(ia[s], ra, "ichimoku_a")
        _cmp(ib[s], rb, "ichimoku_b")
        for i, w in enumerate([7, 14, 25]):
            _cmp(atr[s, i], ref.atr(sh, sl, sc, w), f"atr{w}", atol=1e-12)
        _cmp(vw[s], ref.vwap(sh, sl, sc, sv), "vwap")


def test_nanfill_matches_handle_nan_values(gpu):
    torch = gpu
    from ai_crypto_trader_b200.indicators import nanfill_
    from oracle.indicators_ref import handle_nan
    rng = np.random.default_rng(0)
    n = 20000
    x = rng.normal(size=(6, n)).astype(np.float32)
    x[0, :37] = np.nan                      # leading run
    x[1, 100:9000] = np.nan                 # run spanning two 4096-tiles
    x[2, :] = np.nan                        # all-NaN column -> 0
    x[3, rng.random(n) < 0.3] = np.nan      # scattered
    x[4, -5:] = np.nan                      # trailing
    x[5, :4096] = np.nan                    # exactly one leading tile
    got = nanfill_(torch.from_numpy(x.copy()).cuda()).cpu().numpy()
    for r in range(6):
        want = handle_nan(pd.Series(x[r].astype(np.float64))).to_numpy().astype(np.float32)
        assert np.array_equal(got[r], want), r


def test_flat_series_edge_cases(gpu):
    torch = gpu
    from ai_crypto_trader_b200 import indicators as ind
    from oracle import indicators_ref as ref
    n = 300
    c = np.full((1, n), 50.0, dtype=np.float32)
    c[0, 150:] += np.arange(150, dtype=np.float32) * 0.25
    h, l = c.copy(), c.copy()               # high == low == close: zero ranges
    v = np.ones_like(c)
    tc, th, tl, tv = [torch.from_numpy(a).cuda() for a in (c, h, l, v)]
    sc = pd.Series(c[0].astype(np.float64))
    bb = [t.cpu().numpy()[0] for t in ind.bollinger(tc, fill=True)]
    rh, rm, rlo = ref.bollinger(sc)
    rw, rp = ref.bollinger_width_position(sc, rh, rm, rlo)
    _cmp(bb[4], ref.handle_nan(rp), "bb_position filled", atol=1e-6)          # zero range -> NaN -> filled
    k, d = [t.cpu().numpy()[0] for t in ind.stochastic(th, tl, tc, fill=True)]
    rk, rd = ref.stochastic(sc, sc, sc)
    _cmp(k, ref.handle_nan(rk), "stoch_k filled", atol=1e-5)
    _cmp(d, ref.handle_nan(rd), "stoch_d filled", atol=1e-5)


def test_technical_analyzer_last_bar_scalars(gpu):
    """TechnicalAnalyzer.get_all_indicators vs the same scalars from the float64 oracle columns."""
    from ai_crypto_trader_b200 import synth
    from ai_crypto_trader_b200.indicators import TechnicalAnalyzer
    from ai_crypto_trader_b200.sweep import MarketData
    from oracle import indicators_ref as ref
    ohlcv = synth.synth_ohlcv(3, 10000)
    ta = TechnicalAnalyzer(MarketData(ohlcv))
    for s in range(3):
        so, sh, sl, sc, sv = _series(ohlcv, s)
        want = ref.analyzer_scalars(so, sh, sl, sc, sv)
        got = ta.get_all_indicators(s)
        assert got["trend"] == want["trend"]
        for k in ("rsi", "stoch_k", "stoch_d", "macd", "macd_signal", "williams_r", "bb_position", "volatility", "trend_strength"):
            assert got[k] == pytest.approx(want[k], rel=2e-5, abs=2e-6), (s, k)


@pytest.mark.parametrize("k,offset", [(5, 0), (15, 7), (3, 2)])
def test_resample_and_align(gpu, k, offset):
    """Derived k-minute bars equal pandas' clock-aligned resample; alignment uses completed bars only."""
    torch = gpu
    from ai_crypto_trader_b200 import indicators as ind, synth
    from ai_crypto_trader_b200.sweep import MarketData
    n = 5003
    ohlcv = synth.synth_ohlcv(2, n)
    minute0 = synth.EPOCH_2024_MINUTES + offset
    base = MarketData(ohlcv, minute0=minute0)
    hi = ind.resample(base, k)
    got = hi.ohlcv.cpu().numpy()
    idx = pd.to_datetime((minute0 + np.arange(n)) * 60, unit="s")
    for s in range(2):
        df = pd.DataFrame({f: ohlcv[i, s].astype(np.float64) for i, f in enumerate(synth.FIELDS)}, index=idx)
        r = df.resample(f"{k}min").agg({"open": "first", "high": "max", "low": "min", "close": "last", "volume": "sum"})
        assert got.shape[2] == len(r)
        for i, f in enumerate(("open", "high", "low", "close")):
            assert np.array_equal(got[i, s], r[f].to_numpy().astype(np.float32)), f
        np.testing.assert_allclose(got[4, s], r["volume"].to_numpy(), rtol=1e-6)
        # alignment: at base bar t, the last k-minute bar whose final minute is <= t
        al = ind.align_to_base(hi.close, base, k).cpu().numpy()[s]
        ends = (r.index.astype("datetime64[s]").astype(np.int64) // 60).to_numpy() + k - 1     # last minute of each bucket
        mins = minute0 + np.arange(n)
        j = np.searchsorted(ends, mins, side="right") - 1
        want = np.where(j >= 0, r["close"].to_numpy().astype(np.float32)[np.maximum(j, 0)], np.nan)
        assert np.array_equal(np.isnan(al), np.isnan(want)) and np.array_equal(al[~np.isnan(want)], want[~np.isnan(want)].astype(np.float32))


def test_multi_timeframe_recipe(gpu):
    from ai_crypto_trader_b200 import indicators as ind, synth
    from ai_crypto_trader_b200.sweep import MarketData
    from oracle import indicators_ref as ref
    n = 3000
    ohlcv = synth.synth_ohlcv(1, n)
    base = MarketData(ohlcv)
    got = ind.multi_timeframe_indicators(base, 0)
    idx = pd.to_datetime((synth.EPOCH_2024_MINUTES + np.arange(n)) * 60, unit="s")
    df = pd.DataFrame({f: ohlcv[i, 0].astype(np.float64) for i, f in enumerate(synth.FIELDS)}, index=idx)
    agg = {"open": "first", "high": "max", "low": "min", "close": "last", "volume": "sum"}
    d3, d5, d15 = (df.resample(f"{k}min").agg(agg) for k in (3, 5, 15))
    # resampled bars hold fp32 values on the device
    for d in (d3, d5, d15):
        for c in d.columns:
            d[c] = d[c].astype(np.float32).astype(np.float64)
    last = lambda s: float(s.iloc[-1])
    want = {"rsi": last(ref.rsi(df["close"])), "rsi_3m": last(ref.rsi(d3["close"])), "rsi_5m": last(ref.rsi(d5["close"])),
            "macd": last(ref.macd(df["close"])[0]), "macd_3m": last(ref.macd(d3["close"])[0]), "macd_5m": last(ref.macd(d5["close"])[0]),
            "stoch_k": last(ref.stochastic(df["high"], df["low"], df["close"])[0]),
            "williams_r": last(ref.williams_r(df["high"], df["low"], df["close"]))}
    s1 = (last(df["close"]) - last(ref.sma(df["close"], 20))) / last(ref.sma(df["close"], 20)) * 100
    s5 = (last(df["close"]) - last(ref.sma(d5["close"], 20))) / last(ref.sma(d5["close"], 20)) * 100
    want["trend_strength"] = abs(0.6 * s1 + 0.4 * s5)
    for k, v in want.items():
        assert got[k] == pytest.approx(v, rel=3e-5, abs=3e-6), k
    assert got["price_change_15m"] == pytest.approx((last(df["close"]) - last(d15["open"])) / last(d15["open"]) * 100, rel=1e-5)


@pytest.mark.parametrize("n_bars", [1, 5, 13, 19, 33, 60, 255, 2048, 2049, 5000, 70001])
def test_fused_analyzer_columns_match_the_reference_policy(gpu, n_bars):
    """TechnicalAnalyzer (three fused launches + one batched NaN-policy call) against the float64 restatement of
    _calculate_all_indicators + _handle_nan_values, all 21 columns, incl. series shorter than the windows (all-NaN -> 0),
    a flat stretch (zero ranges: stochastic / Williams / bb_position undefined in mid-series -> ffill)."""
    torch = gpu
    from ai_crypto_trader_b200 import synth
    from ai_crypto_trader_b200.indicators import TechnicalAnalyzer
    from ai_crypto_trader_b200.sweep import MarketData
    from oracle import indicators_ref as ref
    S = 2
    ohlcv = synth.synth_ohlcv(S, n_bars, first_symbol=6)
    if n_bars >= 255:
        a, b = n_bars // 3, n_bars // 3 + 70            # a flat stretch on symbol 1: open == high == low == close
        flat = ohlcv[3, 1, a]
        ohlcv[0:4, 1, a:b] = flat                       # (volume stays positive: a window without volume makes the reference's
                                                        #  VWAP 0/0 on pandas' rolling-sum residue, inf or garbage)
    ta = TechnicalAnalyzer(MarketData(ohlcv))
    assert set(ta.data) == set(TechnicalAnalyzer.COLUMNS)
    for s in range(S):
        so, sh, sl, sc, sv = _series(ohlcv, s)
        want = ref.analyzer_columns(so, sh, sl, sc, sv)
        for name, w in want.items():
            tol = dict(macd_diff=(1e-4, 2e-7), bb_width=(2e-5, 1e-7), bb_position=(2e-5, 1e-6), stoch_k=(RTOL, 1e-5), stoch_d=(RTOL, 1e-5),
                       williams_r=(RTOL, 1e-5), macd=(RTOL, 1e-9), macd_signal=(RTOL, 1e-9)).get(name, (RTOL, 1e-12))
            got = ta.data[name][s].cpu().numpy()
            assert not np.isnan(got).any(), (name, "NaN left after the policy")
            w = w.to_numpy()
            if name == "bb_position":
                # 0/0 where the 20-bar window is exactly flat in mid-series: pandas' online variance leaves a rounding residue
                # there (std ~1e-9 of the price, position = residue / residue = 0.5), the kernel's prefix differences give an exact zero range
                # (-> NaN -> ffill).  Neither is "the" value; compare where the range is a number.
                ill = (want["bb_high"] - want["bb_low"]).to_numpy() <= 1e-7 * want["bb_mid"].to_numpy()
                got, w = got[~ill], w[~ill]
            _cmp(got, w, f"{name} (symbol {s}, N={n_bars})", rtol=tol[0], atol=tol[1])
