#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family once, at sizes a
40-100x slowdown still finishes.  The sweep legs force the speculative paths (no warm-up -> repair passes on the main
and on the side stream, a pool that overflows -> fused fallback) because that is where the bump allocator, the
linked event pool and the overlapped repair stream live.  Results are checked against the C oracle, so a sanitizer
run that "passes" also computed the right thing.

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from ai_crypto_trader_b200 import indicators as ind, synth
from ai_crypto_trader_b200.monte_carlo import PathEngine, risk_statistics
from ai_crypto_trader_b200.sweep import MarketData, PopulationSweep
from oracle import indicators_ref, sim_oracle

legs = sys.argv[1:] or ["sweep", "mc", "indicators"]
torch.cuda.set_device(0)

if "sweep" in legs:
    S, N, POP = 2, 40_000, 70
    ohlcv = synth.synth_ohlcv(S, N, first_symbol=1)
    market = MarketData(ohlcv)
    population = synth.random_population(POP, seed=11)
    population[0].update(rsi_oversold=35, rsi_overbought=65, rsi_period=5, take_profit=1, stop_loss=1)
    population[1].update(rsi_oversold=34, rsi_overbought=66, rsi_period=6, take_profit=10, stop_loss=5)
    cfg = sim_oracle.config_of(market.minute0, 1)
    banks = [indicators_ref.rsi_bank(ohlcv[3, s], list(range(5, 31))) for s in range(S)]
    want = np.array([[int(sim_oracle.lane(ohlcv[3, s], banks[s][p["rsi_period"] - 5], p, cfg)[0]["trade_hash"])
                      for s in range(S)] for p in population], dtype=np.uint64)
    for mode, opts in (("tiled", dict(warm=0, chunks=12, max_repair_rounds=64)),      # repaired chunk by chunk, side stream
                       ("tiled", dict(warm=1024, chunks=6)),                          # speculation mostly right
                       ("tiled", dict(warm=512, chunks=6, pool_blocks=8)),            # pool overflow -> fused fallback
                       ("tiled", dict(warm=1024, chunks=6, order_by="identity")),     # unpacked warps (> 2 RSI rows) -> flagged, fused fallback
                       ("chunked", dict(target_events=300, warm=0, max_chunks=16, max_repair_rounds=64)),
                       ("fused", {})):
        sw = PopulationSweep(market, event_cap=256, mode=mode, chunk_options=opts)
        sw.chunk_min_bars = 0
        for _ in range(2):                       # second sweep: zone map
            sw.evaluate(population)
        got = sw.lane_stats()["trade_hash"]
        assert np.array_equal(got, want), (mode, opts, int((got != want).sum()))
        print("sweep", mode, opts, "ok; invalid lanes", getattr(sw, "last_invalid_lanes", 0), flush=True)

if "mc" in legs:
    eng = PathEngine()
    f, d, paths = eng.gbm(100.0, 0.08, 0.35, 1 / 252, 5000, 64, 2024, store_paths=True)
    f2, d2, _ = eng.gbm(100.0, 0.08, 0.35, 1 / 252, 5000, 64, 2024)
    assert torch.equal(f, f2) and torch.equal(d, d2)
    ret = np.random.default_rng(3).normal(5e-4, 0.02, 60).astype(np.float32)
    eng.bootstrap(ret, True, 100.0, 3000, 45, 77, block_len=5)
    st = risk_statistics(eng, f, d, 100.0, 0.95)
    print("mc ok", st["var"], flush=True)

if "indicators" in legs:
    m = MarketData(synth.synth_ohlcv(2, 9000))
    ta = ind.TechnicalAnalyzer(m)
    ind.multi_timeframe_indicators(m, 0)
    print("indicators ok", ta.get_all_indicators(0)["rsi"], flush=True)
torch.cuda.synchronize()
print("sanitize target done")
