#!/bin/bash
set -u
O=gpurun_out/r2q; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_indicators.py tests/test_gpu_backtest.py -q -x > $O/pytest_ind.log 2>&1; echo "rc=$?" >> $O/pytest_ind.log

python - > $O/analyzer_bench.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
from ai_crypto_trader_b200 import synth
from ai_crypto_trader_b200.indicators import TechnicalAnalyzer
from ai_crypto_trader_b200.sweep import MarketData
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if __import__('os').path.exists('MEASURED_PEAKS.json') else 6583.5
for S in (10, 50):
    m = MarketData(synth.synth_ohlcv(S, 1_000_000)).materialise()
    torch.cuda.synchronize()
    for _ in range(3): ta = TechnicalAnalyzer(m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ta = TechnicalAnalyzer(m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (4 + 21) * 4 * S * 1e6 / 1e9       # high, low, close, volume read once + 21 columns written
    print(json.dumps({"symbols": S, "bars": 1000000, "ms": ms, "algorithmic_GB": gb, "GBps": gb / (ms * 1e-3), "frac_of_hbm_peak": gb / (ms * 1e-3) / peak,
                      "note": "TechnicalAnalyzer(market): 3 fused launches + batched NaN policy + last-bar readback (one small D2H)"}))
PY
tail -n 15 $O/*.log
