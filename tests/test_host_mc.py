"""Host arithmetic of the Monte-Carlo statistics that needs no GPU: NumPy's linear-interpolation percentile rule
rebuilt from exact order statistics (ai_crypto_trader_b200.monte_carlo._virtual_index / _lerp)."""
import numpy as np

from ai_crypto_trader_b200.monte_carlo import PERCENTILES, _lerp, _virtual_index


def test_percentile_rule_equals_numpy():
    rng = np.random.default_rng(12)
    for n in (1, 2, 3, 10, 101, 1000, 99_999):
        x = rng.standard_normal(n).astype(np.float32)
        srt = np.sort(x).astype(np.float64)
        for q in list(PERCENTILES) + [0.0, 100.0, 5.0, 0.1, 99.9, 33.333]:
            lo, hi, g = _virtual_index(n, q)
            assert 0 <= lo <= hi <= n - 1 and 0.0 <= g < 1.0
            got = _lerp(srt[lo], srt[hi], g)
            assert got == np.percentile(x.astype(np.float64), q), (n, q)
