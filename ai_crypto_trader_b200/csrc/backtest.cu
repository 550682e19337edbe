// The reference's single-symbol bar loop (StrategyTester.backtest_strategy,
// backtesting/strategy_tester.py:156-312) as a device state machine, one warp per symbol.
//
// In the reference every indicator fed to the entry test is a whole-frame constant
// (strategy_tester.py:68-71,106-117), so the technical + AI gate is one
// boolean per symbol (`can_enter`), computed once on the host from the GPU indicator
// columns.  What remains per bar is exactly this loop:
//   :192      skip the first `skip` bars
//   :202-218  exit test in PERCENT against thresholds that PositionSizer returns as
//             FRACTIONS (quirk kept): stop-loss first, then take-profit
//   :221-222  still in the position -> next bar WITHOUT an equity point
//   :247-277  entry: PositionSizer.calculate_position_size(balance, ...) (binance_ml_strategy.py:251-291)
//   :280-300  equity / drawdown point (cash only, no mark-to-market)
//   :303-307  forced close at the last bar
// Arithmetic is float64 in the reference's operation order.  The 32 lanes fetch 32 bars
// with one coalesced load; the machine itself is serial in time (each bar depends on the
// balance and position left by the previous one), so all lanes step it redundantly and
// lane 0 writes the records.  Symbols run in parallel (one CTA each).
#include <math.h>
#include "common.cuh"

namespace b200bt {

__device__ __forceinline__ double sizer(double capital, const b200bt_bt_params& q) {
    // binance_ml_strategy.py:265-283
    double size = capital * q.position_pct * q.volume_factor;
    const double max_position = (capital * q.max_risk_per_trade) / q.stop_loss_pct;
    size = fmin(size, max_position);
    size = fmin(size, capital * 0.20);
    size = fmax(size, capital * 0.10);
    size = fmax(size, 40.0);
    return size;
}

__global__ void __launch_bounds__(32)
backtest_ref_kernel(const float* __restrict__ price, int64_t ld, int64_t N, const b200bt_bt_params* __restrict__ params,
                    double* __restrict__ stats, double* __restrict__ trades, int64_t trade_cap,
                    double* __restrict__ equity, int64_t equity_cap) {
    const int sym = blockIdx.x, lane = threadIdx.x;
    const b200bt_bt_params q = params[sym];
    const float* row = price + (int64_t)sym * ld;
    double* tr = trades + (int64_t)sym * trade_cap * 8;
    double* eq = equity + (int64_t)sym * equity_cap * 2;

    double balance = q.initial_balance, max_equity = q.initial_balance;
    double max_dd = 0.0, max_dd_pct = 0.0, total_profit = 0.0, total_loss = 0.0;
    long long n_trades = 0, n_win = 0, n_lose = 0, n_eq = 0;
    bool in_pos = false;
    double entry = 0.0, qty = 0.0, size = 0.0;
    long long entry_bar = 0;

    auto close_pos = [&](double px, long long bar, int reason) {
        const double pnl = (px - entry) * qty;                          // :343
        const double pnl_pct = ((px - entry) / entry) * 100.0;          // :344
        balance += pnl;
        if (lane == 0 && n_trades < trade_cap) {
            double* r = tr + n_trades * 8;
            r[0] = (double)entry_bar; r[1] = (double)bar; r[2] = (double)reason; r[3] = entry;
            r[4] = qty; r[5] = size; r[6] = pnl; r[7] = pnl_pct;
        }
        ++n_trades;
        if (pnl > 0.0) { ++n_win; total_profit += pnl; } else { ++n_lose; total_loss -= pnl; }   // :362-367
        in_pos = false;
    };

    for (int64_t base = 0; base < N; base += 32) {
        const int64_t tl = base + lane;
        const float pv = tl < N ? __ldg(row + tl) : 0.f;
        const int lim = (int)min((int64_t)32, N - base);
        for (int j = 0; j < lim; ++j) {
            const int64_t t = base + j;
            const double px = (double)__shfl_sync(FULL, pv, j);
            if (t < q.skip) continue;                                   // :192
            if (in_pos) {
                const double pnl_pct = ((px - entry) / entry) * 100.0;  // :206
                if (pnl_pct <= -q.stop_loss_pct) close_pos(px, t, 1);   // :209
                else if (pnl_pct >= q.take_profit_pct) close_pos(px, t, 2);   // :215
            }
            if (in_pos) continue;                                       // :221-222
            if (q.can_enter) {                                          // :247-277
                size = sizer(balance, q);
                entry = px;
                qty = size / px;
                entry_bar = t;
                in_pos = true;
            }
            if (lane == 0 && n_eq < equity_cap) { eq[n_eq * 2] = (double)t; eq[n_eq * 2 + 1] = balance; }   // :280-283
            ++n_eq;
            if (balance > max_equity) max_equity = balance;             // :286-287
            const double dd = max_equity - balance;
            const double dd_pct = (dd / max_equity) * 100.0;
            if (dd > max_dd) { max_dd = dd; max_dd_pct = dd_pct; }      // :298-300
        }
    }
    if (in_pos && N > 0) close_pos((double)__ldg(row + (N - 1)), N - 1, 3);   // :303-307 "End of Test"
    if (lane == 0) {
        double* s = stats + (int64_t)sym * 16;
        s[0] = balance; s[1] = (double)n_trades; s[2] = (double)n_win; s[3] = (double)n_lose;
        s[4] = total_profit; s[5] = total_loss; s[6] = max_dd; s[7] = max_dd_pct; s[8] = (double)n_eq;
        s[9] = in_pos ? 1.0 : 0.0;
        for (int i = 10; i < 16; ++i) s[i] = 0.0;
    }
}

}  // namespace b200bt

using namespace b200bt;

extern "C" int b200bt_backtest_ref(const float* price, int64_t ld, int S, int64_t N, const b200bt_bt_params* params,
                                   double* stats, double* trades, int64_t trade_cap, double* equity,
                                   int64_t equity_cap, b200bt_stream_t stream) {
    B200BT_REQUIRE(price && params && stats && trades && equity, B200BT_EINVAL, "backtest_ref: null pointer");
    B200BT_REQUIRE(S > 0 && N > 0 && ld >= N && trade_cap > 0 && equity_cap > 0, B200BT_EINVAL, "backtest_ref: bad sizes");
    int rc = check_device();
    if (rc) return rc;
    backtest_ref_kernel<<<S, 32, 0, (cudaStream_t)stream>>>(price, ld, N, params, stats, trades, trade_cap, equity, equity_cap);
    B200BT_LAUNCH_CHECK("backtest_ref launch");
    return B200BT_OK;
}
