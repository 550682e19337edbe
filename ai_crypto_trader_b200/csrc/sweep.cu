// Family 2: the per-bar entry/exit/stop/PnL state machine over
// (GA-individual x symbol) lanes, with calculate_metrics and the strategy score
// reduced in-kernel (sm_100a).
//
// Reference semantics: services/strategy_evaluation.py:746-878 (_simulate_trades),
// :32-228 (calculate_metrics), :579-633 (_calculate_strategy_score).
//
// Mapping: ONE WARP PER LANE.  The 32 threads of a warp look at 32 consecutive
// bars of the lane's (price, rsi) streams at once; the machine's state
// (flat / long / short + entry price) is warp-uniform.  A ballot finds the
// first bar in the window at which the current state has an event (entry
// signal, or take-profit / stop-loss / RSI exit), the state is advanced, and
// the remaining bars of the window are re-tested under the new state.  Quiet
// stretches therefore cost two coalesced 128-byte loads and four compares per
// 32 bars, and the time axis -- serial in the reference -- is consumed 32 bars
// per step.
//
// Exactness (bit-exact entry/exit bars against the float64 reference):
//  * RSI tests compare the fp32 bank value with thresholds the host rounded so
//    that the fp32 compare decides like the float64 one.
//  * The price exits `(p-e)/e >= tp`, `<= -sl` are screened with fp32 bounds
//    that are provably conservative (relative margin 1e-6); a bar that is not
//    a definite exit is decided by the float64 expression itself.
//  * PnL, equity, drawdown, daily buckets and the score are float64, evaluated
//    from the queued events 32 at a time with warp scans.
#include <math.h>
#include "common.cuh"

namespace b200bt {

constexpr int SW_WARPS = 8;         // warps (lanes of the sweep) per CTA
constexpr int SW_V = 4;             // 32-bar windows loaded per group
constexpr float SW_MARGIN = 1e-6f;  // relative width of the fp32 screening band

struct LaneConst {
    float os_f, ob_f;
    // screening multipliers: candidate (c) and definite (d) bounds
    float hiL_c, hiL_d, loL_c, loL_d;  // long : TP above, SL below
    float hiS_c, hiS_d, loS_c, loS_d;  // short: SL above, TP below
    double tp, sl, size, fee1, fee2;
};

struct LaneAcc {
    double equity, peak, maxdd;
    double tot_profit, tot_loss, largest_p, largest_l;
    double day_sum, pivot, s1, s2;
    long long day_cur;
    long long sum_dur;
    unsigned long long hash;
    unsigned n_win, n_loss, n_days, n_events;
    int day_valid, pivot_set;
};

__device__ __forceinline__ void day_complete(LaneAcc& a, double x) {
    // shifted-data accumulation of the daily pnl sums (exactly 0 variance for equal days)
    if (!a.pivot_set) { a.pivot = x; a.pivot_set = 1; }
    double y = x - a.pivot;
    a.s1 += y;
    a.s2 += y * y;
    a.n_days += 1;
}

// Consume `cnt` queued events (lane j holds event j; even j = entry record,
// odd j = exit record of the same round trip).
__device__ __forceinline__ void process_batch(int cnt, unsigned w, float pf, const LaneConst& c,
                                              LaneAcc& a, const b200bt_sweep_config& cfg,
                                              uint32_t* ev_out, int64_t ev_cap) {
    const int lane = threadIdx.x & 31;
    const bool active = lane < cnt;
    const bool is_exit = active && (w & B200BT_EVENT_EXIT);
    const float p_prev = __shfl_up_sync(FULL, pf, 1);
    const unsigned w_prev = __shfl_up_sync(FULL, w, 1);
    const long long bar = (long long)(w & 0x3fffffffu);

    double pnl = 0.0;
    int dur = 0;
    if (active) {
        if (is_exit) {
            const double e = (double)p_prev, px = (double)pf;
            const double qty = __ddiv_rn(c.size, e);
            const double diff = (w_prev & B200BT_EVENT_SELL) ? __dsub_rn(e, px) : __dsub_rn(px, e);
            pnl = __dsub_rn(__dmul_rn(qty, diff), c.fee2);
            dur = (int)(bar - (long long)(w_prev & 0x3fffffffu));
        } else {
            pnl = -c.fee1;
        }
    }
    // wins / losses
    const bool win = active && pnl > 0.0, loss = active && pnl < 0.0;
    a.n_win += __popc(__ballot_sync(FULL, win));
    a.n_loss += __popc(__ballot_sync(FULL, loss));
    a.tot_profit += warp_sum_d(win ? pnl : 0.0);
    a.tot_loss += warp_sum_d(loss ? pnl : 0.0);
    a.largest_p = fmax(a.largest_p, warp_max_d(win ? pnl : 0.0));
    a.largest_l = fmin(a.largest_l, warp_min_d(loss ? pnl : 0.0));
    a.sum_dur += __reduce_add_sync(FULL, dur);

    // equity curve: inclusive scan of pnl, running peak, drawdown
    double cs = pnl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double up = shfl_up_d(cs, d);
        if (lane >= d) cs += up;
    }
    const double eq = a.equity + cs;
    double pk = active ? eq : -INFINITY;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double up = shfl_up_d(pk, d);
        if (lane >= d) pk = fmax(pk, up);
    }
    pk = fmax(pk, a.peak);
    double dd = 0.0;
    if (active && eq < pk) dd = __ddiv_rn(__dsub_rn(pk, eq), pk);
    a.maxdd = fmax(a.maxdd, warp_max_d(dd));
    a.equity = shfl_d(eq, cnt - 1);
    a.peak = shfl_d(pk, cnt - 1);

    // daily buckets: segmented sum by calendar day over the batch, merged with the carry day
    long long day = active ? (cfg.minute0 + bar * (long long)cfg.bar_minutes) / 1440 : 0;
    const long long day_prev = __shfl_up_sync(FULL, day, 1);
    const long long day_next = __shfl_down_sync(FULL, day, 1);
    const bool head = active && (lane == 0 || day != day_prev);
    const bool tail = active && (lane == cnt - 1 || day != day_next);
    double seg = pnl;
    bool flag = head;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double up = shfl_up_d(seg, d);
        int fup = __shfl_up_sync(FULL, (int)flag, d);
        if (lane >= d && !flag) { seg += up; flag = fup; }
    }
    const long long first_day = __shfl_sync(FULL, day, 0);
    const bool merge_carry = a.day_valid && (first_day == a.day_cur);
    // a tail in the first segment (its day == first_day) absorbs the carry
    if (tail && merge_carry && day == first_day) seg += a.day_sum;
    const bool last_seg_tail = tail && (lane == cnt - 1);
    // completed days: every tail except the batch's last one (which stays open) ...
    double x = (tail && !last_seg_tail) ? seg : 0.0;
    unsigned done = __ballot_sync(FULL, tail && !last_seg_tail);
    // ... plus the carried day when the batch starts on a later day
    const bool carry_done = a.day_valid && !merge_carry;
    // fold completed days in time order (carry first, then lanes ascending): cheap, rare
    if (carry_done) day_complete(a, a.day_sum);
    while (done) {
        int j = __ffs(done) - 1;
        done &= done - 1;
        day_complete(a, shfl_d(x, j));
    }
    a.day_sum = shfl_d(seg, cnt - 1);
    a.day_cur = __shfl_sync(FULL, day, cnt - 1);
    a.day_valid = 1;

    // trade hash + optional event buffer
    unsigned long long h = active ? mix64(((unsigned long long)(a.n_events + lane) << 32) | w) : 0ull;
    unsigned hlo = __reduce_xor_sync(FULL, (unsigned)h);
    unsigned hhi = __reduce_xor_sync(FULL, (unsigned)(h >> 32));
    a.hash ^= ((unsigned long long)hhi << 32) | hlo;
    if (ev_out && active) {
        long long idx = (long long)a.n_events + lane;
        if (idx < ev_cap) ev_out[idx] = w;
    }
    a.n_events += cnt;
}

__global__ void __launch_bounds__(SW_WARPS * 32)
sweep_kernel(const float* __restrict__ price, int64_t ld_price,
             const float* __restrict__ rsi, int64_t ld_rsi, int P, int S, int64_t N,
             const b200bt_individual* __restrict__ indiv, const int32_t* __restrict__ order, int pop,
             const b200bt_sweep_config cfg,  // NOT __grid_constant__: nvcc 12.9 then forwards the param load over a.equity's loop-carried value

             b200bt_lane_stats* __restrict__ stats, uint32_t* __restrict__ events, int64_t ev_cap) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * SW_WARPS + (threadIdx.x >> 5);
    if (gw >= (int64_t)pop * S) return;
    const int sym = (int)(gw / pop);
    const int k = (int)(gw % pop);
    const int ind = order ? order[k] : k;
    const b200bt_individual iv = indiv[ind];

    LaneConst c;
    c.os_f = iv.rsi_lo;
    c.ob_f = iv.rsi_hi;
    c.tp = iv.take_profit;
    c.sl = iv.stop_loss;
    c.size = iv.position_size;
    c.fee1 = __dmul_rn(c.size, 0.001);
    c.fee2 = __dmul_rn(c.size, 0.002);
    {
        const double m = (double)SW_MARGIN;
        c.hiL_c = (float)((1.0 + c.tp) * (1.0 - m));
        c.hiL_d = (float)((1.0 + c.tp) * (1.0 + m));
        c.loL_c = (float)((1.0 - c.sl) * (1.0 + m));
        c.loL_d = (float)((1.0 - c.sl) * (1.0 - m));
        c.loS_c = (float)((1.0 - c.tp) * (1.0 + m));
        c.loS_d = (float)((1.0 - c.tp) * (1.0 - m));
        c.hiS_c = (float)((1.0 + c.sl) * (1.0 - m));
        c.hiS_d = (float)((1.0 + c.sl) * (1.0 + m));
    }
    LaneAcc a;
    a.equity = cfg.initial_capital;
    a.peak = cfg.initial_capital;
    a.maxdd = 0.0;
    a.tot_profit = a.tot_loss = a.largest_p = a.largest_l = 0.0;
    a.day_sum = a.pivot = a.s1 = a.s2 = 0.0;
    a.day_cur = 0;
    a.sum_dur = 0;
    a.hash = 0ull;
    a.n_win = a.n_loss = a.n_days = a.n_events = 0;
    a.day_valid = a.pivot_set = 0;

    const float* __restrict__ pr = price + (int64_t)sym * ld_price;
    const float* __restrict__ rr = rsi + ((int64_t)sym * P + iv.rsi_row) * ld_rsi;
    uint32_t* ev_out = events ? events + ((int64_t)ind * S + sym) * ev_cap : nullptr;

    // machine state (warp-uniform)
    int pos = 0;
    float e = 0.f;
    float rlo = c.os_f, rhi = c.ob_f, plo = -INFINITY, phi = INFINITY, plo_d = -INFINITY, phi_d = INFINITY;
    // event queue: lane j holds event j of the current batch
    unsigned q_w = 0;
    float q_p = 0.f;
    int qn = 0;

    const float qnan = __int_as_float(0x7fc00000);
    const int64_t ngroups = (N + 32 * SW_V - 1) / (32 * SW_V);
    float pn[SW_V], rn[SW_V];
#pragma unroll
    for (int v = 0; v < SW_V; ++v) {
        int64_t t = (int64_t)v * 32 + lane;
        pn[v] = t < N ? __ldg(pr + t) : qnan;
        rn[v] = t < N ? __ldg(rr + t) : qnan;
    }
    for (int64_t g = 0; g < ngroups; ++g) {
        float pv[SW_V], rv[SW_V];
#pragma unroll
        for (int v = 0; v < SW_V; ++v) { pv[v] = pn[v]; rv[v] = rn[v]; }
        if (g + 1 < ngroups) {
#pragma unroll
            for (int v = 0; v < SW_V; ++v) {
                int64_t t = ((g + 1) * SW_V + v) * 32 + lane;
                pn[v] = t < N ? __ldg(pr + t) : qnan;
                rn[v] = t < N ? __ldg(rr + t) : qnan;
            }
        }
#pragma unroll
        for (int v = 0; v < SW_V; ++v) {
            const float p = pv[v], r = rv[v];
            const int64_t t0 = (g * SW_V + v) * 32;
            int start = 0;
            while (true) {
                const bool ev = (lane >= start) && (r < rlo || r > rhi || p <= plo || p >= phi);
                const unsigned m = __ballot_sync(FULL, ev);
                if (m == 0) break;
                const int kk = __ffs(m) - 1;
                const float pk = __shfl_sync(FULL, p, kk);
                const float rk = __shfl_sync(FULL, r, kk);
                start = kk + 1;
                unsigned word;
                if (pos == 0) {
                    // entry (strategy_evaluation.py:784-813): long has priority over short
                    e = pk;
                    if (rk < c.os_f) {
                        pos = 1;
                        rlo = -INFINITY; rhi = c.ob_f;
                        phi = e * c.hiL_c; phi_d = e * c.hiL_d;
                        plo = e * c.loL_c; plo_d = e * c.loL_d;
                        word = (unsigned)(t0 + kk);
                    } else {
                        pos = -1;
                        rlo = c.os_f; rhi = INFINITY;
                        phi = e * c.hiS_c; phi_d = e * c.hiS_d;
                        plo = e * c.loS_c; plo_d = e * c.loS_d;
                        word = (unsigned)(t0 + kk) | B200BT_EVENT_SELL;
                    }
                } else {
                    // exit candidate (:815-847)
                    bool definite = (rk < rlo) || (rk > rhi) || (pk >= phi_d) || (pk <= plo_d);
                    if (!definite) {
                        const double ed = (double)e, pd = (double)pk;
                        const double q = (pos > 0) ? __ddiv_rn(__dsub_rn(pd, ed), ed)
                                                   : __ddiv_rn(__dsub_rn(ed, pd), ed);
                        if (!(q >= c.tp || q <= -c.sl)) continue;  // inside the screening band, no exit
                    }
                    word = (unsigned)(t0 + kk) | B200BT_EVENT_EXIT | (pos > 0 ? B200BT_EVENT_SELL : 0u);
                    pos = 0;
                    rlo = c.os_f; rhi = c.ob_f;
                    plo = plo_d = -INFINITY; phi = phi_d = INFINITY;
                }
                if (lane == qn) { q_w = word; q_p = pk; }
                if (++qn == 32) {
                    process_batch(32, q_w, q_p, c, a, cfg, ev_out, ev_cap);
                    qn = 0;
                }
            }
        }
    }
    if (pos != 0) {
        // force-close at the last bar (:849-876)
        const float pl = __ldg(pr + (N - 1));
        const unsigned word = (unsigned)(N - 1) | B200BT_EVENT_EXIT | (pos > 0 ? B200BT_EVENT_SELL : 0u);
        if (lane == qn) { q_w = word; q_p = pl; }
        ++qn;
    }
    if (qn > 0) process_batch(qn, q_w, q_p, c, a, cfg, ev_out, ev_cap);

    if (lane == 0) {
        if (a.day_valid) day_complete(a, a.day_sum);
        b200bt_lane_stats o;
        const double n_rec = (double)a.n_events;
        o.n_records = n_rec;
        o.n_wins = (double)a.n_win;
        o.n_losses = (double)a.n_loss;
        o.total_profit = a.tot_profit;
        o.total_loss = a.tot_loss;
        o.net_profit = a.tot_profit + a.tot_loss;
        o.max_drawdown = a.maxdd;
        o.n_days = (double)a.n_days;
        o.largest_profit = a.largest_p;
        o.largest_loss = a.largest_l;
        o.sum_duration_bars = (double)a.sum_dur;
        o.trade_hash = a.hash;
        double sharpe = 0.0, win_rate = 0.0, pf = 0.0;
        if (a.n_events >= 2) {
            win_rate = (double)a.n_win / n_rec;
            pf = (a.tot_loss != 0.0) ? fabs(a.tot_profit / a.tot_loss) : INFINITY;
            if (a.n_days > 1) {
                const double n = (double)a.n_days;
                const double mean_y = a.s1 / n;
                double var = a.s2 / n - mean_y * mean_y;
                if (var < 0.0) var = 0.0;
                const double sd = sqrt(var);
                const double mean = a.pivot + mean_y;
                sharpe = sd > 0.0 ? (mean / sd) * sqrt(252.0) : 0.0;
            }
        }
        o.sharpe_ratio = sharpe;
        o.win_rate = win_rate;
        o.profit_factor = pf;
        double primary;
        switch (cfg.primary) {
            case B200BT_PRIMARY_RETURN_PCT: primary = (o.net_profit / cfg.initial_capital) * 100.0; break;
            case B200BT_PRIMARY_PROFIT_FACTOR: primary = pf; break;
            case B200BT_PRIMARY_WIN_RATE: primary = win_rate; break;
            case B200BT_PRIMARY_NET_PROFIT: primary = o.net_profit; break;
            default: primary = sharpe; break;
        }
        double score = primary;
        if (cfg.secondary_mask & B200BT_SEC_MAX_DRAWDOWN) score *= (1.0 - a.maxdd);
        if (cfg.secondary_mask & B200BT_SEC_WIN_RATE) score *= (1.0 + win_rate);
        if (cfg.secondary_mask & B200BT_SEC_PROFIT_FACTOR) score *= (pf / 2.0);
        o.score = score;
        stats[(int64_t)ind * S + sym] = o;
    }
}

__global__ void fitness_reduce_kernel(const b200bt_lane_stats* __restrict__ stats, int pop, int S,
                                      double* __restrict__ fitness) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pop) return;
    double acc = 0.0;
    for (int s = 0; s < S; ++s) acc += stats[(int64_t)i * S + s].score;
    fitness[i] = acc / (double)S;
}

}  // namespace b200bt

using namespace b200bt;

extern "C" int b200bt_sweep(const float* price, int64_t ld_price, const float* rsi, int64_t ld_rsi, int P,
                            int S, int64_t N, const b200bt_individual* indiv, const int32_t* order, int pop,
                            const b200bt_sweep_config* cfg_host, b200bt_lane_stats* stats,
                            uint32_t* events, int64_t event_cap, b200bt_stream_t stream) {
    B200BT_REQUIRE(price && rsi && indiv && cfg_host && stats, B200BT_EINVAL, "sweep: null pointer");
    B200BT_REQUIRE(S > 0 && N > 0 && P > 0 && pop > 0, B200BT_EINVAL, "sweep: bad sizes");
    B200BT_REQUIRE(ld_price >= N && ld_rsi >= N, B200BT_EINVAL, "sweep: row stride shorter than N");
    B200BT_REQUIRE(N < (1ll << 30), B200BT_ELIMIT, "sweep: N must be < 2^30 bars");
    B200BT_REQUIRE(cfg_host->bar_minutes > 0, B200BT_EINVAL, "sweep: bar_minutes must be > 0");
    B200BT_REQUIRE(events == nullptr || event_cap > 0, B200BT_EINVAL, "sweep: event buffer without capacity");
    int rc = check_device();
    if (rc) return rc;
    const int64_t lanes = (int64_t)pop * S;
    const int64_t blocks = (lanes + SW_WARPS - 1) / SW_WARPS;
    B200BT_REQUIRE(blocks < (1ll << 31), B200BT_ELIMIT, "sweep: too many lanes");
    sweep_kernel<<<(unsigned)blocks, SW_WARPS * 32, 0, (cudaStream_t)stream>>>(
        price, ld_price, rsi, ld_rsi, P, S, N, indiv, order, pop, *cfg_host, stats, events, event_cap);
    B200BT_LAUNCH_CHECK("sweep launch");
    return B200BT_OK;
}

extern "C" int b200bt_fitness_reduce(const b200bt_lane_stats* stats, int pop, int S, double* fitness,
                                     b200bt_stream_t stream) {
    B200BT_REQUIRE(stats && fitness && pop > 0 && S > 0, B200BT_EINVAL, "fitness_reduce: bad argument");
    int rc = check_device();
    if (rc) return rc;
    fitness_reduce_kernel<<<(pop + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, pop, S, fitness);
    B200BT_LAUNCH_CHECK("fitness_reduce launch");
    return B200BT_OK;
}
