/*
 * sim_oracle.c -- plain-C float64 restatement of the reference's parameterised
 * backtest rule, its trade metrics and its strategy score, for one
 * (individual, symbol) lane.  TEST INFRASTRUCTURE (see oracle/__init__.py):
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline only.
 *
 * Follows services/strategy_evaluation.py of zd87pl/ai-crypto-trader:
 *   state machine  _simulate_trades            :746-878
 *   metrics        calculate_metrics           :32-228
 *   score          _calculate_strategy_score   :579-633
 * The same arithmetic as oracle/simulate_ref.py, which is validated against the
 * reference's own functions; this C form exists so that full-length (1M bar)
 * lanes can be checked in milliseconds.
 *
 * Inputs are the fp32 series the GPU path consumes, widened to double exactly
 * as Python widens them (float(np.float32)).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o libsim_oracle.so sim_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double rsi_oversold, rsi_overbought; /* thresholds as float64 (ints in the GA) */
    double take_profit, stop_loss;       /* already /100 */
    double position_size;
} oracle_params;

typedef struct {
    double initial_capital;
    int64_t minute0;
    int32_t bar_minutes;
    int32_t primary;        /* 0 sharpe, 1 return_pct, 2 profit_factor, 3 win_rate, 4 net_profit */
    int32_t secondary_mask; /* 1 max_drawdown, 2 win_rate, 4 profit_factor */
    int32_t reserved;
    int32_t gap_bar;        /* > 0: bars >= gap_bar carry timestamps gap_minutes later (glued training folds) */
    int32_t gap_minutes;
} oracle_config;

typedef struct {
    double n_records, n_wins, n_losses, total_profit, total_loss, net_profit, max_drawdown,
        sharpe_ratio, n_days, largest_profit, largest_loss, sum_duration_bars, score, win_rate,
        profit_factor;
    /* calculate_advanced_metrics :250-263, :307-312 */
    double sortino_ratio, n_negative_days, downside_deviation, mean_daily_pnl;
    uint64_t trade_hash;
} oracle_stats;

#define EV_EXIT 0x40000000u
#define EV_SELL 0x80000000u

/* same mixer as csrc/common.cuh event_hash */
static uint64_t event_hash(uint32_t index, uint32_t word) {
    uint32_t a = (index * 0x9E3779B1u) ^ word;
    a *= 0x85EBCA77u;
    a ^= a >> 15;
    uint32_t b = (word * 0xC2B2AE3Du) ^ ((index << 13) | (index >> 19));
    b *= 0x27D4EB2Fu;
    b ^= b >> 16;
    return ((uint64_t)b << 32) | a;
}

typedef struct {
    const oracle_config* cfg;
    double equity, peak, maxdd;
    int have_dd;
    double tot_p, tot_l, largest_p, largest_l;
    uint64_t n_rec, n_win, n_loss, hash;
    int64_t cur_day;
    int day_open;
    double day_sum;
    /* daily sums kept in a growable array: np.std is two-pass */
    double* days;
    size_t n_days, cap_days;
    int64_t sum_dur;
    uint32_t* ev;
    double* ev_pnl;
    int64_t ev_cap;
} acc_t;

static void push_day(acc_t* a, double v) {
    if (a->n_days == a->cap_days) {
        a->cap_days = a->cap_days ? a->cap_days * 2 : 1024;
        a->days = (double*)realloc(a->days, a->cap_days * sizeof(double));
    }
    a->days[a->n_days++] = v;
}

/* one trade record: calculate_metrics' per-record loop (:133-160) */
static void record(acc_t* a, int64_t bar, uint32_t flags, double pnl) {
    uint32_t w = (uint32_t)bar | flags;
    if (a->ev && (int64_t)a->n_rec < a->ev_cap) {
        a->ev[a->n_rec] = w;
        if (a->ev_pnl) a->ev_pnl[a->n_rec] = pnl;
    }
    a->hash ^= event_hash((uint32_t)a->n_rec, w);
    a->n_rec++;
    if (pnl > 0) {
        a->n_win++;
        a->tot_p += pnl;
        if (pnl > a->largest_p) a->largest_p = pnl;
    } else if (pnl < 0) {
        a->n_loss++;
        a->tot_l += pnl;
        if (pnl < a->largest_l) a->largest_l = pnl;
    }
    a->equity = a->equity + pnl;
    if (a->equity > a->peak) {
        a->peak = a->equity;
    } else {
        double dd = (a->peak - a->equity) / a->peak;
        if (!a->have_dd || dd > a->maxdd) a->maxdd = dd;
        a->have_dd = 1;
    }
    int64_t day = (a->cfg->minute0 + bar * (int64_t)a->cfg->bar_minutes +
                   ((a->cfg->gap_bar > 0 && bar >= a->cfg->gap_bar) ? a->cfg->gap_minutes : 0)) / 1440;
    if (a->day_open && day == a->cur_day) {
        a->day_sum += pnl;
    } else {
        if (a->day_open) push_day(a, a->day_sum);
        a->cur_day = day;
        a->day_sum = pnl;
        a->day_open = 1;
    }
}

/* Simulate one lane.  `events`/`event_pnl` (optional) receive the first event_cap
 * event words (bar | EV_EXIT | EV_SELL) and record pnls. Returns 0. */
int oracle_lane(const float* price, const float* rsi, int64_t n, const oracle_params* p,
                const oracle_config* cfg, oracle_stats* out, uint32_t* events, double* event_pnl,
                int64_t event_cap) {
    acc_t a;
    memset(&a, 0, sizeof(a));
    a.cfg = cfg;
    a.equity = cfg->initial_capital;
    a.peak = cfg->initial_capital;
    a.ev = events;
    a.ev_pnl = event_pnl;
    a.ev_cap = event_cap;
    const double size = p->position_size;
    const double fee1 = size * 0.001, fee2 = size * 0.002;
    int side = 0;
    double entry = 0.0;
    int64_t entry_bar = 0;
    for (int64_t t = 0; t < n; ++t) {
        const double px = (double)price[t], r = (double)rsi[t];
        if (side == 0) {
            if (r < p->rsi_oversold) {
                side = 1; entry = px; entry_bar = t;
                record(&a, t, 0u, -fee1);
            } else if (r > p->rsi_overbought) {
                side = -1; entry = px; entry_bar = t;
                record(&a, t, EV_SELL, -fee1);
            }
            continue;
        }
        const double gain = (side > 0 ? (px - entry) : (entry - px)) / entry;
        const int reversal = side > 0 ? (r > p->rsi_overbought) : (r < p->rsi_oversold);
        if (gain >= p->take_profit || gain <= -p->stop_loss || reversal) {
            const double qty = size / entry;
            const double pnl = qty * (side > 0 ? (px - entry) : (entry - px)) - fee2;
            record(&a, t, EV_EXIT | (side > 0 ? EV_SELL : 0u), pnl);
            a.sum_dur += t - entry_bar;
            side = 0;
        }
    }
    if (side != 0 && n > 0) {
        const double px = (double)price[n - 1];
        const double qty = size / entry;
        const double pnl = qty * (side > 0 ? (px - entry) : (entry - px)) - fee2;
        record(&a, n - 1, EV_EXIT | (side > 0 ? EV_SELL : 0u), pnl);
        a.sum_dur += (n - 1) - entry_bar;
    }
    if (a.day_open) push_day(&a, a.day_sum);

    memset(out, 0, sizeof(*out));
    out->n_records = (double)a.n_rec;
    out->n_wins = (double)a.n_win;
    out->n_losses = (double)a.n_loss;
    out->total_profit = a.tot_p;
    out->total_loss = a.tot_l;
    out->net_profit = a.tot_p + a.tot_l;
    out->max_drawdown = a.have_dd ? a.maxdd : 0.0;
    out->n_days = (double)a.n_days;
    out->largest_profit = a.largest_p;
    out->largest_loss = a.largest_l;
    out->sum_duration_bars = (double)a.sum_dur;
    out->trade_hash = a.hash;
    double sharpe = 0.0, win_rate = 0.0, pf = 0.0;
    if (a.n_rec >= 2) {
        win_rate = (double)a.n_win / (double)a.n_rec;
        pf = (a.tot_l != 0.0) ? fabs(a.tot_p / a.tot_l) : INFINITY;
        if (a.n_days > 1) {
            /* np.mean / np.std (ddof 0), two-pass */
            double s = 0.0;
            for (size_t i = 0; i < a.n_days; ++i) s += a.days[i];
            const double mean = s / (double)a.n_days;
            double v = 0.0;
            for (size_t i = 0; i < a.n_days; ++i) v += (a.days[i] - mean) * (a.days[i] - mean);
            const double sd = sqrt(v / (double)a.n_days);
            sharpe = sd > 0.0 ? (mean / sd) * sqrt(252.0) : 0.0;
        }
    }
    out->sharpe_ratio = sharpe;
    out->win_rate = win_rate;
    out->profit_factor = pf;
    if (a.n_rec >= 2 && a.n_days > 0) {
        /* np.mean(daily); np.std of the negative ones (ddof 0, two-pass) */
        double s = 0.0, sn = 0.0;
        size_t nn = 0;
        for (size_t i = 0; i < a.n_days; ++i) {
            s += a.days[i];
            if (a.days[i] < 0) { sn += a.days[i]; ++nn; }
        }
        const double mean = s / (double)a.n_days;
        double sd = 0.0;
        if (nn) {
            const double mn = sn / (double)nn;
            double v = 0.0;
            for (size_t i = 0; i < a.n_days; ++i)
                if (a.days[i] < 0) v += (a.days[i] - mn) * (a.days[i] - mn);
            sd = sqrt(v / (double)nn);
        }
        out->mean_daily_pnl = mean;
        out->n_negative_days = (double)nn;
        out->downside_deviation = sd;
        out->sortino_ratio = sd > 0.0 ? (mean / sd) * sqrt(252.0) : INFINITY;
    }
    /* _calculate_strategy_score :579-633 on the metrics dict (codes as in include/b200bt.h) */
    const double ret_pct = (out->net_profit / cfg->initial_capital) * 100.0;
    const double avg_p = out->n_wins > 0 ? out->total_profit / out->n_wins : 0.0;
    const double avg_l = out->n_losses > 0 ? out->total_loss / out->n_losses : 0.0;
    const double expectancy = a.n_rec > 0 ? win_rate * avg_p - (1.0 - win_rate) * fabs(avg_l) : 0.0;   /* :302-310 */
    double primary;
    switch (cfg->primary) {
        case 1: primary = ret_pct; break;
        case 2: primary = pf; break;
        case 3: primary = win_rate; break;
        case 4: primary = out->net_profit; break;
        case 5: primary = out->n_records; break;
        case 6: primary = out->max_drawdown; break;
        case 7: primary = out->total_profit; break;
        case 8: primary = out->total_loss; break;
        case 9: primary = out->largest_profit; break;
        case 10: primary = out->largest_loss; break;
        case 11: primary = avg_p; break;
        case 12: primary = avg_l; break;
        case 13: primary = out->sortino_ratio; break;
        case 14: primary = expectancy; break;
        case 15: primary = out->max_drawdown > 0.0 ? (ret_pct / 100.0) / out->max_drawdown : INFINITY; break;
        case 16: primary = out->mean_daily_pnl; break;
        case 17: primary = out->max_drawdown > 0.0 ? out->net_profit / (out->max_drawdown * 10000.0) : INFINITY; break;
        case 99: primary = 0.0; break;
        default: primary = sharpe; break;
    }
    double score = primary;
    if (cfg->secondary_mask & 1) score *= (1.0 - out->max_drawdown);
    if (cfg->secondary_mask & 2) score *= (1.0 + win_rate);
    if (cfg->secondary_mask & 4) score *= (pf / 2.0);
    if (cfg->secondary_mask & 8) score *= (1.0 + fmin(expectancy / 100.0, 1.0));
    out->score = score;
    free(a.days);
    return 0;
}

/* Many lanes of one symbol stream pair, used by the CPU baseline timing. */
int oracle_lanes(const float* price, const float* rsi, int64_t n, const oracle_params* p, int n_lanes,
                 const oracle_config* cfg, oracle_stats* out) {
    for (int i = 0; i < n_lanes; ++i) oracle_lane(price, rsi, n, &p[i], cfg, &out[i], 0, 0, 0);
    return 0;
}
