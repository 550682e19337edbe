#!/usr/bin/env python
"""bench.py -- GA population-fitness sweep throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one generation's fitness evaluation: the state-machine sweep over
every (individual x symbol) lane of the population plus the fitness reduction
(and, for N > 1, one NCCL all-gather of the fitness vector).  Workload at N=1 is
BASELINE.json configs[1]: population 1024, 10 symbols, 1M synthetic 1-minute
bars.  For N > 1 the population is sharded by individual, per-GPU work fixed
(weak scaling): global population = 1024*N, market data replicated.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "bar_strategy_evals_per_sec"
UNIT = "evals/s"
POP_PER_GPU = 1024
N_SYMBOLS = 10
N_BARS = 1_000_000
BYTES_PER_EVAL = 8  # price 4 B + RSI 4 B per (individual, symbol, bar), reference RSI rule


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows[-3:]]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port; the Python reference
# itself cannot travel to the GPU box) on all host cores
# --------------------------------------------------------------------------
_REF_STREAMS = []   # market-data dict lists, built in the parent before the fork (shared copy-on-write)


def _ref_worker(job):
    stream, params, goals = job
    from oracle import simulate_ref
    recs = simulate_ref.simulate_trades(params, _REF_STREAMS[stream])
    m = simulate_ref.calculate_metrics(recs)
    return float(simulate_ref.strategy_score(m, goals)), len(recs)


def reference_sample(n_bars: int, lanes: int):
    """Bounded sample of the workload: `lanes` individuals of the seed-42 population on
    n_bars bars of symbol 0, their rsi_period folded onto two streams (7, 14) so the
    prebuilt market-data dict lists stay small.  Returns the job list."""
    from ai_crypto_trader_b200 import synth
    from ai_crypto_trader_b200.sweep import DEFAULT_GOALS
    from oracle import indicators_ref, simulate_ref
    pop = synth.random_population(lanes, seed=42)
    close = synth.synth_symbol(0, n_bars)["close"]
    periods = [7, 14]
    bank = indicators_ref.rsi_bank(close, periods)
    _REF_STREAMS.clear()
    for row in bank:
        _REF_STREAMS.append(simulate_ref.market_points(close, row, "SYN000USDT", synth.EPOCH_2024_MINUTES))
    return [(i % 2, dict(p, rsi_period=periods[i % 2]), DEFAULT_GOALS) for i, p in enumerate(pop)]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    use = max(1, min(cores, 64))
    n_bars = 250_000
    lanes = use * 4
    jobs = reference_sample(n_bars, lanes)
    evals_per_step = n_bars * lanes
    times = []
    with mp.get_context("fork").Pool(use) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = evals_per_step * len(times) / total
    sample = f"{lanes} lanes x {n_bars} bars of configs[1] per step (pure-Python float64 restatement of _simulate_trades+calculate_metrics+score; market-data dicts prebuilt outside the timed region)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"GA fitness sweep pop={POP_PER_GPU}x{args.gpus} symbols={N_SYMBOLS} bars={N_BARS} (bounded sample)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": use, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_inline(budget_s: float = 12.0):
    """Rank-0, N=1 only: the oracle timed on the host cores on a bounded sample."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    use = max(1, min(cores, 64))
    n_bars = 200_000
    lanes = use
    jobs = reference_sample(n_bars, lanes)
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(use) as pool:
        pool.map(_ref_worker, jobs, chunksize=1)
    dt = time.perf_counter() - t0
    py = {"value": n_bars * lanes / dt, "unit": UNIT, "cores": use, "kind": "port",
          "sample": f"{lanes} lanes x {n_bars} bars, pure-Python float64 restatement of the reference loop (its own speed class), {use} processes"}
    # the same algorithm as compiled C (a much stronger CPU baseline than the reference's Python)
    from oracle import sim_oracle
    import numpy as np
    from ai_crypto_trader_b200 import synth
    from oracle import indicators_ref
    close = synth.synth_symbol(0, N_BARS)["close"]
    pop = synth.random_population(16, seed=42)
    bank = indicators_ref.rsi_bank(close, [14])
    cfg = sim_oracle.config_of(synth.EPOCH_2024_MINUTES, 1)
    t0 = time.perf_counter()
    sim_oracle.lanes(close, bank[0], pop, cfg)
    dtc = time.perf_counter() - t0
    py["c_port_single_core"] = {"value": N_BARS * len(pop) / dtc, "unit": UNIT, "cores": 1,
                                "sample": f"{len(pop)} lanes x {N_BARS} bars, oracle/sim_oracle.c"}
    # BASELINE configs[0] (the reference's own CPU-runnable case): StrategyTester.backtest_strategy on 10 000 bars of one
    # symbol, restated in Python with the LLM stubbed (oracle/tester_ref.py), single process like the reference
    from oracle import tester_ref
    df = configs0_frame()
    t0 = time.perf_counter()
    st = tester_ref.backtest(df)
    py["configs0_backtest"] = {"cpu_ms": (time.perf_counter() - t0) * 1e3, "bars": len(df), "trades": int(st["total_trades"]),
                               "sample": "1 strategy x 1 symbol x 10 000 bars, oracle/tester_ref.py, 1 core"}
    return py


def configs0_frame():
    """10 000 synthetic 1-minute bars ending in a sell-off (the constant technical signal then trades, as in the parity
    fixtures tests/golden/bt_reference.*)."""
    import numpy as np
    import pandas as pd
    from ai_crypto_trader_b200 import synth
    n = 10_000
    d = synth.synth_symbol(1, n)
    f = np.ones(n)
    f[-60:] = np.linspace(1.0, 0.90, 60)
    cols = {k: d[k].astype(np.float64) * (f if k in ("open", "high", "low", "close") else 1.0) for k in synth.FIELDS}
    return pd.DataFrame({k: v.astype(np.float32).astype(np.float64) for k, v in cols.items()},
                        index=pd.date_range("2024-01-01", periods=n, freq="min"))


# --------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pop-per-gpu", type=int, default=POP_PER_GPU)
    ap.add_argument("--symbols", type=int, default=N_SYMBOLS)
    ap.add_argument("--bars", type=int, default=N_BARS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="auto", choices=["auto", "fused", "chunked", "tiled"], help="sweep kernel path (auto = product default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    assert args.warmup >= 3 or os.environ.get("B200BT_ALLOW_SHORT_WARMUP"), "timing rules: warm-up >= 3"
    # CPU baseline first (rank 0, N=1 only): fork-based worker pool before CUDA is initialised
    cpu_baseline = None
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_inline()

    import numpy as np
    import torch
    import torch.distributed as dist
    from ai_crypto_trader_b200 import _lib, synth
    from ai_crypto_trader_b200.sweep import MarketData, PopulationSweep, decode_population, evaluation_order

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run for N>1)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pop_local, S, N = args.pop_per_gpu, args.symbols, args.bars
    pop_global = pop_local * world
    # synthetic market, identical on every rank (replicated), pinned on the host for the e2e leg
    ohlcv_host = torch.from_numpy(synth.synth_ohlcv(S, N)).pin_memory()
    market = MarketData(ohlcv_host)
    sweep = PopulationSweep(market, mode=args.mode)
    population = synth.random_population(pop_global, seed=42)
    my_pop = population[rank * pop_local:(rank + 1) * pop_local]
    packed = decode_population(my_pop, sweep.period_row)
    order = evaluation_order(my_pop)
    indiv_dev = torch.from_numpy(packed.view(np.uint8)).to(dev)
    order_dev = torch.from_numpy(order).to(dev)
    fit_local = torch.empty(pop_local, dtype=torch.float64, device=dev)
    fit_global = torch.empty(pop_global, dtype=torch.float64, device=dev)
    plan = sweep.plan(my_pop)          # the path sweep.evaluate() takes: None = fused kernel, else tile / chunk plans
    path = "fused" if plan is None else ("tiled" if type(plan[0]).__name__ == "TilePlan" else "chunked")
    path_kernels = {"fused": "sweep_kernel", "chunked": "chunk_scan_kernel + chunk_sums/partial/lane_combine",
                    "tiled": "lane_scan_kernel + chunk_sums/partial/lane_combine"}[path]
    # DRAM bytes (read + write) of the dominant kernel per launch on THIS workload, from the committed
    # `ncu --set full` captures (profiles/r1_tiled_final_ncu.txt, r1_chunk_scan_ncu.txt, r1_sweep_v2_ncu.txt);
    # null for any other workload size
    ncu_traffic = {"tiled": 5.222e9 + 1.006e9, "chunked": 45.6e9, "fused": 3.3e9}[path] \
        if (pop_local, S, N) == (POP_PER_GPU, N_SYMBOLS, N_BARS) else None

    def step():
        sweep.evaluate_device(indiv_dev, order_dev, pop_local, fit_local, plan=plan)
        if world > 1:
            dist.all_gather_into_tensor(fit_global, fit_local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg ------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    t_wall0 = time.time()
    ev0.record()
    for i in range(args.steps):
        # the sweep (scan [+ verify/repair + metrics] kernels, then the fitness reduction) is bracketed by its own
        # events inside the timed region; the all-gather follows
        k_ev[i][0].record()
        sweep.evaluate_device(indiv_dev, order_dev, pop_local, fit_local, plan=plan)
        k_ev[i][1].record()
        if world > 1:
            dist.all_gather_into_tensor(fit_global, fit_local)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = _lib.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    ms_kernel = statistics.mean(a.elapsed_time(b) for a, b in k_ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    evals_per_step_global = pop_global * S * N
    value = evals_per_step_global * args.steps / (ms_total * 1e-3)

    # ---- end-to-end leg: host OHLCV + host population in, host fitness out ----
    def e2e_step():
        mk = MarketData(ohlcv_host)                      # H2D of the pinned OHLCV (5 fields)
        sw = PopulationSweep(mk, mode=args.mode)         # RSI bank on device
        f = sw.evaluate(my_pop)                          # H2D params, sweep, reduce, D2H fitness
        if world > 1:
            g = torch.from_numpy(f).to(dev)
            out = torch.empty(pop_global, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, g)
            f = out.cpu().numpy()
        return f, sw.h2d_bytes + mk.h2d_bytes, sw.d2h_bytes    # (the sweep reads close prices: the other OHLCV fields stay on the host)

    e2e_steps = max(3, min(args.steps, 5))
    f_e2e, h2d, d2h = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        f_e2e, h2d, d2h = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = evals_per_step_global * e2e_steps / float(te.item())

    # sanity: the e2e result equals the device-resident result
    assert np.allclose(f_e2e[rank * pop_local:(rank + 1) * pop_local] if world > 1 else f_e2e,
                       fit_local.cpu().numpy(), rtol=1e-12, atol=0, equal_nan=True)

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        lanes_evals = pop_local * S * N
        achieved = lanes_evals * BYTES_PER_EVAL / (ms_kernel * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 streams / f64 trade arithmetic", "data": "synthetic",
            "config": {"workload": f"GA fitness sweep (BASELINE configs[1] per GPU): population {pop_local}/GPU x {S} symbols x {N} 1-min bars, reference RSI rule",
                       "global_population": pop_global, "symbols": S, "bars": N, "parallelism": f"individuals sharded x{world}, market replicated, 1 all-gather/generation" if world > 1 else "single GPU",
                       "l2_policy": "inputs larger than L2 (price+RSI bank = %.2f GB per GPU)" % ((S * N * 4 + sweep.bank.numel() * 4) / 1e9)},
            "roofline": {"bound": "hbm", "kernel": path_kernels, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic, "traffic_unit": "bytes per launch of the dominant kernel (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "peak_source": peak_src, "bytes_per_eval": BYTES_PER_EVAL,
                         "kernel_ms": ms_kernel, "sweep_mode": path, "note": "achieved = 8 B x evals per sweep / CUDA-event duration of the sweep kernels (scan, verify/repair, metrics, fitness reduce); lanes sharing a (symbol, period) stream are served from L1/L2, so DRAM traffic is far below the algorithmic bytes (see profiles/)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "path": "MarketData(pinned host OHLCV; close prices uploaded, the fields the sweep does not read stay on the host) -> PopulationSweep (RSI bank) -> evaluate(list of dicts) -> host fitness"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if cpu_baseline is not None:
            if "configs0_backtest" in cpu_baseline:
                # the same configs[0] backtest through the GPU path (indicators, backtest_ref kernel, stats dict)
                import asyncio
                from ai_crypto_trader_b200.backtesting import StrategyTester
                df0 = configs0_frame()
                tester = StrategyTester(config={}, data_manager=None, results_dir=os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpurun_out", "bench_results"), config_path=None)
                asyncio.run(tester.backtest_frame(df0, "SYNUSDC"))
                t0 = time.perf_counter()
                st0 = asyncio.run(tester.backtest_frame(df0, "SYNUSDC"))
                cpu_baseline["configs0_backtest"]["gpu_path_ms"] = (time.perf_counter() - t0) * 1e3
                cpu_baseline["configs0_backtest"]["gpu_trades"] = int(st0["total_trades"])
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
