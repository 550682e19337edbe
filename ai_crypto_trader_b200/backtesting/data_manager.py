"""Historical market-data store: the reference's CSV layout plus a binary fp32 sidecar.

Reference: backtesting/data_manager.py (class HistoricalDataManager).  Kept: directory layout
`backtesting/data/market/<SYMBOL>/<interval>_<YYYYMMDD>_<YYYYMMDD>.csv` (:191-193), the
`timestamp`-indexed DataFrame `load_market_data` returns (:214-265), the cache keyed by
(symbol, interval, range) (:218-220), `available_symbols/intervals`, `get_data_range`.
Out of scope: the Binance / LunarCrush fetchers (network).  New: `save_sidecar` /
`load_ohlcv32` keep a float32 [5][N] + int64 minute-timestamp binary next to each CSV so that
large frames reach pinned host memory without CSV parsing.
"""
from __future__ import annotations

import json
import logging
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd

logger = logging.getLogger("b200bt.data_manager")
FIELDS = ("open", "high", "low", "close", "volume")


class HistoricalDataManager:
    def __init__(self, config_path: Optional[str] = "config.json", data_dir: Optional[str] = None):
        try:
            with open(config_path, "r") as f:
                self.config = json.load(f)
        except (OSError, TypeError):
            self.config = {}
        self.data_dir = Path(data_dir) if data_dir else Path("backtesting/data")
        self.market_data_dir = self.data_dir / "market"
        self.social_data_dir = self.data_dir / "social"
        self.market_data_dir.mkdir(parents=True, exist_ok=True)
        self.social_data_dir.mkdir(parents=True, exist_ok=True)
        self.market_data_cache: Dict[str, pd.DataFrame] = {}
        self.social_data_cache: Dict[str, pd.DataFrame] = {}

    # -- writing (what fetch_and_save_data does after the network fetch, :191-196) -----
    def save_market_data(self, symbol: str, interval: str, df: pd.DataFrame, start_date: datetime,
                         end_date: datetime, sidecar: bool = True) -> Path:
        d = self.market_data_dir / symbol
        d.mkdir(exist_ok=True)
        path = d / f"{interval}_{start_date.strftime('%Y%m%d')}_{end_date.strftime('%Y%m%d')}.csv"
        out = df.copy()
        out.index.name = "timestamp"
        out.to_csv(path)
        if sidecar:
            self.save_sidecar(path, out)
        return path

    @staticmethod
    def save_sidecar(csv_path: Path, df: pd.DataFrame) -> Path:
        """fp32 [5][N] + int64 epoch SECONDS next to the CSV, stamped with the CSV's size and mtime: a CSV that is
        rewritten afterwards (the reference's fetcher, :191-196) invalidates the sidecar."""
        side = Path(str(csv_path) + ".f32.npz")
        seconds = df.index.astype("datetime64[s]").astype(np.int64).to_numpy()
        ohlcv = np.stack([df[f].to_numpy(dtype=np.float32) for f in FIELDS])
        st = Path(csv_path).stat()
        np.savez(side, ohlcv=ohlcv, seconds=seconds, csv_stamp=np.array([st.st_size, st.st_mtime_ns], dtype=np.int64))
        return side

    @staticmethod
    def _sidecar_if_fresh(csv_path: Path):
        """The sidecar's arrays when it exists and still describes `csv_path` (same size and mtime), else None."""
        side = Path(str(csv_path) + ".f32.npz")
        if not side.exists():
            return None
        try:
            z = np.load(side)
            st = Path(csv_path).stat()
            if "csv_stamp" not in z or "seconds" not in z or z["csv_stamp"].tolist() != [st.st_size, st.st_mtime_ns]:
                return None
            return z["seconds"], z["ohlcv"]
        except Exception as e:
            logger.error("Error loading sidecar %s: %s", side, e)
            return None

    # -- reading ------------------------------------------------------------------
    def load_market_data(self, symbol: str, interval: str, start_date: datetime, end_date: datetime = None) -> pd.DataFrame:
        cache_key = f"{symbol}_{interval}_{start_date.strftime('%Y%m%d')}_{end_date.strftime('%Y%m%d') if end_date else 'now'}"
        if cache_key in self.market_data_cache:
            return self.market_data_cache[cache_key]
        symbol_dir = self.market_data_dir / symbol
        if not symbol_dir.exists():
            logger.warning("No data directory found for %s", symbol)
            return pd.DataFrame()
        files = list(symbol_dir.glob(f"{interval}_*.csv"))
        if not files:
            logger.warning("No %s data files found for %s", interval, symbol)
            return pd.DataFrame()
        if end_date is None:
            end_date = datetime.now()
        parts = []
        for path in files:
            try:
                df = pd.read_csv(path)
                df["timestamp"] = pd.to_datetime(df["timestamp"])
                df.set_index("timestamp", inplace=True)
                df = df[(df.index >= start_date) & (df.index <= end_date)]
                if not df.empty:
                    parts.append(df)
            except Exception as e:
                logger.error("Error loading file %s: %s", path, e)
        if not parts:
            logger.warning("No data found for %s in specified date range", symbol)
            return pd.DataFrame()
        result = pd.concat(parts).sort_index()
        result = result[~result.index.duplicated(keep="first")]
        self.market_data_cache[cache_key] = result
        return result

    def load_ohlcv32(self, symbol: str, interval: str, start_date: datetime, end_date: datetime = None):
        """(float32 [5][N], int64 minutes[N]) from the binary sidecars when every matching CSV has a fresh one
        (`_sidecar_if_fresh`); otherwise the CSVs are parsed (same values, rounded to fp32).  Rows are de-duplicated on
        their timestamp in seconds (keep first, like load_market_data :258); `minutes` = seconds // 60."""
        symbol_dir = self.market_data_dir / symbol
        files = sorted(symbol_dir.glob(f"{interval}_*.csv")) if symbol_dir.exists() else []
        sides = [self._sidecar_if_fresh(p) for p in files]
        if files and all(s is not None for s in sides):
            if end_date is None:
                end_date = datetime.now()
            lo = int(pd.Timestamp(start_date).timestamp())
            hi = int(pd.Timestamp(end_date).timestamp())
            chunks = []
            for sec, ohlcv in sides:
                m = (sec >= lo) & (sec <= hi)
                chunks.append((sec[m], ohlcv[:, m]))
            seconds = np.concatenate([c[0] for c in chunks])
            ohlcv = np.concatenate([c[1] for c in chunks], axis=1)
            order = np.argsort(seconds, kind="stable")
            seconds, ohlcv = seconds[order], ohlcv[:, order]
            keep = np.concatenate([[True], seconds[1:] != seconds[:-1]]) if len(seconds) else np.zeros(0, dtype=bool)
            return np.ascontiguousarray(ohlcv[:, keep]), seconds[keep] // 60
        df = self.load_market_data(symbol, interval, start_date, end_date)
        if df.empty:
            return np.zeros((5, 0), dtype=np.float32), np.zeros(0, dtype=np.int64)
        minutes = (df.index.astype("datetime64[s]").astype(np.int64) // 60).to_numpy()
        return np.stack([df[f].to_numpy(dtype=np.float32) for f in FIELDS]), minutes

    def load_social_data(self, symbol: str, start_date: datetime, end_date: datetime = None) -> pd.DataFrame:
        return pd.DataFrame()      # LunarCrush social CSVs feed only the LLM prompt (SURVEY #11): out of scope

    def merge_market_and_social_data(self, symbol: str, interval: str, start_date: datetime,
                                     end_date: datetime = None) -> pd.DataFrame:
        return self.load_market_data(symbol, interval, start_date, end_date)      # :386-389 (no social data)

    def available_symbols(self) -> List[str]:
        return sorted(p.name for p in self.market_data_dir.iterdir() if p.is_dir())

    def available_intervals(self, symbol: str) -> List[str]:
        d = self.market_data_dir / symbol
        return sorted({p.name.split("_")[0] for p in d.glob("*.csv")}) if d.exists() else []

    def get_data_range(self, symbol: str, interval: str) -> Tuple[Optional[datetime], Optional[datetime]]:
        d = self.market_data_dir / symbol
        if not d.exists():
            return None, None
        starts, ends = [], []
        for p in d.glob(f"{interval}_*.csv"):
            parts = p.stem.split("_")
            if len(parts) >= 3:
                try:
                    starts.append(datetime.strptime(parts[1], "%Y%m%d"))
                    ends.append(datetime.strptime(parts[2], "%Y%m%d"))
                except ValueError:
                    pass
        return (min(starts), max(ends)) if starts and ends else (None, None)
