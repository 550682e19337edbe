"""Run the REFERENCE's own code as the oracle of record (build container only).

TEST INFRASTRUCTURE.  /root/reference exists only in the build container, so
this module is imported only by tests/golden/make_golden.py (which writes the
committed fixtures) and by tests that skip when the reference tree is absent.
Nothing on the GPU box imports it.

The reference's numeric code runs UNMODIFIED; only modules it imports for
plotting / networking (matplotlib, seaborn, redis, binance) are replaced by
empty stubs, as verified in SURVEY.md section 8c.  The reference opens
`logs/*.log` and `config.json` relative to the cwd at import time
(services/genetic_algorithm.py:22, services/strategy_evaluation.py:23,513), so
we chdir into a scratch directory holding a copy of its config.json.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("B200BT_REFERENCE_ROOT", "/root/reference"))


def available() -> bool:
    return (REFERENCE_ROOT / "services" / "strategy_evaluation.py").exists()


_scratch = None


def _prepare():
    global _scratch
    if _scratch is not None:
        return _scratch
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    for name in ["matplotlib", "matplotlib.pyplot", "seaborn", "redis", "redis.asyncio", "redis.exceptions",
                 "binance", "binance.client"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["redis.asyncio"].Redis = object
    sys.modules["redis.exceptions"].ConnectionError = Exception
    sys.modules["binance.client"].Client = object
    _scratch = Path(tempfile.mkdtemp(prefix="b200bt_ref_"))
    (_scratch / "logs").mkdir()
    shutil.copy(REFERENCE_ROOT / "config.json", _scratch / "config.json")
    os.chdir(_scratch)
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    return _scratch


def strategy_evaluation():
    """(StrategyEvaluationSystem instance, StrategyPerformanceMetrics class) of the reference."""
    _prepare()
    import logging
    from services.strategy_evaluation import StrategyEvaluationSystem, StrategyPerformanceMetrics
    logging.getLogger("strategy_evaluation").setLevel(logging.ERROR)
    return StrategyEvaluationSystem("config.json"), StrategyPerformanceMetrics


def genetic_algorithm_class():
    _prepare()
    import logging
    from services.genetic_algorithm import GeneticAlgorithm
    logging.getLogger("genetic_algorithm").setLevel(logging.ERROR)
    return GeneticAlgorithm


def monte_carlo_service(mc_params: dict):
    """A MonteCarloService built without __init__ (which needs Binance keys and
    rewrites config.json, monte_carlo_service.py:47-108)."""
    _prepare()
    import logging
    from services.monte_carlo_service import MonteCarloService
    logging.getLogger().setLevel(logging.ERROR)
    svc = object.__new__(MonteCarloService)
    svc.mc_params = dict(mc_params)
    svc.historical_data = {}
    svc.simulation_results = {}
    svc.last_simulation_time = {}
    return svc


def optimization_goals() -> dict:
    import json
    return json.loads((REFERENCE_ROOT / "config.json").read_text())["evolution"]["optimization_goals"]
