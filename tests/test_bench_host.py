"""bench.py's host-side safety net (no GPU): child processes with deadlines, the reference arm's JSON line."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def test_run_child_returns_output_and_kills_the_whole_group_on_timeout(tmp_path):
    import bench
    rc, out, err = bench._run_child([sys.executable, "-c", "import sys; print('hello'); print('oops', file=sys.stderr); sys.exit(3)"], 30)
    assert (rc, out.strip(), err.strip()) == (3, "hello", "oops")
    pidfile = tmp_path / "grandchild.pid"
    code = ("import os, time\n"
            "pid = os.fork()\n"
            "if pid == 0:\n"
            "    time.sleep(120)\n"
            f"open({str(pidfile)!r}, 'w').write(str(pid))\n"
            "time.sleep(120)\n")
    t0 = time.time()
    with pytest.raises(subprocess.TimeoutExpired):
        bench._run_child([sys.executable, "-c", code], 2)
    assert time.time() - t0 < 30
    grandchild = int(pidfile.read_text())
    for _ in range(50):                       # (the kernel needs a moment to reap it)
        try:
            os.kill(grandchild, 0)
        except ProcessLookupError:
            break
        time.sleep(0.1)
    else:
        # still in the process table: it must at least be a zombie of a killed process, not a sleeper
        state = Path(f"/proc/{grandchild}/stat").read_text().split()[2]
        assert state == "Z", state


def test_cpu_baseline_child_reports_a_missed_deadline_instead_of_stalling():
    import bench
    t0 = time.time()
    r = bench.cpu_baseline_child(timeout_s=0.2, attempts=1)
    assert time.time() - t0 < 30
    assert r["value"] is None and "no result within" in r["unavailable"] and r["unit"] == bench.UNIT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the reference's own CPU path through a child process with a deadline): one JSON line."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "bar_strategy_evals_per_sec" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    # a rank other than 0 of a multi-process launch does no work and prints nothing
    r2 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], stdout=subprocess.PIPE,
                        stderr=subprocess.PIPE, text=True, timeout=120, cwd=str(ROOT), env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_claimed_stdout_carries_only_the_result_line():
    """bench._claim_stdout: whatever the process (or a C library) prints to fd 1 afterwards lands on stderr."""
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench\n"
            "fd = bench._claim_stdout()\n"
            "print('a python print'); os.system('echo a child of the shell')\n"
            "os.write(1, b'a C library writing to fd 1\\n')\n"
            "os.write(fd, (json.dumps({'value': 1}) + '\\n').encode())\n") % str(ROOT)
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"value": 1}\n'
    for s in ("a python print", "a child of the shell", "a C library writing to fd 1"):
        assert s in r.stderr
