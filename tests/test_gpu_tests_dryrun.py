"""The GPU test files of the kernels that have not met a GPU yet, dry-run on the CPU against a mock Context built from
the host twins and the oracle (tools/dryrun_gpu_tests.py): keeps their LOGIC (corrupted offsets, expected verdicts and
first-failure codes, wrapper arguments) green in the CPU suite, so that the first `pytest -m gpu` run measures the kernels
and not typos in the tests."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_test_files_dry_run():
    args = [] if os.environ.get("SVB_DRYRUN_FULL") else ["--quick"]      # the full dry run: SVB_DRYRUN_FULL=1 (3 minutes)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dryrun_gpu_tests.py")] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dry run ok:" in r.stdout
