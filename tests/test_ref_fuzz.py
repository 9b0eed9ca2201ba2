"""Randomised sweeps of the oracle against the reference's own compiled code (oracle/_ref/*.so, see tests/test_ref_*.py for
what each library is): scripts/fuzz_ref_{solver,sc,fe}.py with a fixed seed and a size that runs in seconds.  The long
runs behind DESIGN section 2 (9 900 solver cases, 46 000 ScanContext pairs, 1 040 front-end cases, the RadarLoader chain: no mismatch) use the same
scripts with other seeds."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.parametrize("name,lib,seed,n", [("solver", "libref_solver.so", 11, 120), ("sc", "libref_sc.so", 12, 800), ("fe", "libref_frontend.so", 13, 24)])
def test_fuzz(name, lib, seed, n):
    if not os.path.exists(os.path.join(REF, lib)):
        pytest.skip(f"oracle/_ref/{lib} not built (needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", f"fuzz_ref_{name}.py"), str(seed), str(n)], stdin=subprocess.DEVNULL,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    assert "done, mismatches: 0" in r.stderr, r.stderr[-1500:]
