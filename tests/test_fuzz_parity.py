"""Randomly damaged streams (bit flips, truncation, deleted / inserted bytes) through the index pass + the host-compiled device
decode function versus the oracle.  Any difference must be covered by WVB_RF_INEXACT / stopped_early."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_short_campaign():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "120", "7"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "'mismatch': 0" in r.stdout
