"""The oracle against the independent golden vectors (FFmpeg-encoded, see tests/golden/README.md)."""
import hashlib
import json
import os

import numpy as np
import pytest

from _harness import oracle_decode

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "manifest.json")) as f:
    MANIFEST = json.load(f)


@pytest.mark.parametrize("name", sorted(MANIFEST))
@pytest.mark.parametrize("chunk", [4096, 1000])
def test_oracle_matches_golden(name, chunk):
    e = MANIFEST[name]
    data = open(os.path.join(GOLD, e["file"]), "rb").read()
    assert hashlib.md5(data).hexdigest() == e["wv_md5"]
    out, errs, status, info = oracle_decode(data, e["open_flags"], chunk)
    assert status == 0 and errs == 0
    assert out.size == e["samples"] * e["reduced_channels"]
    assert info["bytes_per_sample"] == e["bytes_per_sample"]
    assert info["bits_per_sample"] == e["bits_per_sample"]
    assert hashlib.md5(np.ascontiguousarray(out, dtype="<i4").tobytes()).hexdigest() == e["int32_md5"]
