"""log2/exp2/nbits/ones_count tables: generated header == oracle's libm-derived tables == closed form,
and (when the read-only reference is present, i.e. in the authoring container) == the reference literals."""
import math
import os
import re

from _harness import refdec, wvenc

REF = "/root/reference/WordsUtils.cs"


def closed():
    lg = [int(math.floor(256 * math.log2(1 + i / 256) + 0.5)) for i in range(256)]
    ex = [int(math.floor(256 * (2 ** (i / 256) - 1) + 0.5)) for i in range(256)]
    return lg, ex


def test_tables_agree():
    lg, ex = closed()
    r, w = refdec(), wvenc()
    for i in range(256):
        assert r.rd_dbg_table(0, i) == lg[i] == w.wvenc_dbg_table(0, i)
        assert r.rd_dbg_table(1, i) == ex[i] == w.wvenc_dbg_table(1, i)
        assert r.rd_dbg_table(2, i) == i.bit_length()
        t = 0
        while (i >> t) & 1:
            t += 1
        assert r.rd_dbg_table(3, i) == t


def test_tables_match_reference_literals():
    if not os.path.exists(REF):
        import pytest
        pytest.skip("reference tree not present on this box")
    src = open(REF).read()

    def tab(name):
        m = re.search(name + r"\s*=\s*new int\[\]\s*\{([^}]*)\}", src, re.S)
        body = re.sub(r"//[^\n]*", "", m.group(1))
        return [int(x, 0) for x in body.replace("\n", " ").split(",") if x.strip()]

    lg, ex = closed()
    assert tab("log2_table") == lg
    assert tab("exp2_table") == ex
    assert tab("nbits_table") == [i.bit_length() for i in range(256)]
