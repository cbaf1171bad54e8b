"""The C# shim (csharp/WavPackUtils.cs) cannot be compiled here (no .NET / mono in the image), so its contract is
checked textually: same public static surface as the reference's WavPackUtils.cs, P/Invoke prototypes that match
include/wvb.h, and [StructLayout] mirrors whose offsets equal the compiled C structs (wvb_abi_layout)."""
import json
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import extract_reference_api as X  # noqa: E402

SHIM = open(os.path.join(ROOT, "csharp", "WavPackUtils.cs")).read()
API = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_api.json")))


def test_fixture_is_current_when_the_reference_is_here():
    ref = "/root/reference/WavPackUtils.cs"
    if not os.path.exists(ref):
        pytest.skip("reference tree not on this box; the committed fixture stands")
    assert X.public_static_methods(open(ref, encoding="utf-8-sig").read()) == API["methods"]


def test_every_public_static_method_of_the_reference_exists_with_the_same_signature():
    shim = {m["name"]: m for m in X.public_static_methods(SHIM)}
    assert len(API["methods"]) == 24
    for m in API["methods"]:
        assert m["name"] in shim, "missing in the shim: %s" % m["name"]
        s = shim[m["name"]]
        assert s["returns"] == m["returns"], (m["name"], s["returns"], m["returns"])
        assert [(p["type"], p["name"], p["default"]) for p in s["params"]] == [(p["type"], p["name"], p["default"]) for p in m["params"]], m["name"]


def test_every_call_wvdemo_makes_resolves():
    shim = {m["name"] for m in X.public_static_methods(SHIM)}
    assert set(API["wvdemo_calls"]) <= shim
    assert "SAMPLE_BUFFER_SIZE" in SHIM  # WvDemo.cs:110


CS_SIZES = {"byte": 1, "sbyte": 1, "short": 2, "ushort": 2, "int": 4, "uint": 4, "long": 8, "ulong": 8}


def _cs_struct_layout(name):
    """Sequential layout (natural alignment, Pack = 8) of a struct declared in the shim: {field: (offset, size)}, total size."""
    body = re.search(r"struct\s+%s\b[^{]*\{(.*?)\n        \}" % name, SHIM, re.S).group(1)
    off, fields, max_align = 0, {}, 1
    for decl in re.findall(r"public\s+(fixed\s+)?(\w+)\s+([^;]+);", body):
        fixed, typ, names = decl
        sz = CS_SIZES[typ]
        max_align = max(max_align, sz)
        for nm in [x.strip() for x in names.split(",")]:
            count = 1
            m = re.match(r"(\w+)\[(\d+)\]", nm)
            if m:
                nm, count = m.group(1), int(m.group(2))
            off = (off + sz - 1) // sz * sz
            fields[nm] = (off, sz * count)
            off += sz * count
    return fields, (off + max_align - 1) // max_align * max_align


def test_struct_mirrors_match_the_compiled_layout():
    from wavpackdecoder_b200 import _native as N, build
    build.build()
    lib = N.load()
    names = {"wvb_block_desc": "BlockDesc", "wvb_block_result": "BlockResult", "wvb_file_info": "FileInfo", "wvb_seek_state": "SeekState"}
    seen = 0
    for part in lib.wvb_abi_layout().decode().split("|"):
        items = [x for x in part.split(";") if x]
        cname, size = items[0].split(":")
        fields, total = _cs_struct_layout(names[cname])
        assert total == int(size), (cname, total, size)
        assert len(fields) == len(items) - 1, cname
        for it in items[1:]:
            f, off, sz = it.split(":")
            assert fields[f] == (int(off), int(sz)), (cname, f, fields[f], off, sz)
            seen += 1
    assert seen > 60


def test_pinvoke_prototypes_match_the_header():
    hdr = open(os.path.join(ROOT, "include", "wvb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decls = {m.group(1): m.group(2) for m in re.finditer(r"\b(wvb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, re.S)}

    def nargs(s):
        s = s.strip()
        return 0 if s in ("", "void") else len(s.split(","))

    imports = re.findall(r"\[DllImport\(Lib\)\]\s*internal static extern (?:unsafe )?[\w\*]+ (wvb_\w+)\(([^)]*)\)", SHIM, re.S)
    assert len(imports) >= 12
    for name, params in imports:
        assert name in decls, "the shim imports %s, which include/wvb.h does not declare" % name
        assert nargs(params) == nargs(decls[name]), (name, params, decls[name])
    assert "WVB_ABI_VERSION = %s" % re.search(r"#define WVB_ABI_VERSION (\d+)", hdr).group(1) in SHIM
