#!/usr/bin/env python3
"""Extract the public surface of the reference's WavPackUtils.cs (names, return types, parameter lists) into
tests/golden/reference_api.json, so that the C# shim can be checked against it on boxes without /root/reference.
    python tools/extract_reference_api.py [/root/reference]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIG = re.compile(r"^\s*public\s+static\s+([\w\.\[\]<>]+)\s+(\w+)\s*\(([^)]*)\)", re.M)


def normalise_type(t):
    return t.replace("System.IO.", "").strip()


def parse_params(plist):
    out = []
    for prm in [x.strip() for x in plist.split(",") if x.strip()]:
        default = None
        if "=" in prm:
            prm, default = [x.strip() for x in prm.split("=", 1)]
        ptype, pname = prm.rsplit(None, 1)
        out.append({"type": normalise_type(ptype), "name": pname, "default": default})
    return out


def public_static_methods(text):
    return [{"name": m.group(2), "returns": normalise_type(m.group(1)), "params": parse_params(m.group(3))} for m in SIG.finditer(text)]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    api = public_static_methods(open(os.path.join(ref, "WavPackUtils.cs"), encoding="utf-8-sig").read())
    demo = open(os.path.join(ref, "WvDemo.cs"), encoding="utf-8-sig").read()
    used = sorted(set(re.findall(r"WavPackUtils\.(\w+)\s*\(", demo)))
    out = {"source": "Quake4/WavPackDecoder WavPackUtils.cs (public static methods) and the calls WvDemo.cs makes", "methods": api, "wvdemo_calls": used}
    with open(os.path.join(ROOT, "tests", "golden", "reference_api.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("%d methods, %d used by WvDemo" % (len(api), len(used)))


if __name__ == "__main__":
    main()
