#!/usr/bin/env python3
"""Turn the round-2 final captures in gpurun_out/ (tools/ncu_r2_final.sh) into the committed evidence under profiles/:
per-kernel ncu summaries, the launch list + its share summary, and profiles/roofline_traffic.json (what bench.py reads)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def summary(rep, title, dst):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, title], capture_output=True, text=True).stdout
    open(dst, "w").write(txt)
    print(txt)


def main():
    for rep, title, dst in [
        ("r02_fix_bench", "headline kernel at bench size (200 000 blocks), round-2 final: staged output, 88 registers", "r02_ncu_fixed_kernel_bench_size.txt"),
        ("r02_dsdhigh", "DSD high (mode 3), 160 000 blocks, round-2 final (loop-free renormalisation)", "r02_ncu_dsd_high.txt"),
        ("r02_dsdfast", "DSD fast (mode 1) decode kernel, 160 000 blocks, round-2 final", "r02_ncu_dsd_fast.txt"),
        ("r02_t16", "24-bit 5.1 / 16 terms, all six channels, round-2 final", "r02_ncu_16term.txt"),
    ]:
        path = os.path.join(G, rep + ".ncu-rep")
        if os.path.exists(path):
            summary(path, title, os.path.join(P, dst))
    path = os.path.join(G, "r02_fix_bench.ncu-rep")
    if os.path.exists(path):
        v, _u = raw(path)

        def num(k):
            return float(v[k].replace(",", ""))
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        _v, u = raw(path)
        rd = num("dram__bytes_read.sum") * unit[u["dram__bytes_read.sum"]]
        wr = num("dram__bytes_write.sum") * unit[u["dram__bytes_write.sum"]]
        grid, cta = int(num("launch__grid_size")), int(num("launch__block_size"))
        blocks = 200000 if (grid - 1) * cta < 200000 <= grid * cta else None  # one WavPack block per thread
        json.dump({"kernel": "k_decode_pcm<stereo,lossless,FixedDecorr<-2,3,2,18,18>,F16> (staged output)", "blocks_per_launch": blocks,
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                   "warp_instructions_per_launch": num("smsp__inst_executed.sum"),
                   "source": "profiles/r02_ncu_fixed_kernel_bench_size.txt (ncu --set full, one launch of the bench-size batch, round 2)"},
                  open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
        print(open(os.path.join(P, "roofline_traffic.json")).read())
    lp = os.path.join(G, "r02_launches.csv")
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
        hdr = rows[0]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        tot, cnt = defaultdict(float), defaultdict(int)
        for r in rows[1:]:
            tot[r[ik]] += float(r[iv].replace(",", "")) / 1e6
            cnt[r[ik]] += 1
        allms = sum(tot.values())
        with open(os.path.join(P, "r02_launches_summary.txt"), "w") as f:
            f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400, command: python bench.py --steps 2 --warmup 1\n")
            f.write("(per-launch times under ncu are serialised and cold-cache: compare SHARES; the list covers the warm-up and timed steps of the\n headline workload, its e2e / verify legs and the short launches of the `configs` object, up to 400 launches)\n\n")
            for k, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
                f.write("%5d launches %10.3f ms %5.1f%%  %s\n" % (cnt[k], ms, 100 * ms / allms, k[:150]))
        subprocess.run(["cp", lp, os.path.join(P, "r02_launches.csv")])
        print(open(os.path.join(P, "r02_launches_summary.txt")).read())


if __name__ == "__main__":
    main()
