#!/usr/bin/env python3
"""bench.py -- batch WavPack decode throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W                 # our arm (CUDA, libwvb.so)
  python bench.py --impl reference --gpus N --steps K --warmup W  # reference arm: the reference's CPU decode path
  torchrun --nproc-per-node N bench.py --gpus N ...             # one rank per GPU, no data-path collective

Workload (config.workload): BASELINE.json configs[1] -- 10 000 synthetic 16-bit stereo 44.1 kHz default-mode
.wv files of 10 s per GPU (weak scaling: every rank decodes its own 10 000 files; shards share nothing).
A step = one decode pass over the whole batch.
  value  : decoded complete samples/s, whole job, compressed input and PCM output resident in HBM
  e2e    : same metric through the public C-ABI call with HOST (pinned) buffers, every step: one wvb_batch_decode_files call =
           host index pass (overlapped with the copies) + H2D of the compressed slab + kernels + D2H of the PCM
  e2e.pcie_ceiling: the same pinned buffers and byte counts moved with no decode (H2D and D2H at once, all ranks at once):
           the roofline of the end-to-end path; e2e.frac_of_ceiling = e2e / that
  roofline: algorithmic bytes (compressed block bytes in + PCM bytes out) / CUDA-event kernel time vs measured HBM peak
  configs : the other BASELINE.json configs (1: one 60 s file, the latency case; 3: 24-bit 5.1 with 16 terms; 4: float /
           int32 / hybrid; 5: DSD64 modes 0/1/3), device-resident, short launches, each validated against the oracle.
           Under torchrun configs 3 and 5 are ONE logical corpus cut by decode cost over the ranks (strong scaling),
           per-file results gathered on rank 0
  cpu_baseline: the oracle (C restatement of the reference, one file per thread on all host cores) on a bounded sample
The C# reference cannot run in this image (no .NET); the reference arm therefore times the C restatement ("port").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "decoded_samples_per_s"
UNIT = "samples/s"


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print(*a, file=sys.stderr, flush=True)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------
# synthetic corpus
# --------------------------------------------------------------------------------------
def build_corpus(nfiles, seconds, base_seed, threads, budget_s, pin, cfg_kw=None):
    """Encode up to `nfiles` unique files within ~budget_s (the rest are byte copies of earlier files placed at new
    offsets; warps still hold 32 different blocks because replicas are nfiles_unique files apart)."""
    import _harness as H
    import torch
    cfg = H.make_config(**(cfg_kw or {}))
    n_per = int(cfg.sample_rate * seconds) * (8 if cfg.kind == H.KIND_DSD else 1)  # DSD: byte-times (8 one-bit samples each)
    lib = H.wvenc()
    # probe: size and speed
    probe_n = max(2, min(threads, nfiles))
    bound = lib.wvenc_bound(C.byref(cfg), n_per)
    tmp = np.zeros(bound * probe_n, dtype=np.uint8)
    offs = np.zeros(probe_n, dtype=np.uint64)
    sizes = np.zeros(probe_n, dtype=np.uint64)
    t0 = time.perf_counter()
    used = lib.wvenc_build_corpus(C.byref(cfg), n_per, probe_n, base_seed, threads, tmp.ctypes.data, tmp.size, offs.ctypes.data, sizes.ctypes.data)
    dt = time.perf_counter() - t0
    assert used > 0
    per_file = int(sizes.max())
    rate = probe_n / dt
    unique = int(min(nfiles, max(probe_n, min(nfiles, rate * budget_s))))
    slot = (int(per_file * 1.03) + 4096 + 63) & ~63
    total_cap = slot * nfiles + 4096
    slab_t = torch.empty(total_cap, dtype=torch.uint8, pin_memory=pin)
    slab = slab_t.numpy()
    offsets = np.zeros(nfiles, dtype=np.uint64)
    fsizes = np.zeros(nfiles, dtype=np.uint64)
    t0 = time.perf_counter()
    used = lib.wvenc_build_corpus(C.byref(cfg), n_per, unique, base_seed, threads, slab.ctypes.data, slot * unique, offsets.ctypes.data, fsizes.ctypes.data)
    assert used > 0, "corpus generation overflowed its slab"
    gen_s = time.perf_counter() - t0
    # table order == slab order (the generator's threads bump-allocate): lets the library pipeline H2D / kernels / D2H by segment
    srt = np.argsort(offsets[:unique], kind="stable")
    offsets[:unique] = offsets[:unique][srt]
    fsizes[:unique] = fsizes[:unique][srt]
    pos = (int(used) + 63) & ~63
    for i in range(unique, nfiles):
        j = i % unique
        ln = int(fsizes[j])
        o = int(offsets[j])
        slab[pos:pos + ln] = slab[o:o + ln]
        offsets[i] = pos
        fsizes[i] = ln
        pos += (ln + 63) & ~63
    slab[pos:pos + 64] = 0
    return dict(cfg=cfg, slab_t=slab_t, slab=slab[:pos + 64], slab_bytes=pos + 64, offsets=offsets, sizes=fsizes, unique=unique,
                samples_per_file=n_per, gen_s=gen_s, compressed_bytes=int(fsizes.sum()))


# --------------------------------------------------------------------------------------
# CPU baseline: oracle (restated reference) over a bounded sample, one file per thread
# --------------------------------------------------------------------------------------
def cpu_decode_sample(corpus, nfiles_sample, threads):
    import _harness as H
    lib = H.refdec()
    slab, offsets, sizes = corpus["slab"], corpus["offsets"], corpus["sizes"]
    n_per = corpus["samples_per_file"]
    pcm_cap = n_per * 4 + 64

    def work(i):
        buf = np.empty(pcm_cap, dtype=np.uint8)
        ln = C.c_size_t()
        errs = C.c_long()
        n = lib.rd_decode_file_pcm(slab.ctypes.data + int(offsets[i]), int(sizes[i]), 0, 4096, buf.ctypes.data, pcm_cap, C.byref(ln), C.byref(errs))
        return n, errs.value

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(work, range(nfiles_sample)))
    dt = time.perf_counter() - t0
    samples = sum(r[0] for r in res)
    assert all(r[1] == 0 for r in res)
    return samples, dt


def size_cpu_sample(corpus, threads, target_s):
    """Pick a sample size so that the CPU leg takes about target_s."""
    samples, dt = cpu_decode_sample(corpus, min(threads, len(corpus["offsets"])), threads)
    rate = samples / dt
    n = int(max(threads, min(len(corpus["offsets"]), rate * target_s / corpus["samples_per_file"])))
    return n


# --------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# --------------------------------------------------------------------------------------
# copy-only ceiling of the end-to-end path
# --------------------------------------------------------------------------------------
def pcie_ceiling(slab_t, n_in, out_t, n_out, dev, barrier, max_over_ranks, reps=3):
    """The step's bytes over PCIe with no decode: H2D of the compressed slab and D2H of the PCM at the same time, from / into
    the very pinned buffers the e2e leg uses, all ranks at once (barrier before, max over ranks after).  Best of `reps`."""
    import torch
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(do_in, do_out):
        best = None
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(slab_t[:n_in], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    out_t[:n_out].copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0, dev)
            best = dt if best is None else min(best, dt)
        return best

    run(True, True)
    both, h2d, d2h = run(True, True), run(True, False), run(False, True)
    del d_in, d_out
    return {"both_s": both, "h2d_alone_gbs_per_gpu": n_in / h2d / 1e9, "d2h_alone_gbs_per_gpu": n_out / d2h / 1e9,
            "d2h_gbs_per_gpu_while_h2d": n_out / both / 1e9}


# --------------------------------------------------------------------------------------
# the other BASELINE configs: short device-resident launches
# --------------------------------------------------------------------------------------
T16 = [18, 18, 2, 3, -2, 18, 2, 4, 7, 5, 3, 6, 8, -1, 18, 2]
# name -> (BASELINE config, encoder kwargs, open flags, files per launch at 1 GPU, seconds per file, sharded under torchrun)
EXTRA_CONFIGS = [
    ("config1_one_60s_file", 1, dict(), 0, 1, 60.0, False),
    ("config3_24bit_51_16terms_2ch_max", 3, dict(bits=24, channels=6, sample_rate=48000, block_samples=24000, terms=T16, deltas=[2] * 16), 0x8, 6000, 5.0, True),
    ("config3_24bit_51_16terms_all_channels", 3, dict(bits=24, channels=6, sample_rate=48000, block_samples=24000, terms=T16, deltas=[2] * 16), 0x10000, 2000, 5.0, False),
    ("config4a_float", 4, dict(kind=2, bits=32), 0, 6000, 10.0, False),
    ("config4b_int32_wvx", 4, dict(bits=32, int32_sent_bits=8), 0, 6000, 10.0, False),
    ("config4c_hybrid_stereo", 4, dict(kind=1), 0, 8000, 10.0, False),
    ("config4c_hybrid_mono", 4, dict(kind=1, channels=1, terms=[18, 18, 2, 3], deltas=[2, 2, 2, 2]), 0, 12000, 10.0, False),
    ("config5_dsd64_raw", 5, dict(kind=3, dsd_mode=0, block_samples=22050), 0, 1000, 10.0, True),
    ("config5_dsd64_fast", 5, dict(kind=3, dsd_mode=1, block_samples=22050), 0, 1000, 10.0, True),
    ("config5_dsd64_high", 5, dict(kind=3, dsd_mode=3, block_samples=22050), 0, 1000, 10.0, True),
]


def unique_files(cfg_kw, seconds, seed, threads, budget_s, want):
    """Up to `want` unique encoded files within ~budget_s: (slab, offsets, sizes, samples per file)."""
    c = build_corpus(want, seconds, seed, threads, budget_s, pin=False, cfg_kw=cfg_kw)
    u = c["unique"]
    return c["slab"], c["offsets"][:u].copy(), c["sizes"][:u].copy(), c["samples_per_file"]


def run_extra_config(name, cfg_kw, open_flags, nfiles, seconds, sharded, args, rank, world, dev, threads, peak, barrier):
    """One BASELINE config as a device-resident launch.  The logical corpus is `nfiles` files (unique encodes of two
    durations, replicated); under torchrun a sharded config is cut into contiguous ranges of equal decode cost
    (wavpackdecoder_b200.sharding) and every rank materialises and decodes its own range."""
    import torch
    import torch.distributed as dist
    import _harness as H
    from wavpackdecoder_b200 import _native as N
    from wavpackdecoder_b200.batch import BatchDecoder, Corpus
    from wavpackdecoder_b200.sharding import file_costs, max_over_ranks, shard_contiguous_by_cost
    budget = args.config_gen_budget_s
    want = max(2, min(nfiles, 4000))
    # two durations, so that the cost-based cut is not the trivial equal split
    sets = [unique_files(cfg_kw, seconds, 0xC0DE0000, threads, budget * 0.6, want)]
    if nfiles > 1:
        sets.append(unique_files(cfg_kw, seconds / 2, 0xC0DE8000, threads, budget * 0.4, want))
    fmt = N.OUT_PCM
    set_costs, set_samples = [], []
    for slab_u, offs_u, sizes_u, _ in sets:
        cu = Corpus(slab_u, offs_u, sizes_u, open_flags=open_flags, out_format=fmt, threads=threads)
        assert all(cu.infos[i].status == 0 for i in range(cu.nfiles)), name
        set_costs.append(file_costs(cu))
        set_samples.append(np.array([int(cu.infos[i].indexed_samples) for i in range(cu.nfiles)], dtype=np.int64))
    ids = np.arange(nfiles)
    which = (ids % 3 == 2).astype(np.int64) if len(sets) > 1 else np.zeros(nfiles, dtype=np.int64)  # every third file is a short one
    uid = np.where(which == 0, ids % len(set_costs[0]), ids % len(set_costs[-1]))
    costs = np.where(which == 0, set_costs[0][uid % len(set_costs[0])], set_costs[-1][uid % len(set_costs[-1])])
    w = world if sharded else 1
    lo, hi = shard_contiguous_by_cost(costs, w)[rank if sharded else 0]
    # materialise this rank's range
    sizes = np.array([int(sets[which[i]][2][uid[i]]) for i in range(lo, hi)], dtype=np.uint64)
    offsets = np.zeros(hi - lo, dtype=np.uint64)
    pos = 0
    for k in range(hi - lo):
        offsets[k] = pos
        pos += (int(sizes[k]) + 63) & ~63
    slab = np.zeros(pos + 64, dtype=np.uint8)
    for k, i in enumerate(range(lo, hi)):
        sl, so, ss, _ = sets[which[i]]
        o = int(so[uid[i]])
        slab[int(offsets[k]):int(offsets[k]) + int(sizes[k])] = sl[o:o + int(sizes[k])]
    cp = Corpus(slab, offsets, sizes, open_flags=open_flags, out_format=fmt, threads=threads)
    dec = BatchDecoder(dev.index)
    try:
        d_in = torch.from_numpy(slab).to(dev)
        d_out = torch.empty(cp.out_bytes + 64, dtype=torch.uint8, device=dev)
        d_res = torch.empty(max(cp.nblocks, 1) * 16, dtype=torch.uint8, device=dev)
        dec.prepare(cp.descs, cp.nblocks, fmt)
        FL = N.IN_DEVICE | N.OUT_DEVICE | N.RESULTS_DEVICE

        def step():
            dec.decode(d_in.data_ptr(), slab.size, None, cp.nblocks, d_out.data_ptr(), cp.out_bytes, fmt, FL, d_res.data_ptr())

        for _ in range(2):
            step()
        # validation: every block clean, first and last file of the range byte-identical with the oracle
        res = d_res.cpu().numpy().view(np.uint32).reshape(-1, 4)[:cp.nblocks]
        flagged = int((res[:, 1] != 0).sum())
        ok = flagged == 0
        for i in sorted({0, cp.nfiles - 1}) if cp.nfiles else []:
            data = slab[int(offsets[i]):int(offsets[i]) + int(sizes[i])].tobytes()
            if open_flags & 0x10000:
                continue  # (the all-channels extension has no whole-file oracle; its blocks are CRC-checked above)
            ref, errs, status, info = H.oracle_decode(data, open_flags)
            o = int(cp.file_out_offset[i])
            got = d_out[o:o + ref.size * info["bytes_per_sample"]].cpu().numpy()
            ok = ok and status == 0 and errs == 0 and np.array_equal(got, H.format_samples(ref, info["bytes_per_sample"]))
        kernel_ms = []
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.config_steps):
            step()
            kernel_ms.append(dec.timing()["kernel_ms"])
        torch.cuda.synchronize()
        elapsed = max_over_ranks(time.perf_counter() - t0, dev) if sharded else time.perf_counter() - t0
    finally:
        dec.close()
    local = np.array([cp.total_samples, cp.nblocks, int(sizes.sum()), cp.out_bytes, flagged, int(ok), hi - lo, float(np.mean(kernel_ms)) * 1e3], dtype=np.int64)
    if sharded and world > 1:
        t = torch.from_numpy(local).to(dev)
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)  # per-rank results gathered (rank 0 prints them); the data path itself has no collective
        allr = torch.stack(allr).cpu().numpy()
    else:
        allr = local[None, :]
    tot = allr.sum(axis=0)
    k_ms = float(allr[:, 7].max()) / 1e3
    algo = int(tot[2] + tot[3])
    out = {"baseline_config": None, "scaling": "strong" if sharded and world > 1 else "single-gpu", "files": int(tot[6]), "blocks": int(tot[1]),
           "samples": int(tot[0]), "kernel_ms": k_ms, "value": tot[0] * args.config_steps / elapsed, "unit": UNIT,
           "algorithmic_bytes": algo, "achieved_gbs": algo / (k_ms * 1e-3) / 1e9 / (world if sharded else 1),
           "validated": bool(tot[5] == allr.shape[0] and tot[4] == 0), "flagged_blocks": int(tot[4])}
    out["frac"] = out["achieved_gbs"] / peak
    if sharded and world > 1:
        out["per_rank"] = [{"files": int(r[6]), "blocks": int(r[1]), "samples": int(r[0]), "kernel_ms": r[7] / 1e3} for r in allr]
    return out


# --------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = host_cores()
    nfiles = min(args.files, max(threads * 4, 64))
    corpus = build_corpus(nfiles, args.seconds, 0x5EED0000, threads, args.gen_budget_s, pin=False)
    n = size_cpu_sample(corpus, threads, args.cpu_baseline_s / max(1, (args.steps + args.warmup)))
    n = max(threads, min(n, nfiles))
    for _ in range(args.warmup):
        cpu_decode_sample(corpus, n, threads)
    t_total, s_total = 0.0, 0
    for _ in range(args.steps):
        s, dt = cpu_decode_sample(corpus, n, threads)
        t_total += dt
        s_total += s
    value = s_total / t_total
    sample = "%d of the workload's files (%d x %.0f s 16-bit stereo) per step, one file per thread" % (n, args.files, args.seconds)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": workload_config(args, corpus, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C# reference cannot run here (no .NET runtime); this is its C restatement (oracle/refdec.c, -O2), decode + WavpackFormatSamples, inputs in memory",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, corpus, nblocks):
    cfg = {
        "workload": "BASELINE configs[1]: batch of %d synthetic 16-bit stereo 44.1 kHz default-mode .wv files x %.0f s per GPU, one WavPack block per thread" % (args.files, args.seconds),
        "files_per_gpu": args.files, "seconds_per_file": args.seconds, "block_samples": 22050,
        "files_per_gpu_requested": getattr(args, "requested_files", args.files),
        "unique_files_per_gpu": corpus["unique"] if corpus else None,
        "l2_policy": "inputs (compressed slab + PCM output, GBs) far larger than the 126 MB L2; no flush needed",
        "chunk_samples": 4096,
    }
    if nblocks is not None:
        cfg["blocks_per_gpu"] = nblocks
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--files", type=int, default=10000, help="files per GPU")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--gen-budget-s", type=float, default=60.0, help="time budget for encoding unique files")
    ap.add_argument("--cpu-baseline-s", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` object (the other BASELINE configs)")
    ap.add_argument("--only-configs", default="", help="comma-separated substrings: run only the matching entries of `configs`")
    ap.add_argument("--config-steps", type=int, default=3)
    ap.add_argument("--config-gen-budget-s", type=float, default=6.0)
    args = ap.parse_args()
    if args.warmup < 3:
        log("note: W < 3 requested; the timing rules want >= 3 warm-up steps")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from wavpackdecoder_b200 import _native as N
    from wavpackdecoder_b200.batch import BatchDecoder, Corpus

    lib = N.load()
    if lib.wvb_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; libwvb has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output (it defaults to stdout) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    threads = max(1, host_cores() // max(1, world))
    # host RAM guard: the e2e leg pins the compressed slab and the PCM output of the whole per-GPU batch (~2.9 MB per 10 s file)
    requested_files = args.files
    args.requested_files = requested_files
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        per_file = int(args.seconds * 44100 * 4 * 1.55) + 65536
        cap = int(0.7 * avail / max(1, world) / per_file)
        if cap < args.files:
            args.files = max(64, cap)
            log("host RAM guard: files per GPU reduced %d -> %d (MemAvailable %.0f GB, %d ranks)" % (requested_files, args.files, avail / 1e9, world))
    except Exception:
        pass
    t0 = time.perf_counter()
    corpus = build_corpus(args.files, args.seconds, 0x5EED0000 + rank * args.files, threads, args.gen_budget_s, pin=True)
    log("corpus: %d files (%d unique) %.2f GB compressed, generated in %.1f s (+%.1f s total) on %d threads" % (
        args.files, corpus["unique"], corpus["compressed_bytes"] / 1e9, corpus["gen_s"], time.perf_counter() - t0, threads))

    slab = corpus["slab"]
    t0 = time.perf_counter()
    cp = Corpus(slab, corpus["offsets"], corpus["sizes"], out_format=N.OUT_PCM, threads=threads)
    index_s = time.perf_counter() - t0
    total_samples = cp.total_samples
    pcm_bytes = cp.out_bytes
    log("index: %d blocks, %d samples, %.2f GB PCM, %.3f s" % (cp.nblocks, total_samples, pcm_bytes / 1e9, index_s))
    assert all(cp.infos[i].status == 0 for i in range(cp.nfiles))

    dec = BatchDecoder(local_rank)
    dev = torch.device("cuda", local_rank)
    d_in = torch.empty(slab.size, dtype=torch.uint8, device=dev)
    d_in.copy_(corpus["slab_t"][:slab.size], non_blocking=False)
    d_out = torch.empty(pcm_bytes + 64, dtype=torch.uint8, device=dev)
    d_res = torch.empty(max(cp.nblocks, 1) * 16, dtype=torch.uint8, device=dev)
    dec.prepare(cp.descs, cp.nblocks, N.OUT_PCM)
    FL = N.IN_DEVICE | N.OUT_DEVICE | N.RESULTS_DEVICE

    def step_resident():
        dec.decode(d_in.data_ptr(), slab.size, None, cp.nblocks, d_out.data_ptr(), pcm_bytes, N.OUT_PCM, FL, d_res.data_ptr())

    for _ in range(max(args.warmup, 1)):
        step_resident()
    # validation (untimed): no block may report a CRC error; three files are compared with the oracle byte for byte
    res_host = d_res.cpu().numpy().view(np.uint32).reshape(-1, 4)
    crc_errors = int((res_host[:cp.nblocks, 1] & 1).sum())
    flagged = int((res_host[:cp.nblocks, 1] != 0).sum())
    import _harness as H
    validated = crc_errors == 0 and flagged == 0
    for i in sorted({0, cp.nfiles // 2, cp.nfiles - 1}):
        o = int(cp.file_out_offset[i])
        nb = int(cp.infos[i].indexed_samples) * 4
        got = d_out[o:o + nb].cpu().numpy()
        data = slab[int(corpus["offsets"][i]):int(corpus["offsets"][i]) + int(corpus["sizes"][i])].tobytes()
        ref, errs, status, info = H.oracle_decode(data)
        validated = validated and status == 0 and errs == 0 and np.array_equal(got, H.format_samples(ref, 2))
    log("validation vs oracle: %s (crc_errors=%d flagged=%d)" % (validated, crc_errors, flagged))

    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    kernel_ms, launches = [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
        tm = dec.timing()
        kernel_ms.append(tm["kernel_ms"])
        launches += tm["launches"]
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    clk = clocks.stop()
    from wavpackdecoder_b200.sharding import max_over_ranks
    elapsed_max = max_over_ranks(elapsed, dev)
    value = total_samples * world * args.steps / elapsed_max

    # ---- e2e: host buffers through the public C-ABI call, index pass + H2D + kernels + D2H every step ----
    e2e = None
    if not args.no_e2e:
        out_t = torch.empty(pcm_bytes + 64, dtype=torch.uint8, pin_memory=True)
        out_np = out_t.numpy()

        def step_e2e():
            # One library call: wvb_batch_decode_files indexes the files on the host threads WHILE the segments already indexed
            # are uploaded, decoded and downloaded.  (cap_hint: a caller that decodes batch after batch sizes its block table
            # from the previous one.)  WVB_BENCH_TWO_CALLS=1: the round-1 flow, wvb_index_many then wvb_batch_decode.
            if os.environ.get("WVB_BENCH_TWO_CALLS"):
                res = (N.BlockResult * max(cp.nblocks, 1))()
                c2 = Corpus(slab, corpus["offsets"], corpus["sizes"], out_format=N.OUT_PCM, threads=threads, cap_hint=cp.nblocks)
                dec.decode(slab.ctypes.data, slab.size, c2.descs, c2.nblocks, out_np.ctypes.data, pcm_bytes, N.OUT_PCM, 0, res)
                return c2, res
            return dec.decode_slab(slab, corpus["offsets"], corpus["sizes"], out_np, cap_hint=cp.nblocks, out_format=N.OUT_PCM, threads=threads,
                                   out_cap=pcm_bytes)

        for _ in range(max(args.warmup, 1)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            c2, results = step_e2e()
        torch.cuda.synchronize()
        e_elapsed = time.perf_counter() - t0
        e_elapsed_max = max_over_ranks(e_elapsed, dev)
        # validation of what came back over PCIe: every block's result clean, and the PCM bytes of three files compared with
        # the oracle's decode of the same .wv bytes
        rt = N.result_table(results, cp.nblocks)
        e2e_ok = bool((rt["rflags"] == 0).all()) and c2.nblocks == cp.nblocks and np.array_equal(c2.file_out_offset, cp.file_out_offset)
        for i in sorted({0, cp.nfiles // 2, cp.nfiles - 1}):
            o = int(c2.file_out_offset[i])
            data = slab[int(corpus["offsets"][i]):int(corpus["offsets"][i]) + int(corpus["sizes"][i])].tobytes()
            ref, errs, status, info = H.oracle_decode(data)
            want = H.format_samples(ref, 2)
            e2e_ok = e2e_ok and status == 0 and errs == 0 and np.array_equal(out_np[o:o + want.size], want)
        e2e_value = total_samples * world * args.steps / e_elapsed_max
        e2e = {"value": e2e_value, "unit": UNIT,
               "h2d_bytes_per_step": int(slab.size + cp.nblocks * (160 + 4)), "d2h_bytes_per_step": int(pcm_bytes + cp.nblocks * 16),
               "includes": "host block-index pass + H2D + kernels + D2H, one wvb_batch_decode_files call per step (index pass overlapped with the copies)",
               "validated": bool(e2e_ok), "ms_per_step": 1000.0 * e_elapsed_max / args.steps}

    # ---- verify-only end to end: host .wv bytes in, PCM stays in HBM, MD5 per file computed on the device,
    # only 16 B per file and the per-block results come back.  Extra to the contract's `e2e`; same timing rules. ----
    e2e_verify = None
    if not args.no_e2e:
        import hashlib
        d_out = torch.empty(pcm_bytes + 64, dtype=torch.uint8, device=dev)

        def step_verify():
            c2, _res = dec.decode_slab(slab, corpus["offsets"], corpus["sizes"], None, cap_hint=cp.nblocks, out_format=N.OUT_PCM, threads=threads,
                                       mem_flags=N.OUT_DEVICE, out_ptr=d_out.data_ptr(), out_cap=pcm_bytes)
            lens = np.array([int(c2.infos[i].indexed_samples) * 4 for i in range(c2.nfiles)], dtype=np.uint64)
            return c2, dec.md5_ranges(c2.file_out_offset, lens, pcm_bytes, d_out.data_ptr())

        for _ in range(max(args.warmup, 1)):
            step_verify()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            c2, digests = step_verify()
        torch.cuda.synchronize()
        v_elapsed = max_over_ranks(time.perf_counter() - t0, dev)
        # spot check against hashlib over the PCM the e2e leg brought back (same decode, host copy)
        v_ok = True
        for i in (0, c2.nfiles // 2, c2.nfiles - 1):
            o = int(c2.file_out_offset[i])
            ln = int(c2.infos[i].indexed_samples) * 4
            v_ok = v_ok and hashlib.md5(out_np[o:o + ln].tobytes()).digest() == digests[i].tobytes()
        e2e_verify = {"value": total_samples * world * args.steps / v_elapsed, "unit": UNIT,
                      "h2d_bytes_per_step": int(slab.size + cp.nblocks * (160 + 4) + c2.nfiles * 16),
                      "d2h_bytes_per_step": int(c2.nfiles * 16 + cp.nblocks * 16),
                      "includes": "host block-index pass + H2D + kernels + device MD5 per file; PCM never leaves HBM", "validated": bool(v_ok)}
        del d_out

    if e2e is not None:
        # the e2e leg's own roofline: its bytes over PCIe with no decode (this overwrites the pinned output buffer: last use)
        ceil = pcie_ceiling(corpus["slab_t"], slab.size, out_t, pcm_bytes, dev, barrier, max_over_ranks)
        ceil_value = total_samples * world / ceil["both_s"]
        ceil.update({"value": ceil_value, "unit": UNIT,
                     "what": "H2D of the compressed slab and D2H of the PCM at once, same pinned buffers, no decode, all ranks at once, best of 3"})
        e2e["pcie_ceiling"] = ceil
        e2e["frac_of_ceiling"] = e2e["value"] / ceil_value

    # ---- roofline of the dominant kernel ----
    peak, peak_kind = measured_peak()
    k_ms = float(np.mean(kernel_ms))
    algo_bytes = corpus["compressed_bytes"] + pcm_bytes
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if traffic and traffic.get("blocks_per_launch") == cp.nblocks else None,
                "traffic_source": traffic.get("source") if traffic else None, "peak_kind": peak_kind,
                "kernel": "k_decode_pcm<stereo,lossless,FixedDecorr<-2,3,2,18,18>,F16>", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": int(algo_bytes),
                "bytes_per_sample": algo_bytes / total_samples}
    # What actually binds the kernel (explanatory, the contract's bound stays "hbm"): warp-instruction issue.  Instructions per
    # launch come from the same ncu capture as `traffic`; the rate is this run's, against SMs x 4 schedulers x the sampled clock.
    if traffic and traffic.get("blocks_per_launch") == cp.nblocks and traffic.get("warp_instructions_per_launch") and clk.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = sms * 4 * clk["sm_mhz"] * 1e6
        issue_rate = traffic["warp_instructions_per_launch"] / (k_ms * 1e-3)
        roofline["issue"] = {"warp_instructions_per_launch": traffic["warp_instructions_per_launch"], "achieved": issue_rate, "peak": issue_peak,
                             "unit": "warp-inst/s", "frac": issue_rate / issue_peak,
                             "warp_instructions_per_32_samples": traffic["warp_instructions_per_launch"] / total_samples * 32}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:  # (the other ranks wait at the next barrier)
        cores = host_cores()
        n = size_cpu_sample(corpus, cores, args.cpu_baseline_s if world == 1 else args.cpu_baseline_s / 2)
        s, dt = cpu_decode_sample(corpus, n, cores)
        cpu_baseline = {"value": s / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d of the %d files, one file per thread, decode + WavpackFormatSamples, %.1f s" % (n, args.files, dt)}
    barrier()

    # ---- the other BASELINE configs ----
    configs = None
    if not args.no_configs:
        # release the headline batch's memory first
        d_in = d_out = d_res = out_t = out_np = None  # noqa: F841
        dec.close()
        corpus.pop("slab_t", None)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        configs = {}
        wanted = [x for x in args.only_configs.split(",") if x]
        for name, bcfg, kw, oflags, nfiles, secs, sharded in EXTRA_CONFIGS:
            if wanted and not any(x in name for x in wanted):
                continue
            if world > 1 and not sharded:
                continue  # at N > 1 only the configs BASELINE.json shards (3 and 5) run, as one corpus cut over the ranks
            t0 = time.perf_counter()
            try:
                r = run_extra_config(name, kw, oflags, nfiles * (world if sharded else 1), secs, sharded, args, rank, world, dev, threads, peak, barrier)
                r["baseline_config"] = bcfg
            except Exception as e:  # one failing config must not take the headline line with it
                r = {"baseline_config": bcfg, "error": repr(e)[:300]}
                barrier()
            configs[name] = r
            log("%s: %s (%.1f s)" % (name, {k: v for k, v in r.items() if k != "per_rank"}, time.perf_counter() - t0))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * elapsed_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic (in-repo encoder, %d unique files per GPU, validated vs oracle: %s)" % (corpus["unique"], validated),
            "config": workload_config(args, corpus, cp.nblocks), "clocks": clk, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e_verify": e2e_verify, "configs": configs,
            "pcm_gb_per_s": pcm_bytes * world * args.steps / elapsed_max / 1e9, "validated": bool(validated),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
