#!/usr/bin/env python3
"""Copy-only ceiling of the end-to-end path: the bytes one bench step moves over PCIe, with no decode.

    python tools/pcie_ceiling.py [--h2d-gb 8.2 --d2h-gb 17.64 --reps 3]
    torchrun --nproc-per-node N tools/pcie_ceiling.py ...      # all ranks copy at once (whole-box ceiling)

Times, with CUDA events on their own streams: H2D alone, D2H alone, and both directions at once (what a pipelined
decode step does).  Also prints the host topology facts that decide where the ceiling comes from (NUMA nodes, the
cores this process may use, `nvidia-smi topo -m`).  bench.py measures the same thing inline (`e2e.pcie_ceiling`);
this tool exists for experiments (write-combined input slab, NUMA placement).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def topo():
    out = {}
    try:
        out["cores"] = len(os.sched_getaffinity(0))
    except Exception:
        out["cores"] = os.cpu_count()
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        out["numa_nodes"] = len(nodes)
        out["numa_cpulist"] = {n: open("/sys/devices/system/node/%s/cpulist" % n).read().strip() for n in sorted(nodes)}
    except Exception:
        out["numa_nodes"] = None
    try:
        out["memavail_gb"] = [int(l.split()[1]) / 1e6 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
    except Exception:
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h2d-gb", type=float, default=8.2)
    ap.add_argument("--d2h-gb", type=float, default=17.64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--wc", action="store_true", help="write-combined pinned memory for the H2D source")
    ap.add_argument("--topo", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    nin, nout = int(args.h2d_gb * 1e9), int(args.d2h_gb * 1e9)
    rt = C.CDLL("libcudart.so.12") if args.wc else None
    if args.wc:
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nin), 4) == 0  # cudaHostAllocWriteCombined
        import numpy as np
        h_in = torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * nin).from_address(p.value)))
    else:
        h_in = torch.empty(nin, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nout, dtype=torch.uint8, pin_memory=True)
    h_in[::4096] = 1
    h_out[::4096] = 1
    d_in = torch.empty(nin, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nout, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(do_in, do_out):
        best = None
        for _ in range(args.reps):
            barrier()
            t0 = time.perf_counter()
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            best = dt if best is None else min(best, dt)
        return best

    timed(True, True)
    t_in, t_out, t_both = timed(True, False), timed(False, True), timed(True, True)
    if rank == 0:
        line = {"n_gpus": world, "h2d_gb": args.h2d_gb, "d2h_gb": args.d2h_gb, "wc": args.wc,
                "h2d_alone_gbs_per_gpu": nin / t_in / 1e9, "d2h_alone_gbs_per_gpu": nout / t_out / 1e9,
                "both_s": t_both, "both_h2d_gbs_per_gpu": nin / t_both / 1e9, "both_d2h_gbs_per_gpu": nout / t_both / 1e9,
                "box_d2h_gbs_both": nout * world / t_both / 1e9}
        if args.topo:
            line["topo"] = topo()
            try:
                line["nvidia_smi_topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            except Exception as e:
                line["nvidia_smi_topo"] = str(e)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
