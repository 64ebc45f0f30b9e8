#!/usr/bin/env python
"""bench.py -- k-mers/s of the repart->superk->count->merge hot path on synthetic FASTQ.

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line (rank 0).
A "step" is one pass of the whole hot path over one batch of synthetic input:
  workload cfg2 (BASELINE.json configs[1]): S samples x R reads x 150 nt, k=31, hash:bf:bin,
  P partitions, Bloom size B  ->  per-partition dense Bloom slabs (the .cmbf bodies).
`value`  : k-mer occurrences / s with the FASTQ text already resident in HBM (device timing).
`e2e`    : same metric through the host-buffer C-ABI path: FASTQ in pinned host memory,
           H2D copies and the D2H read of every .cmbf body inside the timed region.
`--impl reference` times the unmodified reference CPU pipeline (oracle/_ref/bin/kmtricks) on a
bounded sample of the same workload with all host threads.
Inputs are far larger than L2 (315 MB of text per sample launch vs 126 MB L2).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# several lanes = several streams and NCCL communicators per GPU: give every stream its own hardware queue, so that a
# collective kernel waiting for its peers never sits in front of another lane's kernel (must be set before CUDA starts)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

# survey §8(d) "Algorithmic bytes": S1 = read FASTQ + write buckets at the reference's compact
# super-k-mer format (1.03 B per k-mer, measured on the reference's skp files)
BUCKET_BYTES_PER_KMER = 1.03


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kmx", choices=["kmx", "reference"])
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--partitions", type=int, default=64)
    ap.add_argument("--bloom-size", type=int, default=200_000_000)
    ap.add_argument("--hard-min", type=int, default=2)
    ap.add_argument("--kmer-size", type=int, default=31)
    ap.add_argument("--mode", default="hash:bf:bin")
    ap.add_argument("--lanes", type=int, default=4, help="samples in flight (kmx_run_samples lanes)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-samples", type=int, default=0, help="samples in the bounded CPU-reference run (0 = one per host thread, at most 16: "
                                                               "the reference parses a sample on ONE thread, fewer samples leave cores idle)")
    ap.add_argument("--ref-reads", type=int, default=0, help="reads per sample of the CPU-reference run (0 = the workload's own, scaled down only "
                                                             "when steps+warmup would push the run past --ref-budget-s)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    ap.add_argument("--hang-timeout-s", type=float, default=150.0,
                    help="N > 1: if warm-up + timed steps have not finished after this many seconds, every rank re-executes the "
                         "bench with ONE lane (one stream and communicator per GPU) instead of never printing a line")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the oracle-checked pass after timing")
    ap.add_argument("--other-configs", default="auto", choices=["auto", "on", "off"],
                    help="BASELINE configs 3-5 (1000 x 5M kmer:count, 1000 x 5M hash:bft, 500 x 5M k=63 kmer:pa + rescue) after the main "
                         "measurement, reported under other_configs; auto = only on 8 GPUs, where they fit at full size")
    ap.add_argument("--oc-child", default="", help=argparse.SUPPRESS)      # internal: run ONE other config in this (child) process
    ap.add_argument("--oc-timeout-s", type=float, default=120.0, help="other_configs: hard limit per config (each runs in child processes)")
    ap.add_argument("--oc-samples-scale", type=float, default=1.0, help="other_configs: fraction of the samples (testing on fewer GPUs)")
    ap.add_argument("--oc-reads-scale", type=float, default=1.0, help="other_configs: fraction of the reads per sample")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        self.mode = os.environ.get("KMX_BENCH_CLOCKS", "nvml")
        if self.mode == "none":
            return
        if self.mode == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                self.nv = pynvml
                self.hd = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
                self.stop_flag = False
                self.t = threading.Thread(target=self._poll, daemon=True)
                self.t.start()
                return
            except Exception:
                self.mode = "smi"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        R = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
             "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        period = float(os.environ.get("KMX_BENCH_CLOCKS_MS", "200")) / 1e3
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.hd, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.hd, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.hd) / 1e3
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.hd)
                self.rows.append([str(sm), str(mx), str(pw)] + ["Active" if rs & v else "Not Active" for v in R.values()])
            except Exception:
                pass
            time.sleep(period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.mode == "nvml":
            self.stop_flag = True
            self.t.join(timeout=2)
        elif not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.mode}


def workload_name(args, world=1):
    return (f"cfg2: {args.samples} samples x {args.reads} reads x {args.read_len} nt, k={args.kmer_size}, {args.mode}, "
            f"P={args.partitions}, bloom={args.bloom_size}, hard-min {args.hard_min}, --static-repart, m=10"
            + (f"; samples per GPU (x{world} GPUs = {args.samples * world} columns), partitions sharded over GPUs, "
               f"one NCCL all-to-all-v of bucket regions per sample" if world > 1 else ""))


def n_kmers(args):
    return args.samples * args.reads * (args.read_len - args.kmer_size + 1)


# --------------------------------------------------------------------------- reference arm
def size_reference_run(args, nruns):
    """Fills args.ref_samples / args.ref_reads: the workload's own per-sample shape when the whole run (nruns pipeline runs)
    fits --ref-budget-s at ~8e7 k-mers/s, else fewer reads per sample (stated in the line)."""
    threads = os.cpu_count() or 1
    if args.ref_samples <= 0:
        args.ref_samples = max(2, min(16, threads, args.samples))
    if args.ref_reads <= 0:
        per_read = args.read_len - args.kmer_size + 1
        fit = args.ref_budget_s / max(nruns, 1) * 8.0e7 / (per_read * args.ref_samples)
        args.ref_reads = int(max(100_000, min(args.reads, fit // 50_000 * 50_000)))
    return args.ref_samples * args.ref_reads == args.samples * args.reads


def run_reference_once(args, workdir, threads):
    """One bounded run of the unmodified reference pipeline; returns (seconds, kmers)."""
    from oracle import oracle as O
    fof = os.path.join(workdir, "fof.txt")
    rd = os.path.join(workdir, "run")
    shutil.rmtree(rd, ignore_errors=True)
    kind, what = args.mode.split(":")[:2]
    # same per-partition window as the full workload: bloom scaled to keep W
    cmd = [O.timed_ref_bin()[0], "pipeline", "--file", fof, "--run-dir", rd, "--kmer-size", str(args.kmer_size),
           "--mode", f"{kind}:{what}:bin", "--hard-min", str(args.hard_min), "--nb-partitions", str(args.partitions),
           "--minimizer-size", "10", "--static-repart", "--bloom-size", str(args.bloom_size), "-t", str(threads)]
    t0 = time.perf_counter()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    shutil.rmtree(rd, ignore_errors=True)
    return dt, args.ref_samples * args.ref_reads * (args.read_len - args.kmer_size + 1)


def _write_ref_sample(job):
    from kmtricks_b200 import synth
    path, s, reads, read_len, genome = job
    with open(path, "wb") as g:
        step = 50_000
        for r0 in range(0, reads, step):
            g.write(synth.make_fastq(1234, s, min(step, reads - r0), L=read_len, G=genome, d=2e-3, e=2e-3, revcomp=True, first_read=r0))
    return path


def make_ref_inputs(args, workdir):
    """Same generator as the GPU arm (numpy twin of kmx_synth_fastq), one worker process per sample."""
    import multiprocessing as mp
    jobs = [(os.path.join(workdir, f"S{s}.fastq"), s, args.ref_reads, args.read_len, args.genome) for s in range(args.ref_samples)]
    with mp.get_context("spawn").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        pool.map(_write_ref_sample, jobs)
    with open(os.path.join(workdir, "fof.txt"), "w") as f:
        for s, j in enumerate(jobs):
            f.write(f"S{s}: {j[0]}\n")


def ref_workdir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix="kmx_ref_", dir=base)


def cpu_baseline(args):
    from oracle import oracle as O
    if not O.have_ref():
        return {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/bin/kmtricks missing"}
    wd = ref_workdir()
    try:
        size_reference_run(args, 2)
        make_ref_inputs(args, wd)
        threads = os.cpu_count() or 1
        dt, km = run_reference_once(args, wd, threads)
        return {"value": km / dt, "unit": "k-mers/s", "cores": threads, "kind": "reference",
                "sample": f"{args.ref_samples} samples x {args.ref_reads} reads x {args.read_len} nt of the same generator, "
                          f"kmtricks pipeline --mode {args.mode} -t {threads} on {wd.split('/')[1]}, {dt:.2f} s wall incl. file I/O; "
                          f"binary built {O.timed_ref_bin()[1]}"}
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    same = size_reference_run(args, args.steps + args.warmup)
    cfg = {"workload": workload_name(args, int(os.environ.get("WORLD_SIZE", "1"))),
           "sample": f"each step = {args.ref_samples} samples x {args.ref_reads} reads x {args.read_len} nt of that workload's generator "
                     f"(same k / P / bloom / hard-min / mode; one sample per host thread keeps every core busy; bounded so that "
                     f"{args.steps}+{args.warmup} runs end within ~{int(args.ref_budget_s)} s)",
           "same_config": bool(same), "same_per_sample_shape": args.ref_reads == args.reads,
           "reference_build": O.timed_ref_bin()[1], "host_threads": os.cpu_count() or 1,
           "l2": "inputs larger than L2"}
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/kmtricks not built (run oracle/build_ref.sh where /root/reference exists)"}))
        return
    wd = ref_workdir()
    try:
        make_ref_inputs(args, wd)
        threads = os.cpu_count() or 1
        for _ in range(args.warmup):
            run_reference_once(args, wd, threads)
        tot = 0.0; km = 0
        for _ in range(args.steps):
            dt, k1 = run_reference_once(args, wd, threads)
            tot += dt; km += k1
        val = km / tot
        line = {"impl": "reference", "metric": "k-mers/s end-to-end (repart->merge)", "value": val, "unit": "k-mers/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": threads, "kind": "reference",
                                 "sample": f"{args.ref_samples} x {args.ref_reads} reads per step, files on {wd.split('/')[1]}, binary built {O.timed_ref_bin()[1]}"},
                "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
    finally:
        shutil.rmtree(wd, ignore_errors=True)



# --------------------------------------------------------------------------- BASELINE configs 3-5
OTHER_CONFIGS = {
    # SURVEY 8(d) "Synthetic inputs": generator parameters that bound the number of matrix rows
    "cfg3": dict(what="1000 samples x 5M reads x 150 nt, k=31, kmer:count:bin, P=512, hard-min 3", samples=1000, reads=5_000_000, k=31,
                 mode="kmer:count:bin", P=512, hard_min=3, genome=10_000_000, d=1e-5, e=1e-3,
                 twin="tests/test_gpu_atscale.py::test_cfg345_twins_equal_reference_binary[cfg3_kmer_count_P512]"),
    "cfg4": dict(what="1000 samples x 5M reads x 150 nt, k=31, hash:bft:bin (Bloom rows + bit transpose), P=512, bloom 4e8, hard-min 3", samples=1000,
                 reads=5_000_000, k=31, mode="hash:bft:bin", P=512, hard_min=3, genome=10_000_000, d=1e-5, e=1e-3, bloom=400_000_000,
                 twin="tests/test_gpu_atscale.py::test_cfg345_twins_equal_reference_binary[cfg4_hash_bft_P512]"),
    "cfg5": dict(what="500 samples x 5M reads x 150 nt, k=63, kmer:pa:bin, P=256, hard-min 1, soft-min 3, share-min 2 (rescue)", samples=500,
                 reads=5_000_000, k=63, mode="kmer:pa:bin", P=256, hard_min=1, soft_min=3, share_min=2, recurrence_min=1, genome=10_000_000,
                 d=1e-4, e=2e-4, twin="tests/test_gpu_atscale.py::test_cfg345_twins_equal_reference_binary[cfg5_k63_kmer_pa_rescue_P256]"),
}


def run_other_config(name, c, args, world, rank, local, peak):
    """One pass of the whole hot path over a BASELINE config that does not fit as text: the FASTQ of a batch of samples is
    generated on the device (untimed), the batch goes through stage 1 / exchange / stage 2 (timed, CUDA events), the lists
    stay resident; then every rank merges its partitions (timed).  Returns the line's entry for this config."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from kmtricks_b200 import _lib, engine, synth
    S = max(world, int(round(c["samples"] * args.oc_samples_scale)))
    R = max(1000, int(round(c["reads"] * args.oc_reads_scale)))
    Lr, k, P = 150, c["k"], c["P"]
    cfg = engine.Config(kmer_size=k, nb_partitions=P, mode=c["mode"], hard_min=c["hard_min"], soft_min=c.get("soft_min", 1),
                        recurrence_min=c.get("recurrence_min", 1), share_min=c.get("share_min", 0), bloom_size=c.get("bloom", 10_000_000))
    n_local = (S + world - 1) // world
    eng = engine.Engine(cfg, S, device=local)
    L, h = eng.lib, eng.h

    def ck(rc, what):
        if rc:
            raise RuntimeError(f"{name} {what}: {L.kmx_last_error(h).decode()} ({rc})")
    try:
        if world > 1:
            from kmtricks_b200 import dist as kd
            kd.init_engine(eng, nlanes=args.lanes)
            my_parts = list(kd.owned_partitions(P, world, rank))
        else:
            my_parts = list(range(P))
        sb = R * synth.record_bytes(Lr)
        B = max(1, min(8, n_local))
        d_text = C.c_void_p()
        ck(L.kmx_dev_alloc(h, B * sb + 64, C.byref(d_text)), "dev_alloc")
        stream = torch.cuda.ExternalStream(L.kmx_stream(h), device=torch.device("cuda", local))

        def timed(fn):
            ck(L.kmx_sync(h), "sync")
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            ck(L.kmx_sync(h), "sync"); torch.cuda.synchronize()
            return e0.elapsed_time(e1)

        def batch(b0, timed_run):
            nb = min(B, n_local - b0)
            sizes = []
            for i in range(nb):
                slot = rank * n_local + b0 + i
                if slot < S:
                    ck(L.kmx_synth_fastq(h, 1234, slot, 0, R, Lr, c["genome"], c["d"], c["e"], 1, d_text.value + i * sb), "synth")
                    sizes.append(sb)
                else:
                    sizes.append(0)                      # padding sample of the last rank
            ck(L.kmx_sync(h), "sync")
            ptrs = (C.c_void_p * nb)(*[d_text.value + i * sb for i in range(nb)])
            sz = (C.c_size_t * nb)(*sizes)
            hm = (C.c_uint32 * nb)(*([c["hard_min"]] * nb))
            if world > 1:
                fn = lambda: ck(L.kmx_dist_run_batch(h, nb, ptrs, sz, 1, hm, b0, n_local, None), "dist_run_batch")
            else:
                ids = (C.c_uint32 * nb)(*[min(b0 + i, S - 1) for i in range(nb)])
                fn = lambda: ck(L.kmx_run_samples(h, nb, ptrs, sz, 1, ids, hm, args.lanes, None), "run_samples")
            if timed_run:
                return timed(fn)
            fn()
            return 0.0

        t_start = time.perf_counter()

        def log(msg):
            if rank == 0:
                print(f"[{name} {time.perf_counter() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)
        log(f"S={S} R={R} world={world} n_local={n_local} B={B}")
        batch(0, False)                                  # warm-up: allocations, bucket geometry, table sizes
        ck(L.kmx_reset(h), "reset")
        log("warm-up batch done")
        if world > 1:
            dist.barrier()
        ms_sc = 0.0
        for b0 in range(0, n_local, B):
            ms_sc += batch(b0, True)
            log(f"batch {b0 // B + 1}/{(n_local + B - 1) // B}: {ms_sc:.0f} ms so far, device bytes {int(L.kmx_device_bytes(h)) >> 20} MiB")
        soft = np.full(S, cfg.soft_min, dtype=np.uint32)
        mp = _lib.KmxMergeParams(soft.ctypes.data_as(C.POINTER(C.c_uint32)), cfg.recurrence_min, cfg.share_min,
                                 {"count": 0, "pa": 1, "bf": 2, "bft": 3}[cfg.fmt], 0)
        res = _lib.KmxMergeResult()
        body = [0, 0]

        def merges():
            for p in my_parts:
                ck(L.kmx_merge_partition(h, p, C.byref(mp), C.byref(res)), "merge")
                body[0] += res.n_rows * res.row_bytes; body[1] += res.n_rows
        ms_merge = timed(merges)
        log(f"merges done: {ms_merge:.0f} ms")
        D = 0
        nsz = C.c_uint64()
        for s_ in range(S):
            for p_ in my_parts:
                L.kmx_counts_size(h, s_, p_, C.byref(nsz)); D += nsz.value
        tot = torch.tensor([ms_sc + ms_merge, ms_sc, ms_merge], device="cuda", dtype=torch.float64)
        sums = torch.tensor([float(D), float(body[0]), float(body[1])], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        ms, ms_sc, ms_merge = [float(x) for x in tot.tolist()]
        D, body_b, rows = [float(x) for x in sums.tolist()]
        kmers = S * R * (Lr - k + 1)
        w = (k + 31) // 32
        key_b = 8 if cfg.key_kind == "hash" else 8 * w
        bucket = BUCKET_BYTES_PER_KMER if k <= 32 else 1.17     # reference record density at k=63: (1 + ceil((63 + l - 1)/4)) / l, l ~ 18
        alg = S * sb + 2 * bucket * kmers + 2 * (key_b + 4) * D + body_b * (2 if cfg.fmt == "bft" else 1)
        return {"workload": c["what"], "value": kmers / (ms * 1e-3), "unit": "k-mers/s", "n_gpus": world, "ms_per_step": ms,
                "ms_superk_exchange_count": ms_sc, "ms_merge": ms_merge, "kmers_per_step": kmers,
                "scale": {"samples": S, "reads_per_sample": R, "full_size": S == c["samples"] and R == c["reads"]},
                "surviving_key_sample_pairs": int(D), "matrix_rows": int(rows), "matrix_bytes": int(body_b),
                "pipeline_roofline": {"algorithmic_bytes_per_step": alg, "achieved_GBps_per_gpu": alg / world / (ms * 1e-3) / 1e9,
                                      "frac": alg / world / (ms * 1e-3) / 1e9 / peak, "peak": peak},
                "steps": 1, "warmup": "one batch", "data": "synthetic, generated on the device per batch of <= 8 samples per GPU (untimed)",
                "parity_twin": c["twin"], "device_bytes": int(L.kmx_device_bytes(h))}
    finally:
        eng.close()


# --------------------------------------------------------------------------- kmx arm
def main_kmx(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from kmtricks_b200 import _lib, engine, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    t_begin = time.perf_counter()

    def log(msg):                       # progress on stderr (rank 0): a run that is cut off still says how far it got
        if rank == 0:
            print(f"[bench {time.perf_counter() - t_begin:6.1f}s] {msg}", file=sys.stderr, flush=True)

    # Safety net for N > 1 (first attempt only): the lanes of a rank drive independent NCCL communicators from their own host
    # threads; should that ever wedge (it did once at 8 GPUs before buffers stopped being freed mid-run), the ranks re-execute
    # themselves with one lane, where no cross-lane wait exists, and say so in the line.
    attempt = int(os.environ.get("KMX_BENCH_ATTEMPT", "1"))
    watchdog = None
    if world > 1 and attempt == 1 and args.lanes > 1 and args.hang_timeout_s > 0:
        def _rexec():
            log(f"no progress after {args.hang_timeout_s:.0f} s: re-executing with --lanes 1")
            env = dict(os.environ)
            env["KMX_BENCH_ATTEMPT"] = "2"
            env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + 57)
            env["TORCHELASTIC_USE_AGENT_STORE"] = "False"
            try:
                sys.stdout.flush(); sys.stderr.flush()
            except Exception:
                pass
            os.execve(sys.executable, [sys.executable, os.path.abspath(__file__)] + sys.argv[1:] + ["--lanes", "1"], env)
        watchdog = threading.Timer(args.hang_timeout_s, _rexec)
        watchdog.daemon = True
        watchdog.start()

    # multi-GPU: samples shard over ranks (weak scaling: every rank parses `samples` samples)
    cfg = engine.Config(kmer_size=args.kmer_size, nb_partitions=args.partitions, mode=args.mode, hard_min=args.hard_min,
                        bloom_size=args.bloom_size)
    N = args.samples                       # samples parsed by THIS rank (weak scaling)
    N_tot = N * world                      # sample columns of every matrix
    eng = engine.Engine(cfg, N_tot, device=local)
    L = eng.lib
    h = eng.h
    if world > 1:
        from kmtricks_b200 import dist as kd
        kd.init_engine(eng, nlanes=args.lanes)
        my_parts = list(kd.owned_partitions(args.partitions, world, rank))
    else:
        my_parts = list(range(args.partitions))
    rb = synth.record_bytes(args.read_len)
    sample_bytes = args.reads * rb
    kmers_step = n_kmers(args)
    P = args.partitions
    Wb = cfg.window_bits
    row_bytes = (N_tot + 7) // 8
    slab_bytes = Wb * row_bytes
    fmt = cfg.fmt

    def ck(rc, what):
        if rc:
            raise RuntimeError(f"{what}: {L.kmx_last_error(h).decode()} ({rc})")

    # ---- synthetic FASTQ resident in HBM (all samples)
    d_text = C.c_void_p()
    ck(L.kmx_dev_alloc(h, N * sample_bytes + 64, C.byref(d_text)), "dev_alloc text")
    for s in range(N):
        ck(L.kmx_synth_fastq(h, 1234, rank * N + s, 0, args.reads, args.read_len, args.genome, 2e-3, 2e-3, 1,
                             d_text.value + s * sample_bytes), "synth")
    ck(L.kmx_sync(h), "sync")
    # all P bodies stay in HBM
    d_out = C.c_void_p()
    ck(L.kmx_dev_alloc(h, len(my_parts) * (slab_bytes + 64) if fmt in ("bf", "bft") else 64, C.byref(d_out)), "dev_alloc out")
    soft = np.full(N_tot, cfg.soft_min, dtype=np.uint32)
    mp = _lib.KmxMergeParams(soft.ctypes.data_as(C.POINTER(C.c_uint32)), cfg.recurrence_min, cfg.share_min,
                             {"count": 0, "pa": 1, "bf": 2, "bft": 3}[fmt], 0)
    res = _lib.KmxMergeResult()

    wall = {}

    def tm(name, rc_fn, *a):
        t0 = time.perf_counter()
        rc = rc_fn(*a)
        wall[name] = wall.get(name, 0.0) + time.perf_counter() - t0
        ck(rc, name)

    dev_ptrs = (C.c_void_p * N)(*[d_text.value + s * sample_bytes for s in range(N)])
    sizes = (C.c_size_t * N)(*([sample_bytes] * N))
    hmins = (C.c_uint32 * N)(*([args.hard_min] * N))

    body_max = [0]; body_sum = [0]

    def merges(to_host=None):
        body_sum[0] = 0
        for j, p in enumerate(my_parts):
            if to_host is None and fmt in ("bf", "bft"):
                tm("set_out", L.kmx_set_merge_output, h, d_out.value + j * (slab_bytes + 64), slab_bytes + 64)
            tm("merge", L.kmx_merge_partition, h, p, C.byref(mp), C.byref(res))
            body_max[0] = max(body_max[0], res.n_rows * res.row_bytes); body_sum[0] += res.n_rows * res.row_bytes
            if to_host is not None:
                tm("merge_get", L.kmx_merge_get, h, to_host + (j * slab_bytes if fmt in ("bf", "bft") else 0), None, None)
        tm("set_out", L.kmx_set_merge_output, h, None, 0)

    def make_step(ptrs, on_device, lanes, to_host=None):
        def step():
            tm("reset", L.kmx_reset, h)
            if world > 1:
                tm("run_samples", L.kmx_dist_run_samples, h, N, ptrs, sizes, on_device, hmins, None)
            else:
                tm("run_samples", L.kmx_run_samples, h, N, ptrs, sizes, on_device, None, hmins, lanes, None)
            merges(to_host)
        return step

    step_device = make_step(dev_ptrs, 1, args.lanes)
    step_device_1lane = make_step(dev_ptrs, 1, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ck(L.kmx_sync(h), "sync")

    stream = torch.cuda.ExternalStream(L.kmx_stream(h), device=torch.device("cuda", local))

    def timed(fn, steps):
        barrier()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)          # lane 0's stream; every merge waits for all lanes first
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    log(f"inputs generated ({N} samples per GPU, {world} GPU(s))")
    for _ in range(args.warmup):
        step_device()
    log("warm-up done")
    launches0 = L.kmx_launch_count(h)
    wall.clear()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_device, args.steps)
    clocks = sampler.stop()
    host_wall = {k: round(1e3 * v / args.steps, 2) for k, v in wall.items()}
    launches = (L.kmx_launch_count(h) - launches0) // max(args.steps, 1)
    ms_step = ms / args.steps
    value = world * kmers_step / (ms_step * 1e-3)
    log(f"timed: {ms_step:.1f} ms/step")

    # per-kernel device time: same step, ONE lane (kernels back to back on one stream, so the
    # CUDA-event spans around each launch are exclusive), events recorded inside the library
    if world > 1:
        ck(L.kmx_dist_set_lanes(h, 1), "dist_set_lanes")      # one lane: samples back to back on one stream, exclusive spans
    step_device_1lane()
    ck(L.kmx_profile_enable(h, 1), "prof")
    ck(L.kmx_profile_reset(h), "prof")
    psteps = max(1, min(args.steps, 2))
    exch0 = int(L.kmx_stat(h, 4))
    ms_1lane = timed(step_device_1lane, psteps) / psteps
    exch_bytes_step = (int(L.kmx_stat(h, 4)) - exch0) // psteps
    prof = {}
    for i, name in enumerate(_lib.PROF_KINDS):
        tms = C.c_double(); cnt = C.c_uint64()
        ck(L.kmx_profile_get(h, i, C.byref(tms), C.byref(cnt)), "prof_get")
        if cnt.value:
            prof[name] = {"ms_per_step": tms.value / psteps, "launches_per_step": cnt.value // psteps}
    ck(L.kmx_profile_enable(h, 0), "prof")
    log(f"1-lane profile pass done: {ms_1lane:.1f} ms/step")
    if watchdog is not None:
        watchdog.cancel()
    if world > 1:
        ck(L.kmx_dist_set_lanes(h, args.lanes), "dist_set_lanes")
    exchange = None
    if world > 1 and "exchange" in prof:
        # bytes this rank put on NVLink per step (its buckets for the other ranks' partitions) over the CUDA-event time of the
        # grouped send/recv; 770 GB/s per direction is the measured peer-copy figure of B200_PROFILING.md (900 nominal)
        ems = prof["exchange"]["ms_per_step"]
        exchange = {"sent_bytes_per_step_rank0": exch_bytes_step, "ms_per_step": round(ems, 3), "GBps_per_direction": exch_bytes_step / (ems * 1e-3) / 1e9 if ems > 0 else None,
                    "nvlink_peer_copy_GBps_measured": 770.0, "frac_of_nvlink": exch_bytes_step / (ems * 1e-3) / 1e9 / 770.0 if ems > 0 else None,
                    "bytes_per_kmer": exch_bytes_step / kmers_step}

    # ---- roofline of the dominant kernel (device time share from the event spans)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kern_time = {k: v["ms_per_step"] for k, v in prof.items() if k not in ("fill", "exchange")}
    top = max(kern_time, key=kern_time.get) if kern_time else None
    roof = None
    if top:
        n_l = prof[top]["launches_per_step"]
        dur_ms = prof[top]["ms_per_step"] / n_l
        kmers_launch = kmers_step / N
        if top == "s1_superk":
            alg = sample_bytes + BUCKET_BYTES_PER_KMER * kmers_launch
            what = "S1: FASTQ text read + super-k-mer buckets written (1.03 B/k-mer)"
        elif top in ("hash_hist", "expand", "radix_sort", "rle", "hash_emit"):
            alg = BUCKET_BYTES_PER_KMER * kmers_launch
            what = "S2: buckets read (1.03 B/k-mer) [+12 B per surviving (key,sample)]; hash_hist = pass A (hash + bin), hash_emit = pass B (count + emit)"
        else:
            alg = body_sum[0] / max(len(my_parts), 1)
            what = "S3/S4: matrix body written"
        ach = alg / (dur_ms * 1e-3) / 1e9
        traffic = None
        try:      # dram__bytes_read+write per sample launch of this span's kernels, from the committed ncu --set full capture (same launch shape only)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["spans"]
            if top in tr and args.reads == 1_000_000 and args.read_len == 150 and args.mode == "hash:bf:bin" and args.partitions == 64 and args.bloom_size == 200_000_000:
                traffic = tr[top]["dram_read_bytes"] + tr[top]["dram_write_bytes"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg, "launch_ms": dur_ms, "what": what, "peak_source": peak_src}
        # whole hot path against the same roofline: SURVEY 8(d) algorithmic bytes of all four stages over the step time
        try:
            D = 0
            nsz = C.c_uint64()
            for s_ in range(N_tot):
                for p_ in my_parts:
                    ck(L.kmx_counts_size(h, s_, p_, C.byref(nsz)), "counts_size")
                    D += nsz.value
            key_b = 8 * ((args.kmer_size + 31) // 32) if not args.mode.startswith("hash") else 8
            alg_s1 = N * (sample_bytes + BUCKET_BYTES_PER_KMER * kmers_step / N)
            alg_s2 = BUCKET_BYTES_PER_KMER * kmers_step + (key_b + 4) * D / max(world, 1)
            alg_s34 = (key_b + 4) * D + body_sum[0] * (2 if fmt == "bft" else 1)
            alg_all = alg_s1 + alg_s2 + alg_s34
            # every timed span against the same roofline (algorithmic bytes of its stage per step / its CUDA-event time per step)
            span_alg = {"fq_index": N * sample_bytes, "s1_superk": alg_s1, "hash_hist": BUCKET_BYTES_PER_KMER * kmers_step,
                        "hash_emit": (key_b + 4) * D / max(world, 1), "expand": BUCKET_BYTES_PER_KMER * kmers_step,
                        "radix_sort": (key_b + 4) * D / max(world, 1), "rle": (key_b + 4) * D / max(world, 1), "merge": alg_s34}
            roof["per_span"] = {k: {"ms_per_step": round(v["ms_per_step"], 3), "algorithmic_bytes_per_step": span_alg[k],
                                    "achieved": span_alg[k] / (v["ms_per_step"] * 1e-3) / 1e9,
                                    "frac": span_alg[k] / (v["ms_per_step"] * 1e-3) / 1e9 / peak}
                                for k, v in prof.items() if k in span_alg and v["ms_per_step"] > 0}
            roof["pipeline"] = {"algorithmic_bytes_per_step": alg_all, "achieved": alg_all / (ms_step * 1e-3) / 1e9,
                                "frac": alg_all / (ms_step * 1e-3) / 1e9 / peak, "surviving_key_sample_pairs": D,
                                "what": "S1 text+buckets, S2 buckets+lists, S3/S4 lists+bodies (SURVEY 8d) over ms_per_step"}
        except Exception as e:      # reporting only
            roof["pipeline"] = {"error": str(e)}

    # ---- end to end through host buffers (pinned FASTQ in, bodies out)
    e2e = None
    if not args.no_e2e:
        h_text = C.c_void_p(); h_out = C.c_void_p()
        # pinned staging for the FASTQ: all N samples if host RAM allows (all ranks share the box),
        # otherwise the first K samples, cycled (same bytes moved, same stage-1/2 work)
        K = N
        try:
            import psutil
            avail = psutil.virtual_memory().available
            budget = int(avail * 0.6) // max(world, 1) - len(my_parts) * slab_bytes
            K = max(1, min(N, budget // sample_bytes))
        except Exception:
            pass
        if world > 1:
            t = torch.tensor([K], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN); K = int(t.item())
        rc = L.kmx_host_alloc(K * sample_bytes, C.byref(h_text))
        out_bytes = len(my_parts) * slab_bytes if fmt in ("bf", "bft") else int(body_max[0] * 1.05) + 4096
        rc2 = L.kmx_host_alloc(max(out_bytes, 64), C.byref(h_out))
        if rc or rc2:
            e2e = {"value": None, "unit": "k-mers/s", "error": "pinned host allocation failed"}
        else:
            ck(L.kmx_memcpy_d2h(h, h_text, d_text, K * sample_bytes), "d2h text")
            host_ptrs = (C.c_void_p * N)(*[h_text.value + (s % K) * sample_bytes for s in range(N)])
            step_host = make_step(host_ptrs, 0, args.lanes, to_host=h_out.value)
            step_host()
            ns = max(1, min(args.steps, 2))
            t0 = time.perf_counter()
            ms_e = timed(step_host, ns)
            wall_e = time.perf_counter() - t0
            e2e = {"value": world * kmers_step / (ms_e / ns * 1e-3), "unit": "k-mers/s", "h2d_bytes_per_step": N * sample_bytes * world,
                   "d2h_bytes_per_step": int(body_sum[0]) * world, "ms_per_step": ms_e / ns, "wall_s_per_step": wall_e / ns,
                   "what": "kmx_run_samples on pinned host FASTQ + kmx_merge_partition/kmx_merge_get into pinned host memory",
                   "host_text_samples": K}
            L.kmx_host_free(h_text); L.kmx_host_free(h_out)

    log(f"e2e done: {e2e}")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)

    # ---- N > 1: the path just timed, on small seeded samples, against the CPU oracle (the checker; tests/dist_check.py).
    # Runs in child processes (one per rank, own rendezvous port) under a time limit: a check that fails or hangs is
    # reported as such and cannot take the measurement down with it.
    parity = None
    if world > 1 and not args.no_parity_check:
        env = dict(os.environ)
        env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + 100)
        env["TORCHELASTIC_USE_AGENT_STORE"] = "False"
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--oc-child", "parity", "--gpus", str(world)], env=env,
                               capture_output=True, text=True, timeout=120)
            out = [l for l in r.stdout.splitlines() if l.startswith("{")]
            parity = json.loads(out[-1])["parity_check"] if out else ("ok" if r.returncode == 0 and rank != 0 else f"error: {(r.stderr or 'no output')[-200:]}")
        except subprocess.TimeoutExpired:
            parity = "no result within 120 s"
        except Exception as e:
            parity = f"error: {e}"
        log(f"parity check: {parity}")

    # ---- BASELINE configs 3-5 (their own engines; the main one is closed first to free its HBM)
    # Every rank starts a child process per config (same rank / world, its own rendezvous port) under a hard time limit, so a
    # config that fails or hangs at full size is reported as such and cannot take the headline measurement down with it.
    other = None
    want_other = args.other_configs == "on" or (args.other_configs == "auto" and world == 8)
    dev_bytes_main = int(L.kmx_device_bytes(h))
    if want_other:
        eng.close()
        torch.cuda.empty_cache()
        other = {}
        t_oc = time.perf_counter()
        for idx, (name, c) in enumerate(OTHER_CONFIGS.items()):
            if time.perf_counter() - t_oc > 2.2 * args.oc_timeout_s:
                other[name] = {"workload": c["what"], "skipped": "time budget of the other_configs block used up"}
                continue
            env = dict(os.environ)
            env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + 101 + idx)
            env["TORCHELASTIC_USE_AGENT_STORE"] = "False"      # the children rendezvous among themselves (rank 0's child hosts the store)
            cmd = [sys.executable, os.path.abspath(__file__), "--oc-child", name, "--gpus", str(world), "--lanes", str(args.lanes),
                   "--oc-samples-scale", str(args.oc_samples_scale), "--oc-reads-scale", str(args.oc_reads_scale)]
            try:
                r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=args.oc_timeout_s)
                out = [l for l in r.stdout.splitlines() if l.startswith("{")]
                other[name] = json.loads(out[-1]) if out else {"workload": c["what"], "error": (r.stderr or "no output")[-300:]}
                log(f"{name}: {str(other[name])[:200]}")
            except subprocess.TimeoutExpired as e:
                tail = e.stderr.decode(errors="replace") if isinstance(e.stderr, bytes) else (e.stderr or "")
                tail = " | ".join(l for l in tail.splitlines() if l.startswith("["))[-400:]
                other[name] = {"workload": c["what"], "error": f"no result within {args.oc_timeout_s:.0f} s", "progress": tail}
                for rest in list(OTHER_CONFIGS)[idx + 1:]:
                    other[rest] = {"workload": OTHER_CONFIGS[rest]["what"], "skipped": "an earlier config ran into the time limit"}
                break
            except Exception as e:
                other[name] = {"workload": c["what"], "error": str(e)[:300]}
    if rank == 0:
        line = {"metric": "k-mers/s end-to-end (repart->merge)", "value": value, "unit": "k-mers/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload_name(args, world),
                           "kmers_per_step": kmers_step, "l2": "inputs larger than L2 (315 MB text per launch)",
                           "value_clock": "FASTQ resident in HBM -> all .cmbf bodies in HBM", "lanes": args.lanes,
                           **({"fallback": f"the first attempt made no progress within {args.hang_timeout_s:.0f} s; this is the rerun with one lane per GPU"}
                              if attempt > 1 else {})},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "parity_check": parity, "exchange": exchange,
                "kernel_ms_per_step": {k: round(v["ms_per_step"], 3) for k, v in prof.items()}, "ms_per_step_1lane": ms_1lane,
                "host_wall_ms_per_step": host_wall, "device_bytes": dev_bytes_main, "other_configs": other}
        print(json.dumps(line))
    if not want_other:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


def main_oc_child(args):
    """One other config in a process of its own (see main_kmx)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        peak = 6650.0
    if args.oc_child == "parity":
        try:
            from tests import dist_check
            res = {"parity_check": "ok" if dist_check.check(rank, world, local, quiet=True) else "FAIL"}
        except Exception as e:
            res = {"parity_check": f"error: {str(e)[:200]}"}
        if rank == 0:
            print(json.dumps(res), flush=True)
        os._exit(0)
    c = OTHER_CONFIGS[args.oc_child]
    try:
        res = run_other_config(args.oc_child, c, args, world, rank, local, peak)
    except Exception as e:
        res = {"workload": c["what"], "error": str(e)[:300]}
    if rank == 0:
        print(json.dumps(res), flush=True)
    os._exit(0)                 # no teardown collectives: a rank that failed must not make the others wait


if __name__ == "__main__":
    a = parse_args()
    if a.oc_child:
        main_oc_child(a)
    elif a.impl == "reference":
        main_reference(a)
    else:
        main_kmx(a)
