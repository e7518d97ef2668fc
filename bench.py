#!/usr/bin/env python
"""Benchmark of the scan hot path: read-names/s scanned + matched (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R]

One step = one pass of the whole scan path over one synthetic NovaSeq-lane shaped input
(config C2: 10+10 bp UDI, 384-sample sheet, `scan -n 1 -rc`): fused parse+pack+count kernel ->
sorted unique list -> matcher pass 1 (forward + reverse-complement) -> per-sample orientation
call -> matcher pass 2.  `value` is kernel-only (input resident in HBM, results left on the
device); `e2e` feeds HOST buffers through the C-ABI (H2D of every input byte and D2H of the
per-key results inside the timed region).  N > 1: one process per GPU (torchrun), each rank owns
one lane (weak scaling); inside the step the per-rank unique tables are merged over NCCL so that every
key ends on one rank (frb_shardmerge), each rank matches its share, and the per-sample orientation sums
are all-reduced (the call of F:354-388 is over all reads of the job).
`--impl reference` times the CPU oracle port of the reference (the reference is a Python script
that cannot travel to the GPU box) on a bounded sample with all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIG = "C2"
N_SUBS = 1
GEN_CHUNK = 8_000_000          # reads per generator call
CPU_SAMPLE_READS = 150_000     # bounded sample for the CPU arm
E2E_MAX_BYTES = 6 << 30        # host-resident sample for the end-to-end leg


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: 400M or what HBM holds)")
    ap.add_argument("--table-log2", type=int, default=0, help="slots of the unique-key tables (default: by input size)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference on a bounded sample (shared by cpu_baseline and
# --impl reference).  This is the one place outside tests/ that executes oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_sample(tmpdir, reads=CPU_SAMPLE_READS):
    import gzip

    from frender_b200 import synth
    spec = synth.make_spec(CONFIG)
    data = synth.generate_big(spec, 0, reads)
    path = os.path.join(tmpdir, "Undetermined_S0_L001_R1_001.fastq.gz")
    with gzip.open(path, "wb", compresslevel=1) as fh:
        fh.write(data)
    return spec, path, len(data)


def cpu_step(spec, path, cores):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import frender_oracle as O
    t0 = time.perf_counter()
    counter = O.tally_barcodes(cores, [path])
    t1 = time.perf_counter()
    results, calls, _ = O.scan_analysis(cores, counter, spec.indexes(), N_SUBS, True)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1, len(results)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    with tempfile.TemporaryDirectory() as d:
        spec, path, nbytes = cpu_sample(d)
        times, parts = [], []
        for i in range(args.warmup + args.steps):
            total, t_tally, t_match, uniq = cpu_step(spec, path, cores)
            if i >= args.warmup:
                times.append(total)
                parts.append((t_tally, t_match))
    ms = 1e3 * sum(times) / len(times)
    value = CPU_SAMPLE_READS / (ms / 1e3)
    sample = (f"{CPU_SAMPLE_READS} reads of the {CONFIG} lane ({nbytes} B decompressed) from one .fastq.gz: "
              f"gzip inflate + tally (1 core: one file) + matcher both passes on {cores} cores; "
              f"{uniq} unique keys")
    line = {
        "impl": "reference", "metric": "read_names_per_s_scanned_matched", "value": value, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64",
        "data": "synthetic", "config": workload_config(CPU_SAMPLE_READS, 1),
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample,
                         "tally_s": sum(p[0] for p in parts) / len(parts),
                         "match_s": sum(p[1] for p in parts) / len(parts)},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(reads_per_gpu, n_gpus):
    return {"workload": "scan -n 1 -rc, synthetic NovaSeq-lane Undetermined R1, 10+10 bp UDI, 384-sample sheet "
                        "(BASELINE configs[1])",
            "reads_per_gpu": reads_per_gpu, "reads_total": reads_per_gpu * n_gpus, "read_len": 151,
            "index": "10+10", "samples": 384, "n_mismatches": N_SUBS, "rc": True,
            "l2_policy": "inputs larger than L2 (no flush needed)" if reads_per_gpu * 373 > (256 << 20)
            else "input smaller than L2"}


# ------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock and throttle-reason sampler for the timed region (the counters of B200_PROFILING.md's
    nvidia-smi line; every 50 ms through NVML, every 200 ms through the nvidia-smi child).  Read through NVML in-process when pynvml is importable: an
    `nvidia-smi -lms` child re-queries the whole device each period and was seen to hold up the sampled
    rank's CUDA calls for 10-600 ms at a time; falls back to that child otherwise."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, device):
        self.proc = self.thread = None
        self.sm, self.mx, self.reasons = [], 0, set()
        try:
            if os.environ.get("FRB_BENCH_NO_CLOCKS"):
                raise RuntimeError("clock sampling disabled")
            import threading

            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip()]
            index = int(ids[device]) if ids and all(v.strip().isdigit() for v in ids) else device
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.stop_flag = threading.Event()

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                        self.reasons.update(name for name, bit in self.BITS.items() if mask & bit)
                    except pynvml.NVMLError:
                        pass
                    self.stop_flag.wait(0.05)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        self.source = "nvidia-smi"
        self.path = tempfile.mktemp(suffix=".csv")
        self.fh = open(self.path, "w")
        try:
            if os.environ.get("FRB_BENCH_NO_CLOCKS"):
                raise OSError("clock sampling disabled")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread:
            self.stop_flag.set()
            self.thread.join()
        else:
            if self.proc:
                self.proc.terminate()
                self.proc.wait()
            self.fh.close()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for row in open(self.path):
                f = [x.strip() for x in row.split(",")]
                if len(f) < 9:
                    continue
                try:
                    self.sm.append(float(f[1]))
                    self.mx = max(self.mx, float(f[2]))
                except ValueError:
                    continue
                for name, flag in zip(names, f[5:9]):
                    if flag.lower().startswith("active"):
                        self.reasons.add(name)
            os.unlink(self.path)
        sm = self.sm
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": self.mx or None,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": self.source}


def run_b200(args):
    from frender_b200 import _lib as L
    from frender_b200 import synth
    from frender_b200.engine import C, Context, PackedSheet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist  # host-side rendezvous only (gloo); the data path is NCCL in the .so
        dist.init_process_group("gloo", rank=rank, world_size=world)

    spec = synth.make_spec(CONFIG, lane=1 + rank % 8)
    lib = L.lib
    probe = Context(local, table_log2=10)
    free_b, total_b = C.c_uint64(), C.c_uint64()
    probe._ck(lib.frb_mem_info(probe._h, C.byref(free_b), C.byref(total_b)))
    probe.close()
    reads = args.reads or 400_000_000
    bytes_per_read = 375.0
    table_log2 = args.table_log2 or (26 if reads > 100_000_000 else (24 if reads > 8_000_000 else 21))
    overhead = 2 * (32 << table_log2) + (8 << 30)
    fit = int((free_b.value - overhead) / bytes_per_read)
    if reads > fit:
        reads = max(1_000_000, fit // 1_000_000 * 1_000_000)
    ctx = Context(local, table_log2=table_log2)
    h = ctx._h
    ck = ctx._ck

    # ---- synthetic lane, generated on the device, resident in HBM -------------------------------
    emit_i7 = np.array([sum(int(c) << (2 * p) for p, c in enumerate(row)) for row in spec.sheet_i7], np.uint32)
    emit_i5 = np.array([sum(int(c) << (2 * p) for p, c in enumerate(row)) for row in spec.emit_i5()], np.uint32)
    cdf = np.ascontiguousarray(spec.cdf, np.uint64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ck(lib.frb_synth_load(h, spec.seed, spec.l1, spec.l2, spec.n_samples, vp(emit_i7), vp(emit_i5), vp(cdf),
                          spec.lane, spec.read_len, spec.sub_t, spec.n_t, spec.rand_t, spec.hop_t))
    cap = int(reads * 374.6) + (64 << 20)
    dbuf = C.c_void_p()
    ck(lib.frb_dev_alloc(h, cap, C.byref(dbuf)))
    g_base = rank * reads
    off, bounds = 0, [0]
    t_gen = time.perf_counter()
    for g in range(0, reads, GEN_CHUNK):
        n = C.c_uint64()
        ck(lib.frb_synth_generate(h, g_base + g, g_base + min(g + GEN_CHUNK, reads), 1,
                                  C.c_void_p(dbuf.value + off), cap - off, C.byref(n)))
        off += n.value
        bounds.append(off)
    nbytes = off
    t_gen = time.perf_counter() - t_gen

    sheet = ctx.load_sheet(PackedSheet(spec.indexes()))
    if world > 1:
        ident = (C.c_char * 128)()
        if rank == 0:
            ck(lib.frb_nccl_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        ck(lib.frb_nccl_init(h, box[0], rank, world))

    def barrier():
        if dist:
            dist.barrier()

    def analyze(want_outputs):
        """matcher pass 1 (rc) -> orientation call -> pass 2; returns unique count"""
        r = ctx.match(N_SUBS, True, None, want_outputs=False)
        f_sum, rc_sum = r["f_sum"], r["rc_sum"]
        if world > 1:   # every rank matched its share of the merged keys: the call is over the whole job
            f_sum, rc_sum = ctx.allreduce(f_sum), ctx.allreduce(rc_sum)
        use = np.array([f_sum[g] < rc_sum[g] for g in sheet.group], np.uint8)               # F:376
        out = ctx.match(N_SUBS, False, use, want_outputs=want_outputs)
        return out

    def step_resident():
        ck(lib.frb_reset(h))
        ck(lib.frb_scan_begin(h, rank, 0))
        ck(lib.frb_scan_chunk_dev(h, dbuf, nbytes, 0, L.RULE_SCAN, None, None))
        n_reads, n_uniq = C.c_uint64(), C.c_uint64()
        ck(lib.frb_scan_end(h, C.byref(n_reads), C.byref(n_uniq)))
        if world > 1:
            ck(lib.frb_shardmerge(h, C.byref(n_uniq)))
        analyze(False)
        return n_reads.value

    def timed(fn, steps, warmup, after_warmup=None):
        for _ in range(warmup):
            fn()
        ck(lib.frb_sync(h))
        if after_warmup:
            after_warmup()
        barrier()
        ms_all = []
        t_wall = time.perf_counter()
        for _ in range(steps):
            ck(lib.frb_timer_start(h))
            got = fn()
            ms = C.c_float()
            ck(lib.frb_timer_stop(h, C.byref(ms)))
            ms_all.append(ms.value)
            if os.environ.get("FRB_BENCH_VERBOSE"):
                print(f"[rank {rank}] step {len(ms_all)}: {ms.value:.2f} ms", file=sys.stderr)
        ck(lib.frb_sync(h))
        t_wall = time.perf_counter() - t_wall
        barrier()
        return sum(ms_all), t_wall, got

    ctx.prof(True)
    mark = {}

    def start_counting():      # the timed steps only: per-class device time and launch count start after warm-up
        for k in range(L.K_NUM):
            ctx.prof_read(k, reset=True)
        mark["launches"] = ctx.launches()

    clocks = Clocks(local)
    ms_total, wall_s, got_reads = timed(step_resident, args.steps, args.warmup, start_counting)
    clk = clocks.stop()
    launches = ctx.launches() - mark["launches"]
    assert got_reads == reads, (got_reads, reads)
    prof = {k: ctx.prof_read(k, reset=True) for k in range(L.K_NUM)}
    n_uniq = C.c_uint64()
    ck(lib.frb_total_finish(h, C.byref(n_uniq)))
    if dist:    # the shares are disjoint: the job's unique keys are their sum
        box = [None] * world
        dist.all_gather_object(box, n_uniq.value)
        n_uniq = C.c_uint64(sum(box))

    # device time of the step, max over ranks
    ms_step = ms_total / args.steps
    if dist:
        box = [None] * world
        dist.all_gather_object(box, ms_step)
        ms_step = max(box)
    value = reads * world / (ms_step / 1e3)

    # ---- roofline of the dominant kernel (parse+pack+count), live CUDA events -------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    scan_ms, scan_n = prof[L.K_SCAN]
    scan_ms_per = scan_ms / max(scan_n, 1)
    achieved = nbytes / (scan_ms_per * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_scan_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = nbytes * tj["traffic_per_algorithmic_byte"]
        traffic_src = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (%s), scaled "
                       "from %d to %d algorithmic bytes" % (tj["source"], tj["algorithmic_bytes"], nbytes))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "scan_ws_kernel (parse+pack+count, warp-specialised)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": scan_ms_per,
                "step_share": {name: prof[k][0] / args.steps for name, k in
                               (("scan_ms", L.K_SCAN), ("export_sort_merge_ms", L.K_EXPORT),
                                ("match_ms", L.K_MATCH), ("clear_ms", L.K_OTHER), ("verify_ms", L.K_VERIFY))}}

    # ---- end to end: host buffers -> H2D -> kernels -> D2H of per-key results -------------------
    e2e = None
    if not args.no_e2e:
        k = max(1, sum(1 for b in bounds[1:] if b <= E2E_MAX_BYTES))
        e_bytes = bounds[k]
        e_reads = min(k * GEN_CHUNK, reads)
        hbuf = C.c_void_p()
        ck(lib.frb_host_alloc(C.byref(hbuf), e_bytes))
        ck(lib.frb_d2h(h, hbuf, dbuf, e_bytes))
        d2h = [0]

        def step_host():
            ck(lib.frb_reset(h))
            ck(lib.frb_scan_begin(h, rank, 0))
            ck(lib.frb_scan_chunk_host(h, hbuf, e_bytes, 0, L.RULE_SCAN))
            n_reads, n_uniq = C.c_uint64(), C.c_uint64()
            ck(lib.frb_scan_end(h, C.byref(n_reads), C.byref(n_uniq)))
            if world > 1:
                ck(lib.frb_shardmerge(h, C.byref(n_uniq)))
            keys, counts, _ = ctx.total_arrays()
            out = analyze(True)
            d2h[0] = keys.nbytes + counts.nbytes + sum(v.nbytes for v in out.values())
            return n_reads.value

        e_steps, e_warm = max(2, min(args.steps, 5)), max(1, min(args.warmup, 3))
        _, e_wall, e_got = timed(step_host, e_steps, e_warm)
        assert e_got == e_reads, (e_got, e_reads)
        e_s = e_wall / e_steps
        if dist:
            box = [None] * world
            dist.all_gather_object(box, e_s)
            e_s = max(box)
        e2e = {"value": e_reads * world / e_s, "unit": "reads/s", "h2d_bytes_per_step": e_bytes,
               "d2h_bytes_per_step": d2h[0], "reads_per_step_per_gpu": e_reads, "ms_per_step": e_s * 1e3,
               "h2d_gbs": e_bytes / e_s / 1e9,
               "sample": "host-pinned decompressed FASTQ (first %d reads of the lane) through frb_scan_chunk_host; "
                         "wall clock, sync both sides" % e_reads}
        ck(lib.frb_host_free(hbuf))

    # ---- end to end WITH host gzip: a .fastq.gz sample through frb_scan_gz (zlib inflate thread, pinned
    #      ring, H2D overlapped with the kernels), as the CLI does; inflate-bound by construction ----------
    e2e_gzip = None
    if rank == 0 and world == 1 and not args.no_e2e:
        import zlib
        g_reads = min(2_000_000, reads)
        g_bytes = bounds[1] if GEN_CHUNK <= g_reads else None
        raw = np.empty(int(g_reads * 376), np.uint8)
        n_raw = C.c_uint64()
        tmpbuf = C.c_void_p()
        ck(lib.frb_dev_alloc(h, raw.size, C.byref(tmpbuf)))
        ck(lib.frb_synth_generate(h, g_base, g_base + g_reads, 1, tmpbuf, raw.size, C.byref(n_raw)))
        ck(lib.frb_d2h(h, vp(raw), tmpbuf, n_raw.value))
        ck(lib.frb_dev_free(h, tmpbuf))
        with tempfile.TemporaryDirectory() as d:
            gz_path = os.path.join(d, "Undetermined_S0_L001_R1_001.fastq.gz")
            z = zlib.compressobj(1, zlib.DEFLATED, 31)
            with open(gz_path, "wb") as fh:
                view = memoryview(raw)[:n_raw.value]
                for off in range(0, n_raw.value, 64 << 20):
                    fh.write(z.compress(view[off:off + (64 << 20)]))
                fh.write(z.flush())
            gz_size = os.path.getsize(gz_path)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                ctx.reset()
                got_r, got_u, got_raw = ctx.scan_gz(gz_path, 0)
                ctx.total_arrays()
                analyze(True)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            assert got_r == g_reads and got_raw == n_raw.value
            # the same through the CLI's `-c N` path: N files inflated and scanned at the same time, one
            # context per stream on this GPU, per-file lists folded into one total on the device (F:189-203)
            from frender_b200.cli import scan_files_concurrent
            streams = max(2, min(8, len(os.sched_getaffinity(0))))
            files = []
            for k in range(streams):
                pth = os.path.join(d, f"lane{k}_R1.fastq.gz")
                os.link(gz_path, pth)
                files.append(pth)
            best_m = None
            for _ in range(2):
                t0 = time.perf_counter()
                per_file, _tot = scan_files_concurrent(files, None, streams, local, 21, ctx)
                analyze(True)
                dt = time.perf_counter() - t0
                best_m = dt if best_m is None else min(best_m, dt)
            assert all(per_file[k][0] == g_reads for k in range(streams))
        e2e_gzip = {"value": g_reads / best, "unit": "reads/s", "reads": g_reads, "gz_bytes": gz_size,
                    "raw_bytes": n_raw.value, "inflate_gbs": n_raw.value / best / 1e9,
                    "note": "one .fastq.gz stream; single zlib inflate thread is the bound (SURVEY 8f-1)",
                    "multi_file": {"value": streams * g_reads / best_m, "unit": "reads/s", "files": streams,
                                   "streams": streams, "inflate_gbs": streams * n_raw.value / best_m / 1e9,
                                   "note": "`-c N`: N files at the same time, one context and one zlib thread each"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = len(os.sched_getaffinity(0))
        with tempfile.TemporaryDirectory() as d:
            cspec, path, cbytes = cpu_sample(d)
            total, t_tally, t_match, uniq = cpu_step(cspec, path, cores)
        cpu = {"value": CPU_SAMPLE_READS / total, "unit": "reads/s", "cores": cores, "kind": "port",
               "sample": f"{CPU_SAMPLE_READS} reads of the same lane shape from one .fastq.gz: inflate+tally "
                         f"{t_tally:.2f}s (1 core, one file) + matcher both passes {t_match:.2f}s on {cores} cores"}

    if rank == 0:
        line = {
            "metric": "read_names_per_s_scanned_matched", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64", "data": "synthetic (device-generated, "
            "counter-based; byte-identical to frender_b200/synth.py)",
            "config": workload_config(reads, world), "clocks": clk, "e2e": e2e, "e2e_gzip": e2e_gzip,
            "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu,
            "unique_keys": n_uniq.value, "input_bytes_per_gpu": nbytes, "gen_s": t_gen,
            "kernel_only_reads_per_s_per_gpu": reads / (scan_ms_per * 1e-3),
        }
        print(json.dumps(line))
    ck(lib.frb_dev_free(h, dbuf))
    ctx.close()
    if dist:
        dist.destroy_process_group()


def main():
    # stdout carries the one JSON line: NCCL's version/warning banner (NCCL_DEBUG=VERSION|WARN in some
    # environments) goes to stderr, or away
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        del os.environ["NCCL_DEBUG"]
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
