#!/usr/bin/env python
"""Benchmark of the BreaKmer per-target k-mer assembly hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

A "step" is one pass of the whole hot path (target.compare_kmers: k-mer counting,
sample-only selection, read grouping, init_assembly) over one batch of synthetic
target regions.  At N=1 the workload is BASELINE.json configs[1]: the 500-target
gene panel (k=15).  With N ranks every rank owns its own batch of 500 regions -- one sample of
the panel per GPU, the same per-GPU problem on every rank (weak scaling, regions sharded by rank,
no data-path collective).

One JSON line is printed by rank 0.  `value` is whole-job regions/s with the
batch already resident in HBM; `e2e` is the same metric through the C-ABI entry
point bk_compare_kmers_batch with HOST buffers (host->device copies and the
result read-back inside the timed region).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "C1": ("C1: single synthetic 20 kb target region, 100 bp reads at 200x, planted 1.5 kb deletion, k=15", 1),
    "C2": ("C2: 500-target gene panel, synthetic tumor reads with planted indels/inversions/tandem dups, k=15", 500),
    "C3": ("C3: tumor/normal pair, 500 targets with normal-k-mer subtraction, k=15", 500),
    "C4": ("C4: 2000x amplicon-depth panel, 100 amplicons, k=21", 100),
    "C5": ("C5: exome-scale 20,000 target regions with planted translocations, k=15 (2,500 per GPU)", 2500),
}
METRIC = "target_regions_assembled_per_sec"
UNIT = "regions/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_regions(workload, n, slice_index):
    """The regions of one GPU.  Weak scaling runs the SAME per-GPU problem on every rank (one sample of the panel per
    GPU: slice 0 of the generator everywhere), so that the per-N values differ by scaling effects only -- distinct slices
    of the generator differ by +-5 % in DP work and 15 % in their longest region (profiles/r1_scaling.md); `--slice`
    selects another one."""
    from breakmer_b200 import synth
    return list(synth.config_regions(workload, n=n, start=slice_index * n))


# ---------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.dev, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------
# CPU side (the oracle; used only as the reported baseline / reference arm)
# ---------------------------------------------------------------------------------
def cpu_region(region):
    """The reference's CPU path for one region, as restated by the oracle: pure-Python
    k-mer counting + set algebra + init_assembly with the pure-Python olc.nw loop
    (that loop is what the reference itself executes in CPython)."""
    from oracle import assembler_py, kmers_py, nw_py
    normal = [x[1] for x in region.normal_reads] if region.normal_reads else None
    _r, _c, _s, only = kmers_py.sample_only(region.ref_fwd, [x[1] for x in region.reads],
                                            [x[1] for x in region.sc_records], region.k, normal)
    ctg = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len, nw=nw_py.nw)
    return len(only), len(ctg)


def _cpu_region_by_index(args):
    workload, idx = args
    from breakmer_b200 import synth
    t0 = time.time()
    n_only, n_ctg = cpu_region(synth.config_region(workload, idx))
    return n_only, n_ctg, time.time() - t0


def cpu_c_nw_single(regions, max_regions=40):
    """Context only: the same oracle with its C restatement of olc.nw (what a compiled single-core port of the
    reference would roughly do)."""
    from oracle import assembler_py, kmers_py
    t0 = time.time()
    done = 0
    for r in regions[:max_regions]:
        normal = [x[1] for x in r.normal_reads] if r.normal_reads else None
        _r, _c, _s, only = kmers_py.sample_only(r.ref_fwd, [x[1] for x in r.reads], [x[1] for x in r.sc_records], r.k, normal)
        assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
        done += 1
    return done / (time.time() - t0)


def cpu_baseline_single(workload, regions, budget_s=20.0, max_regions=8):
    t0 = time.time()
    done = 0
    kmers = 0
    for r in regions[:max_regions]:
        n_only, _ = cpu_region(r)
        kmers += n_only
        done += 1
        if time.time() - t0 >= budget_s:
            break
    dt = time.time() - t0
    return {"value": done / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first %d regions of %s, single process, regions serial (as sv_processor.py:185), "
                      "oracle port with the pure-Python olc.nw loop; jellyfish absent so the k-mer stage is the "
                      "oracle's dict counter (under-states the reference)" % (done, workload),
            "seconds": round(dt, 2), "sample_only_kmers_per_s": kmers / dt}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    desc, per_gpu = WORKLOADS[args.workload]
    n_total = per_gpu * args.gpus
    per_step = min(2 * cores, n_total)        # two regions per core per step keeps the pool busy past the stragglers
    budget = float(os.environ.get("BK_REF_BUDGET_S", "170"))
    ctx = mp.get_context("fork")
    t_start = time.time()
    timed_regions = 0
    timed_s = 0.0
    timed_kmers = 0
    steps_done = 0
    with ctx.Pool(cores) as pool:
        nxt = 0
        for step in range(args.warmup + args.steps):
            idx = [(args.workload, (nxt + j) % n_total) for j in range(per_step)]
            nxt += per_step
            t0 = time.time()
            res = pool.map(_cpu_region_by_index, idx, chunksize=1)
            dt = time.time() - t0
            if step >= args.warmup:
                timed_regions += len(res)
                timed_s += dt
                timed_kmers += sum(r[0] for r in res)
                steps_done += 1
            if time.time() - t_start > budget and steps_done >= 1:
                break
    value = timed_regions / timed_s if timed_s > 0 else 0.0
    sample = ("%d regions per step (two per host core, dynamic pool) of %s, %d of %d timed steps completed within the %.0f s budget; "
              "oracle port of the reference's CPython path over multiprocessing" %
              (per_step, args.workload, steps_done, args.steps, budget))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * timed_s / max(1, steps_done), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": desc, "regions_per_gpu": per_gpu, "k": 21 if args.workload == "C4" else 15},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_only_kmers_per_s": timed_kmers / timed_s if timed_s > 0 else 0.0,
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE (%d) != --gpus (%d)" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    from breakmer_b200 import _lib, batch
    desc, per_gpu = WORKLOADS[args.workload]
    if args.regions:
        per_gpu = args.regions
    regions = make_regions(args.workload, per_gpu, args.slice)
    pk = batch.PackedBatch(regions).pin()        # pinned host buffers for the end-to-end leg
    h = _lib.Handle(local_rank)
    hbm_peak, peak_src = load_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident-input run: `value` -----------------------------------------------------
    # Steps are independent batches, so the framework keeps `inflight` of them on the device at once (one
    # handle = one stream + its own buffers each, one host thread per handle): the tail of one batch -- a few
    # regions with long serial chains -- overlaps the bulk of the next.  inflight=1 is the strictly sequential mode.
    n_fly = max(1, min(args.inflight, args.steps))
    # one host thread per in-flight batch waits on its stream; when the ranks of this box have fewer cores than such threads,
    # let them sleep on a blocking event instead of spinning (bk_set_option "blocking_sync")
    cores_per_rank = max(1, (os.cpu_count() or 1) // max(1, world))
    blocking = (n_fly > cores_per_rank and not os.environ.get("BK_BENCH_SPIN")) or bool(os.environ.get("BK_BLOCKING_SYNC"))
    handles = [h] + [_lib.Handle(local_rank) for _ in range(n_fly - 1)]
    for hh in handles:
        hh.set_option("blocking_sync", 1 if blocking else 0)
    for hh in handles:
        batch.upload(hh, pk)
    for _ in range(args.warmup):                # W untimed warm-up steps on every handle (arenas reach steady state)
        for hh in handles:
            batch.run(hh, pk, resident=True, decode=False)
    # per-kernel device times: a short SEQUENTIAL pass on one handle with the library's CUDA-event timers on
    # (in the pipelined region kernels of different batches share the SMs, so their durations are not comparable)
    lat_ms = []
    h.set_option("spec_width", int(os.environ.get("BK_BENCH_SEQ_W", "4")))   # latency-oriented setting for the sequential pass
    h.kernel_times_reset(True)
    for _ in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        r = batch.run(h, pk, resident=True, decode=False)
        lat_ms.append(float(r.gpu_ms))
    ktimes = h.kernel_times()
    kt_steps = 3
    spec_w = args.spec_width if args.spec_width else 4
    for hh in handles:
        hh.set_option("spec_width", spec_w)
        batch.run(hh, pk, resident=True, decode=False)
        hh.kernel_times_reset(False)             # timers off, launch counters zeroed for the timed region
    sampler = ClockSampler(local_rank)
    sampler.start()
    last_box = {}

    stagger_s = (min(lat_ms) / 1000.0 / n_fly) if (n_fly > 1 and lat_ms) else 0.0

    def worker(j):
        hh = handles[j]
        if stagger_s:
            time.sleep(j * stagger_s)       # spread the batches over one step latency so that tails and bulks overlap
        for step in range(j, args.steps, n_fly):
            if n_fly == 1:
                flush.zero_()               # L2 flush between timed iterations (sequential mode only)
                torch.cuda.synchronize()
            res = batch.run(hh, pk, resident=True, decode=False)
            last_box[j] = (int(res.n_contigs), int(res.n_check_align), int(res.n_dp_cells),
                           int(res.n_kmer_occurrences), int(res.so_off[res.n_regions]), float(res.gpu_ms),
                           int(res.n_sorted_keys))

    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ev0.record()
    threads = [threading.Thread(target=worker, args=(j,)) for j in range(n_fly)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    barrier()
    wall_s = time.time() - t0
    clocks = sampler.stop()
    gpu_launches = 0
    for hh in handles:
        gpu_launches += int(sum(v[1] for v in hh.kernel_times().values()))
    dev_s = max_over_ranks(ev0.elapsed_time(ev1) / 1000.0)      # device clock across the K steps, max over ranks
    step_ms = [1000.0 * dev_s / args.steps] * args.steps
    n_regions_total = per_gpu * world
    value = n_regions_total * args.steps / dev_s
    n_contigs, n_check, n_cells, n_occ, n_only, _, n_sorted = last_box[0]
    kmers_per_s = sum_over_ranks(float(n_only)) * args.steps / dev_s

    # ---- end to end through the C ABI with host buffers: `e2e` ---------------------------------------
    for hh in handles:
        batch.run(hh, pk, decode=False)
    e2e_box = {}

    def e2e_worker(j):
        hh = handles[j]
        if stagger_s:
            time.sleep(j * stagger_s)
        for step in range(j, args.steps, n_fly):
            e2e_box[j] = batch.run(hh, pk, decode=False)

    barrier()
    t0 = time.time()
    threads = [threading.Thread(target=e2e_worker, args=(j,)) for j in range(n_fly)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    e2e_local = time.time() - t0
    barrier()
    res = e2e_box[0]
    e2e_s = max_over_ranks(e2e_local)
    e2e_value = n_regions_total * args.steps / e2e_s
    h2d = pk.input_bytes + 8 * (len(pk.read_off) + len(pk.sc_off) + len(pk.ref_off)) + len(pk.read_flags)
    out = batch.BatchOutput(res, pk)
    d2h = int(out.seq.nbytes + out.kmer_locs.nbytes + out.indel_only.nbytes + out.others.nbytes + out.reads.nbytes +
              out.kmer_mer.nbytes + 2 * out.kmer_pos.nbytes + out.so_mers.nbytes + out.so_counts.nbytes +
              out.uniq_rec.nbytes + out.uniq_mult.nbytes)

    # ---- same steps with the persistent reference k-mer cache (reported beside the headline, not as it) ----
    ref_cache = None
    if not args.no_ref_cache_leg:
        pk_nr = batch.PackedBatch(regions, with_ref=False).pin()
        for hh in handles:
            hh.ref_cache_build([r.ref_fwd for r in regions], pk.k)
            batch.upload(hh, pk_nr)
            batch.run(hh, pk_nr, resident=True, decode=False)

        def rc_worker(j):
            hh = handles[j]
            if stagger_s:
                time.sleep(j * stagger_s)
            for step in range(j, args.steps, n_fly):
                batch.run(hh, pk_nr, resident=True, decode=False)

        barrier()
        ev0.record()
        threads = [threading.Thread(target=rc_worker, args=(j,)) for j in range(n_fly)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        ev1.record()
        ev1.synchronize()
        barrier()
        rc_s = max_over_ranks(ev0.elapsed_time(ev1) / 1000.0)
        ref_cache = {"value": n_regions_total * args.steps / rc_s, "unit": UNIT, "ms_per_step": 1000.0 * rc_s / args.steps,
                     "note": "reference k-mers of the targets counted once and kept on the device (bk_ref_cache_build), as the "
                             "reference keeps its reference dumps behind marker files (utils.py:157)"}
        for hh in handles:
            hh.ref_cache_clear()

    # ---- ingest row: the same steps starting from FASTA/FASTQ FILES (tmpfs) -------------------
    from_files = None
    if not args.no_ingest_leg:
        from_files = ingest_leg(args, regions, pk, handles, n_fly, stagger_s, barrier, max_over_ranks, n_regions_total, rank)

    # ---- roofline of the dominant kernel (the assembler) and of the dominant k-mer stage kernel -----
    asm_ms, asm_n = ktimes["assemble"]
    asm_ms_per_launch = asm_ms / max(1, asm_n)
    # algorithmic bytes of one assemble launch (DESIGN.md "A-stage"): every unique read once, the sample-only
    # table (12 B/mer), the posting lists (8 B/entry), and the contigs written
    NU = int(out.uniq_reg_off[-1])
    read_lens = pk.read_off[1:] - pk.read_off[:-1]
    uniq_bases = int(read_lens[out.uniq_rec].sum()) if NU else 0
    asm_bytes = uniq_bases + 12 * n_only + d2h
    asm_gbs = asm_bytes / (asm_ms_per_launch * 1e-3) / 1e9 if asm_ms_per_launch > 0 else 0.0
    sc_ms, sc_n = ktimes["sort_scatter"]
    sort_bytes = 2 * 12 * n_sorted       # one pass: keys+values of the sorted (sample) windows read once, written once
    sort_gbs = sort_bytes / (sc_ms / max(1, sc_n) * 1e-3) / 1e9 if sc_ms > 0 else 0.0
    cells_per_s = n_cells * kt_steps / (asm_ms * 1e-3) if asm_ms > 0 else 0.0
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    # INT32 ALU-pipe ceiling of the DP: 8 alu-pipe instructions per cell (nw.cuh), 16 lanes/clk/SMSP (B300_MICROARCH.md)
    int_peak_cells = 148 * 4 * 16 * sm_mhz * 1e6 / 8.0
    # DRAM traffic per launch from the committed ncu --set full captures (profiles/r1_*.md); only valid for the default workload
    asm_traffic = 34850560 if (args.workload == "C2" and per_gpu == 500) else None
    sort_traffic = 191033344 if (args.workload == "C2" and per_gpu == 500) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "from_files": from_files,
        "config": {"workload": desc, "regions_per_gpu": per_gpu, "k": pk.k, "rc_thresh": pk.rc_thresh,
                   "per_gpu_problem": "every rank runs its own batch of the same %d regions (generator slice %d)" % (per_gpu, args.slice),
                   "input_bytes_per_gpu": pk.input_bytes, "steps_in_flight": n_fly, "assembler_spec_width": spec_w,
                   "host_cores_per_rank": cores_per_rank, "host_wait": "blocking" if blocking else "spin",
                   "l2": ("256 MB buffer written between timed steps (flush)" if n_fly == 1 else
                          "%d independent batches in flight on separate buffers; the per-step working set (~0.8 GB of "
                          "key/value, scratch and state arrays) exceeds the 126 MB L2" % n_fly),
                   "timing": "CUDA events bracketing the K steps (barrier + synchronize on both sides), max over ranks",
                   "wall_ms_per_step": 1000.0 * wall_s / args.steps,
                   "sequential_latency_ms_per_step": (min(lat_ms) if lat_ms else None)},
        "sample_only_kmers_per_s": kmers_per_s,
        "with_ref_kmer_cache": ref_cache,
        "per_step": {"contigs": n_contigs, "check_align_calls": n_check, "dp_cells": n_cells,
                     "kmer_occurrences": n_occ, "sorted_keys": n_sorted, "sample_only_kmers": n_only},
        "roofline": {"kernel": "assemble_kernel", "bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": asm_gbs / hbm_peak, "traffic": asm_traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": asm_bytes, "ms_per_launch": asm_ms_per_launch,
                     "share_of_step": asm_ms / (1000.0 * sum(lat_ms) / 1000.0) if lat_ms else None,
                     "note": "the dominant kernel is an integer-issue/latency bound DP state machine that moves "
                             "O(m+n) bytes per O(m*n) cell updates; its HBM fraction is small by construction, "
                             "see roofline_alu and roofline_kstage"},
        "roofline_alu": {"kernel": "assemble_kernel", "bound": "int32 issue", "achieved": cells_per_s, "peak": int_peak_cells,
                         "unit": "DP cell updates/s", "frac": cells_per_s / int_peak_cells if int_peak_cells else None,
                         "achieved_in_flight": n_cells * args.steps / dev_s,
                         "frac_in_flight": (n_cells * args.steps / dev_s) / int_peak_cells if int_peak_cells else None,
                         "note": "`achieved`/`frac`: one launch alone (sequential pass; bounded by the longest region's serial "
                                 "chain, most SMs idle in the tail); `*_in_flight`: all DP cells of the timed region / its device "
                                 "time with %d batches in flight (other stages included)" % n_fly,
                         "peak_source": "148 SM x 4 SMSP x 16 alu lanes/clk x measured sm clock / 8 alu-pipe instructions per cell; "
                                        "ncu: pipe_alu 48% busy on active SMs, launch bounded by the longest region (profiles/r1_assemble_kernel.md)"},
        "roofline_kstage": {"kernel": "rs_scatter_kernel", "bound": "hbm", "achieved": sort_gbs, "peak": hbm_peak,
                            "unit": "GB/s", "frac": sort_gbs / hbm_peak, "traffic": sort_traffic,
                            "algorithmic_bytes_per_launch": sort_bytes, "launches": sc_n},
        "kernel_ms_per_step": {k: round(v[0] / kt_steps, 4) for k, v in ktimes.items() if v[1]},
        "kernel_timing": "sequential pass of %d steps with L2 flush, CUDA events per kernel on the library stream" % kt_steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                "ms_per_step": 1000.0 * e2e_s / args.steps},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single(args.workload, regions)
        line["cpu_baseline"]["value_with_c_nw_for_context"] = cpu_c_nw_single(regions)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    for hh in handles:
        hh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def ingest_leg(args, regions, pk, handles, n_fly, stagger_s, barrier, max_over_ranks, n_regions_total, rank):
    """Every step = bk_ingest_files (text of the four files of each target -> page-locked arrays, host threads) followed by
    bk_compare_kmers_batch.  Files live in tmpfs; writing them is not timed.  Python marshalling of the same regions
    (batch.PackedBatch) is timed beside it for context."""
    import shutil
    import tempfile
    import torch
    from breakmer_b200 import batch, ingest
    root = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    d = tempfile.mkdtemp(prefix="bk_bench_r%d_" % rank, dir=root)
    try:
        refs, fqs, scs, nms = [], [], [], []
        any_normal = any(r.normal_reads for r in regions)
        n_bytes = 0
        for i, r in enumerate(regions):
            base = os.path.join(d, "t%05d" % i)
            texts = {
                "_ref.fa": ">%s\n%s\n" % (r.name, "\n".join(r.ref_fwd[a:a + 60] for a in range(0, len(r.ref_fwd), 60))),
                "_reads.fastq": "".join("%s\n%s\n+\n%s\n" % (rec[0], rec[1], rec[2]) for rec in r.reads),
                "_sc.fa": "".join(">%s\n%s\n" % (rec[0], rec[1]) for rec in r.sc_records),
            }
            if any_normal:
                texts["_normal.fastq"] = "".join("@%s\n%s\n+\n%s\n" % (rec[0].lstrip("@"), rec[1], "I" * len(rec[1]))
                                                 for rec in r.normal_reads)
            for suffix, t in texts.items():
                with open(base + suffix, "w", newline="\n") as f:
                    f.write(t)
                n_bytes += len(t)
            refs.append(base + "_ref.fa"); fqs.append(base + "_reads.fastq"); scs.append(base + "_sc.fa")
            nms.append(base + "_normal.fastq" if any_normal else None)
        normal = nms if any_normal else None
        world = int(os.environ.get("WORLD_SIZE", "1"))
        cores = max(1, (os.cpu_count() or 1) // world)     # this rank's share of the host
        per = max(1, cores // n_fly)
        ings = [ingest.Ingest(n_threads=per) for _ in range(n_fly)]
        kw = dict(normal=normal, k=pk.k, rc_thresh=pk.rc_thresh)
        # parser alone, all cores on one batch
        solo = ingest.Ingest(n_threads=cores)
        solo.files(refs, fqs, scs, **kw)
        t0 = time.time()
        reps = 5
        for _ in range(reps):
            solo.files(refs, fqs, scs, **kw)
        parse_s = (time.time() - t0) / reps
        solo.close()
        t0 = time.time()
        batch.PackedBatch(regions)
        py_s = time.time() - t0
        for j, hh in enumerate(handles):
            batch.run(hh, ings[j].files(refs, fqs, scs, **kw), decode=False)
        box = {}

        def worker(j):
            hh = handles[j]
            if stagger_s:
                time.sleep(j * stagger_s)
            for step in range(j, args.steps, n_fly):
                box[j] = batch.run(hh, ings[j].files(refs, fqs, scs, **kw), decode=False)

        barrier()
        t0 = time.time()
        threads = [threading.Thread(target=worker, args=(j,)) for j in range(n_fly)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        local = time.time() - t0
        barrier()
        ff_s = max_over_ranks(local)
        n_contigs = int(box[0].n_contigs)
        # contig hand-off (row f.3): the contig.setup files of every contig of one batch result, written to tmpfs
        pk0 = ings[0].files(refs, fqs, scs, **kw)
        res0 = batch.run(handles[0], pk0, decode=False)
        hand_s, n_files = [], 0
        writer = ingest.Ingest(n_threads=cores)
        for rep in range(3):
            out_root = os.path.join(d, "handoff%d" % rep)
            t0 = time.time()
            n_files = writer.write_contigs(res0, pk0, [os.path.join(out_root, "t%05d" % i, "contigs") for i in range(len(regions))],
                                            [os.path.join(out_root, "t%05d_clusters.out" % i) for i in range(len(regions))])
            hand_s.append(time.time() - t0)
        writer.close()
        for g in ings:
            g.close()
        return {"value": n_regions_total * args.steps / ff_s, "unit": UNIT, "ms_per_step": 1000.0 * ff_s / args.steps,
                "text_bytes_per_step": n_bytes, "n_contigs": n_contigs,
                "parse_only": {"ms_per_batch": 1000.0 * parse_s, "regions_per_s": len(regions) / parse_s,
                               "text_MB_per_s": n_bytes / parse_s / 1e6, "host_threads": cores},
                "python_marshalling_ms_per_batch": 1000.0 * py_s,
                "handoff": {"ms_per_batch": 1000.0 * min(hand_s), "files_per_batch": n_files, "host_threads": cores,
                            "what": "bk_write_contigs: <id>.fq + <id>.fa per contig and the cluster file per target "
                                    "(sv_processor.py:749-782), tmpfs"},
                "note": "each step parses the targets' FASTA/FASTQ files (tmpfs) with bk_ingest_files into page-locked memory "
                        "and calls bk_compare_kmers_batch; %d ingest threads per in-flight batch" % per}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--regions", type=int, default=0, help="override regions per GPU (debugging)")
    ap.add_argument("--spec-width", type=int, default=0, help="assembler warps per region (0 = auto)")
    ap.add_argument("--inflight", type=int, default=6, help="independent batches (steps) kept on the device at once")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cache-leg", action="store_true")
    ap.add_argument("--no-ingest-leg", action="store_true")
    ap.add_argument("--slice", type=int, default=0, help="which 500-region slice of the generator every rank runs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
