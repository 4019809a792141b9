"""Race detection for the kernel source: the emulator of tests/sim/simt_host.h under ThreadSanitizer.  Every emulated GPU
thread is a TSan fiber; fiber switches carry no synchronisation, so the only happens-before edges TSan sees are the ones
the kernels create -- warp collectives, __syncthreads(), kernel boundaries; atomics are real atomics.  A report is then
what CUDA calls a data race: two threads of a block touching the same address, at least one writing, with no barrier
between them.  (compute-sanitizer's racecheck covers shared memory on the device; this covers global memory too, in the
container without a GPU.)

  * the detector is tested on the pattern it was built for, with and without the barrier;
  * the whole library (tests/sim/simt_tsan_driver.cpp: ingest -> bk_compare_kmers_batch, then bk_count_kmers,
    bk_sample_only, bk_nw_batch, bk_dedup_reads and the reference k-mer cache: every kernel) runs golden regions under
    it: no reports, results equal to the oracle."""
import os
import subprocess
import tempfile

import pytest

import sim_util
from breakmer_b200 import synth
from oracle import assembler_py
from oracle.make_golden import oracle_sample_only, region_scenarios

SUPP = os.path.join(sim_util.SIM_DIR, "tsan_suppressions.txt")      # one documented benign race (plain read next to an atomicOr)
ENV = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 suppressions=%s" % SUPP)


def _tsan_works():
    return subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", os.devnull], input="int main(){return 0;}",
                          capture_output=True, text=True).returncode == 0


pytestmark = pytest.mark.skipif(not _tsan_works(), reason="g++ -fsanitize=thread is not usable here")


def test_the_detector_sees_the_flag_race_and_accepts_the_fix(tmp_path):
    exe = os.path.join(str(tmp_path), "selftest")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-w", "-DSIMT_TSAN", "-fsanitize=thread", "-I", sim_util.SIM_DIR,
                           "-o", exe, os.path.join(sim_util.SIM_DIR, "simt_tsan_selftest.cpp")])
    racy = subprocess.run([exe, "racy"], env=ENV, capture_output=True, text=True)
    if "unexpected memory mapping" in racy.stderr:
        pytest.skip("the ThreadSanitizer runtime cannot map its shadow memory on this kernel configuration")
    assert racy.stderr.count("WARNING: ThreadSanitizer: data race") >= 1 and "simt_tsan_selftest.cpp" in racy.stderr
    fixed = subprocess.run([exe, "fixed"], env=ENV, capture_output=True, text=True)
    assert fixed.returncode == 0 and "ThreadSanitizer" not in fixed.stderr and "done" in fixed.stdout


def _write_region(r, d):
    base = os.path.join(d, r.name)
    with open(base + "_ref.fa", "w") as f:
        f.write(">%s\n%s\n" % (r.name, r.ref_fwd))
    with open(base + ".fastq", "w") as f:
        for rid, seq, qual, _io in r.reads:
            f.write("%s\n%s\n+\n%s\n" % (rid, seq, qual))
    with open(base + "_sc.fa", "w") as f:
        for name, seq in r.sc_records:
            f.write(">%s\n%s\n" % (name, seq))
    cols = [base + "_ref.fa", base + ".fastq", base + "_sc.fa"]
    if r.normal_reads:
        with open(base + "_normal.fastq", "w") as f:
            for rec in r.normal_reads:
                f.write("@%s\n%s\n+\n%s\n" % (rec[0].lstrip("@"), rec[1], "I" * len(rec[1])))
        cols.append(base + "_normal.fastq")
    return "\t".join(cols)


def test_whole_library_under_address_sanitizer():
    """The same executable under AddressSanitizer + UBSan with every arena allocation turned into an exact-size malloc: a
    kernel touching one element past any device / pinned array would be reported.  Opt-in (BK_TEST_ASAN=1: the build takes
    a minute); the result on the final source is in profiles/r2_emulator_runs.md."""
    if not os.environ.get("BK_TEST_ASAN"):
        pytest.skip("set BK_TEST_ASAN=1 to build and run the AddressSanitizer executable")
    exe = sim_util.build_asan_driver()
    scen = [s for s in region_scenarios() if s[1]["k"] == 15]
    regions = [synth.make_region(n, **kw) for n, kw in scen[:8]] + [synth.config_region("C3", 2)]
    d = tempfile.mkdtemp(prefix="bk_asan_")
    man = os.path.join(d, "manifest.txt")
    with open(man, "w") as f:
        f.write("\n".join(_write_region(r, d) for r in regions) + "\n")
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    out = subprocess.run([exe, man, "15", str(regions[0].rc_thresh)], env=env, capture_output=True, text=True, timeout=1800)
    assert out.returncode == 0 and "Sanitizer" not in out.stderr and "runtime error" not in out.stderr, out.stderr[:4000]
    assert out.stdout.count("region ") == len(regions)


@pytest.mark.parametrize("order", ["random:3"])
def test_whole_library_has_no_warp_or_block_level_race_on_golden_regions(order):
    exe = sim_util.build_tsan_driver()
    scen = [s for s in region_scenarios() if s[1]["k"] == 15]
    regions = [synth.make_region(n, **kw) for n, kw in (scen[i] for i in (0, 2, 4, 9, 13, 14))]
    if order:
        regions.append(synth.config_region("C3", 2))            # a tumour / normal region (normal subtraction, K4)
    d = tempfile.mkdtemp(prefix="bk_tsan_")
    man = os.path.join(d, "manifest.txt")
    with open(man, "w") as f:
        f.write("\n".join(_write_region(r, d) for r in regions) + "\n")
    env = dict(ENV)
    if order:
        env["SIMT_ORDER"] = order
    out = subprocess.run([exe, man, "15", str(regions[0].rc_thresh)], env=env, capture_output=True, text=True, timeout=1200)
    if "unexpected memory mapping" in out.stderr:
        pytest.skip("the ThreadSanitizer runtime cannot map its shadow memory on this kernel configuration")
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ThreadSanitizer" not in out.stderr, out.stderr[:4000]
    got, cur = {}, None
    for line in out.stdout.splitlines():
        p = line.split()
        if p[0] == "region":
            cur = int(p[1]); got[cur] = (int(p[5]), [])
        elif p[0] == "contig":
            got[cur][1].append(p[1])
    for i, r in enumerate(regions):
        _a, _b, _c, only = oracle_sample_only(r)
        exp = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
        assert got[i] == (len(only), [c["seq"] for c in exp]), r.name
    assert "ThreadSanitizer" not in out.stderr
    # the entry points beside the batched path ran under the detector too (bk_count_kmers / bk_sample_only, bk_nw_batch
    # with and without alignment strings, bk_dedup_reads, the reference k-mer cache, a pair above 4,095 bases = nw_long_kernel)
    legs = {l.split()[0]: l.split() for l in out.stdout.splitlines()
            if l.split()[0] in ("count_kmers", "nw_batch", "nw_long", "dedup", "ref_cache")}
    assert sorted(legs) == ["count_kmers", "dedup", "nw_batch", "nw_long", "ref_cache"]
    assert all(v[1:3] == ["rc", "0"] for v in legs.values())
    assert int(legs["nw_long"][4]) > 100 and legs["nw_long"][4] == legs["nw_long"][7].rstrip(")")      # (a 4,300-base pair)
    assert legs["ref_cache"][4] == legs["ref_cache"][-1].rstrip(")")            # same contigs with the cache as with the sequences
