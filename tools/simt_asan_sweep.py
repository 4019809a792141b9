"""Out-of-bounds sweep of the kernel SOURCE without a GPU: the whole library on the host SIMT emulator (tests/sim), built
with AddressSanitizer + UBSan and with every device / pinned arena allocation turned into a malloc of exactly the size
asked for (tests/sim_util.build_asan_driver), run over groups of regions and compared with the oracle.  TEST TOOL.

    python tools/simt_asan_sweep.py golden            # the 66 golden scenarios (k=15 and k=21 groups)
    python tools/simt_asan_sweep.py edge              # empty / ragged regions, 300-base reads
    python tools/simt_asan_sweep.py C4 0 2            # regions 0..1 of a BASELINE config (C2, C3, C4, C5)
    SIMT_ORDER=random:3 python tools/simt_asan_sweep.py C5 0 40
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sim_util                                                   # noqa: E402
from breakmer_b200 import synth                                   # noqa: E402
from oracle import assembler_py                                   # noqa: E402
from oracle.make_golden import oracle_sample_only, region_scenarios   # noqa: E402
from test_simt_tsan import _write_region                          # noqa: E402


def run_group(exe, regions, k, label):
    d = tempfile.mkdtemp(prefix="bk_asan_sweep_")
    man = os.path.join(d, "manifest.txt")
    with open(man, "w") as f:
        f.write("\n".join(_write_region(r, d) for r in regions) + "\n")
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    out = subprocess.run([exe, man, str(k), str(regions[0].rc_thresh)], env=env, capture_output=True, text=True, timeout=7200)
    bad = out.returncode != 0 or "Sanitizer" in out.stderr or "runtime error" in out.stderr
    if bad:
        print("%s: SANITIZER REPORT / FAILURE (rc %d)\n%s" % (label, out.returncode, out.stderr[:6000]))
        return False
    got, cur = {}, None
    for line in out.stdout.splitlines():
        p = line.split()
        if p[0] == "region":
            cur = int(p[1]); got[cur] = (int(p[3]), int(p[5]), [])
        elif p[0] == "contig":
            got[cur][2].append(p[1])
    mism = 0
    for i, r in enumerate(regions):
        _a, _b, _c, only = oracle_sample_only(r)
        exp = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
        if got.get(i) != (0, len(only), [c["seq"] for c in exp]):
            mism += 1
            print("%s: region %s differs from the oracle" % (label, r.name))
    print("%s: %d regions, no sanitizer report, %d oracle mismatches" % (label, len(regions), mism))
    return mism == 0


def main(argv):
    exe = sim_util.build_asan_driver()
    ok = True
    if argv[0] == "golden":
        scen = region_scenarios()
        for k in sorted(set(s[1]["k"] for s in scen)):
            regs = [synth.make_region(n, **kw) for n, kw in scen if kw["k"] == k]
            ok &= run_group(exe, regs, k, "golden k=%d" % k)
    elif argv[0] == "edge":
        # empty and ragged regions (tests/test_gpu_pipeline.py::test_empty_and_ragged_regions), first and last in the batch,
        # and 300-base reads (blocked DP with edge buffers)
        r0 = synth.make_region("e0", seed=5, L=800, cov=0, k=15, e=0.0, event=("none",))
        r1 = synth.make_region("e1", seed=6, L=900, cov=200, k=15, e=0.01, event=("del", 100, None))
        r2 = synth.Region(name="e2", k=15, ref_fwd="ACGT" * 100, reads=[("@a:1:1:1:1/1_0", "ACGTN", "IIIII", False)],
                          sc_records=[("a", "AC")])
        r3 = synth.make_region("e3", seed=5, L=800, cov=0, k=15, e=0.0, event=("none",))
        ok &= run_group(exe, [r1, r0, r2, r3], 15, "edge: ragged regions, empty ones last")
        ok &= run_group(exe, [r2, r1], 15, "edge: a five-base read first")
        lr = [synth.make_region("lr%d" % i, seed=700 + i, L=3000, cov=120, k=21, e=0.01, event=("del", 200, None), rl=300,
                                rl_jitter=40) for i in range(2)]
        ok &= run_group(exe, lr, 21, "edge: 300-base reads, k=21")
    else:
        cfg, lo, hi = argv[0], int(argv[1]), int(argv[2])
        regs = [synth.config_region(cfg, i) for i in range(lo, hi)]
        ok &= run_group(exe, regs, regs[0].k, "%s[%d:%d]" % (cfg, lo, hi))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
