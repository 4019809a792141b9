"""Time bk_dedup_reads on batches shaped like deep amplicon seeds (device kernel vs whole call) and the oracle
beside it on a bounded sample.  python tools/dedup_profile.py [n_batches] [reads_per_batch]"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    from breakmer_b200 import get_handle
    from oracle import redundancy_py as R
    from test_oracle_redundancy import random_batch
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rng = random.Random(9)
    batches = [random_batch(rng, per, "b%d" % i) for i in range(nb)]
    seqs = [r[1] for b in batches for r in b]
    pos = [r[2] for b in batches for r in b]
    off = [0]
    for b in batches:
        off.append(off[-1] + len(b))
    h = get_handle(0)
    h.dedup_reads(seqs, pos, off, 0.9)
    h.kernel_times_reset(True)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        check, flags, n_pairs = h.dedup_reads(seqs, pos, off, 0.9)
    dt = (time.perf_counter() - t0) / reps
    cells = 0
    for b in batches:
        ls = [len(r[1]) for r in b]
        s = sum(ls)
        cells += (s * s - sum(x * x for x in ls)) // 2
    print("bk_dedup_reads: %d batches x %d reads, %d alignments (both directions in one sweep), %.3g DP cells per sweep set"
          % (nb, per, n_pairs, cells))
    print("  whole call (host buffers in, flags out): %.2f ms = %.1f k reads/s, %.2f G cell updates/s (x2 directions)"
          % (dt * 1e3, len(seqs) / dt / 1e3, 2 * cells / dt / 1e9))
    for name, (ms, launches) in h.kernel_times().items():
        if launches:
            print("  kernel %s: %.2f ms per launch (%d launches) = %.1f G cells/s per sweep"
                  % (name, ms / launches, launches, cells / (ms / launches * 1e-3) / 1e9))
    h.kernel_times_reset(False)
    t0 = time.perf_counter()
    n_or = min(nb, 3)
    for b in batches[:n_or]:
        R.dedup_batch(b, 0.9)
    dt_or = (time.perf_counter() - t0) / n_or
    print("  oracle (C nw, sequential chain, only the alignments it needs): %.2f ms per batch -> %.1f k reads/s"
          % (dt_or * 1e3, per / dt_or / 1e3))
    print("  kept %d of %d reads, %d flagged redundant" % (int((flags & 1).sum()), len(seqs), int((flags & 2 > 0).sum())))


if __name__ == "__main__":
    main()
