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
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    nb = int(args[0]) if len(args) > 0 else 200
    per = int(args[1]) if len(args) > 1 else 64
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
    all_pairs = sum(len(b) * (len(b) - 1) // 2 for b in batches)
    mean_len = sum(len(x) for x in seqs) / len(seqs)
    cells = n_pairs * mean_len * mean_len                      # estimate: m*n per sweep at the mean read length
    print("bk_dedup_reads: %d batches x %d reads: %d alignments in %d launches (all ordered pairs would be %d), ~%.3g DP cells"
          % (nb, per, n_pairs, h.last_dedup_launches, all_pairs, cells))
    print("  whole call (host buffers in, flags out): %.2f ms = %.1f k reads/s" % (dt * 1e3, len(seqs) / dt / 1e3))
    for name, (ms, launches) in h.kernel_times().items():
        if launches:
            print("  kernel %s: %.3f ms per call in %d launches = ~%.0f G cells/s (one sweep gives both directions)"
                  % (name, ms / reps, launches // reps, cells / (ms / reps * 1e-3) / 1e9))
    h.kernel_times_reset(False)
    t0 = time.perf_counter()
    for _ in range(reps):
        h.dedup_reads(seqs, pos, off, 0.9)
    print("  whole call with the per-kernel event timers off: %.2f ms" % ((time.perf_counter() - t0) / reps * 1e3))
    if "--trace" in sys.argv:
        os.environ["BK_DEDUP_TRACE"] = "1"                      # per-round host timing on stderr (csrc/dedup.cuh)
        h.dedup_reads(seqs, pos, off, 0.9)
        del os.environ["BK_DEDUP_TRACE"]
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
