"""Small workload for compute-sanitizer runs (memcheck / racecheck): nw batch, k-mer counting, a 6-region pipeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from breakmer_b200 import _lib, batch, synth
h = _lib.Handle(0)
out, alns = h.nw_batch(["ACGTACGTTTGACCA" * 9, "TTGACCAACGTAC" * 5, "A" * 300 + "CGT" * 30], [0, 1, 2], [1, 0, 0], want_aln=True)
print(out[:, :5].tolist())
m, c = h.count_kmers(["ACGTNACGTACGGTACCA" * 20, "TTTT", ""], 15)
print(len(m))
regions = [synth.make_region("s%d" % i, seed=900 + i, L=900, cov=120, k=15, e=0.01, event=[("del", 80, None), ("ins", 40), ("inv", 200)][i % 3],
                             indel_p=0.3, rl=[100, 100, 150, 300][i % 4]) for i in range(6)]
for w in (4, 1):
    h.set_option("spec_width", w)
    o = batch.run(h, batch.PackedBatch(regions))
    print(w, o.n_contigs, o.n_check_align)
# normal-sample subtraction (probe mode with two streamed sets) and a batch without reference windows
nregs = list(synth.config_regions("C3", 2))
o = batch.run(h, batch.PackedBatch(nregs, with_normal=True))
print("normal", o.n_contigs, int(o.so_off[-1]))
h.ref_cache_build([r.ref_fwd for r in regions], 15)
o = batch.run(h, batch.PackedBatch(regions, with_ref=False))
print("ref cache", o.n_contigs)
h.ref_cache_clear()
