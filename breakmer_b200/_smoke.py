"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked
against the oracle.  (The oracle import below is the checker, as the task allows
for smoke(); the product modules never import it.)"""


def run():
    from breakmer_b200 import _lib, batch, synth
    from oracle import assembler_py, kmers_py

    regions = [synth.make_region("smoke%d" % i, seed=31 + i, L=1200 + 300 * i, cov=150, k=15, e=0.005,
                                 event=[("del", 200, None), ("ins", 40), ("tdup", 150)][i], indel_p=0.3) for i in range(3)]
    h = _lib.Handle(0)
    try:
        out = batch.run(h, batch.PackedBatch(regions))
        n_ctg = 0
        for i, r in enumerate(regions):
            _ref, _case, _sc, only = kmers_py.sample_only(r.ref_fwd, [x[1] for x in r.reads], [x[1] for x in r.sc_records], r.k)
            exp = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
            assert out.sample_only(i) == only, "sample-only k-mers differ from the oracle"
            assert out.contig_records(i) == exp, "contigs differ from the oracle"
            n_ctg += len(exp)
        assert n_ctg > 0
        print("smoke ok: %d regions, %d contigs, %d check_align, %.2f ms on device" %
              (len(regions), n_ctg, out.n_check_align, out.gpu_ms))
    finally:
        h.close()
