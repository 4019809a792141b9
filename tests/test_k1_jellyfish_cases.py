"""K1 hardening: every place where the reader / counter of jellyfish 1.1.11 (`jellyfish count -m k -s ... -o`,
`jellyfish dump -c`; call sites utils.py:160,166 of the reference) could differ from this repo's restatement gets an
explicit known-answer case.  jellyfish is a third-party binary that is absent from the reference tree and from this
image, and the reference has no test that pins it, so these answers are derived from its documented behaviour
(jellyfish 1.1 manual, sections "Counting k-mers" and "Input files"; parse_dna / parse_qual_dna of its source):
the row stays PARITY UNPINNED -- what these cases pin is that the oracle, the Python drop-in
(utils.read_sequences / run_jellyfish / load_kmers), the native ingest (bk_ingest_*) and the CUDA counters all
implement the SAME stated semantics, case by case.

  C1  multi-line FASTA: the lines of a record are concatenated, k-mers run across line breaks      (manual: "multi-fasta")
  C2  any character outside ACGTacgt (N, IUPAC codes, '-', '*') ends the current k-mer window       (parse_dna: codes < 0 reset)
  C3  lower case is counted as upper case; the dump prints upper case                               (parse_dna folds case)
  C4  FASTQ is read by line position: a quality line that starts with '@' is not a header           (parse_qual_dna: 4-line records)
  C5  CR LF line ends: the '\\r' is white space to the readers here (stripped); see the note in the test
  C6  empty records and records shorter than k contribute nothing; k-mers never span records        (manual: "each sequence separately")
  C7  the format is chosen by the FIRST byte of the file: '>' FASTA, '@' FASTQ; an empty file has no k-mers
  C8  counts are occurrences, not presence: duplicated records and repeats add up                   (manual: "count")
  C9  the strand is the one given: no canonicalisation without -C (utils.py:160 passes no -C)
"""
import os

import pytest

from oracle import ingest_py, kmers_py

CASES = [
    # name, text, k, expected {mer: count}
    ("c1_multiline_fasta", ">r1 two lines\nACG\nTAC\n", 3, {"ACG": 1, "CGT": 1, "GTA": 1, "TAC": 1}),
    ("c1_multiline_two_records", ">a\nAC\nGT\n>b\nTTT\nT\n", 3, {"ACG": 1, "CGT": 1, "TTT": 2}),
    ("c2_n_breaks_window", ">r\nACGNACGT\n", 3, {"ACG": 2, "CGT": 1}),
    ("c2_iupac_and_gap_chars", ">r\nACGRTTT-GGG*CCC\n", 3, {"ACG": 1, "TTT": 1, "GGG": 1, "CCC": 1}),
    ("c3_lower_case_folded", ">r\nacgTAcg\n", 3, {"ACG": 2, "CGT": 1, "GTA": 1, "TAC": 1}),
    ("c4_fastq_quality_line_starting_with_at", "@i:1:1:1:1/1_0\nACGT\n+\n@III\n@i:1:1:1:2/1_0\nTTTT\n+\n>>>>\n", 3,
     {"ACG": 1, "CGT": 1, "TTT": 2}),
    ("c4_fastq_plus_line_with_text", "@i:1:1:1:1/1_0\nGGGA\n+i:1:1:1:1/1_0\nIIII\n", 3, {"GGG": 1, "GGA": 1}),
    ("c5_crlf_single_line", ">r\r\nACGT\r\n", 3, {"ACG": 1, "CGT": 1}),
    ("c5_crlf_fastq", "@i:1:1:1:1/1_0\r\nACGT\r\n+\r\nIIII\r\n", 3, {"ACG": 1, "CGT": 1}),
    ("c6_empty_and_short_records", ">e\n\n>s\nAC\n>ok\nACGA\n>e2\n", 3, {"ACG": 1, "CGA": 1}),
    ("c6_no_span_across_records", ">a\nAAC\n>b\nCGG\n", 3, {"AAC": 1, "CGG": 1}),
    ("c6_fastq_empty_sequence", "@i:1:1:1:1/1_0\n\n+\n\n@i:1:1:1:2/1_0\nCCCC\n+\nIIII\n", 3, {"CCC": 2}),
    ("c7_empty_file", "", 3, {}),
    ("c7_fasta_sniffed_by_first_byte", ">@not a fastq\nACGT\n", 4, {"ACGT": 1}),
    ("c8_occurrences_add_up", ">a\nACACAC\n>b\nACACAC\n", 2, {"AC": 6, "CA": 4}),
    ("c9_strand_specific", ">r\nAAAC\n", 4, {"AAAC": 1}),           # GTTT (the reverse complement) is NOT counted
    ("c2_window_longer_than_valid_run", ">r\nACNGTNAA\n", 3, {}),
]
IDS = [c[0] for c in CASES]


@pytest.mark.parametrize("name,text,k,expected", CASES, ids=IDS)
def test_oracle_matches_the_stated_semantics(name, text, k, expected):
    assert kmers_py.count_kmers(ingest_py.kmer_sequences(text), k) == expected


@pytest.mark.parametrize("name,text,k,expected", CASES, ids=IDS)
def test_python_reader_and_host_ingest_see_the_same_records(name, text, k, expected, tmp_path):
    from breakmer_b200 import ingest, utils
    fn = os.path.join(str(tmp_path), name + ".txt")
    with open(fn, "w", newline="") as f:
        f.write(text)
    seqs = utils.read_sequences(fn)
    assert seqs == ingest_py.kmer_sequences(text)
    assert kmers_py.count_kmers(seqs, k) == expected
    # native ingest (host code, no device needed when pinned=False): the same records as the soft-clip input
    g = ingest.Ingest(n_threads=2, pinned=False)
    try:
        pk = g.files([None], [None], [fn], k=k)
        assert pk.sequences("sc") == [seqs]
    finally:
        g.close()


def test_crlf_inside_a_multi_line_fasta_record_is_a_documented_choice():
    """With CR LF line ends real jellyfish sees '\\r' as a non-ACGT byte, which for a MULTI-line record would end the
    window at every line break.  The readers here strip it with the other white space (as CPython's str.strip() does in
    the reference's own FastqFile), so windows still run across the lines.  BreaKmer writes all four inputs itself with
    '\\n' line ends (utils.py:367-371, 436-443; sv_processor.py:499-519), so the two readings cannot differ on its files;
    the choice is pinned here so that it cannot change silently."""
    text = ">r\r\nACG\r\nTAC\r\n"
    assert kmers_py.count_kmers(ingest_py.kmer_sequences(text), 3) == {"ACG": 1, "CGT": 1, "GTA": 1, "TAC": 1}


@pytest.mark.gpu
@pytest.mark.parametrize("name,text,k,expected", CASES, ids=IDS)
def test_gpu_run_jellyfish_and_load_kmers(name, text, k, expected, tmp_path):
    """The drop-in pair the reference calls (sv_processor.py:615-620): CUDA counting behind run_jellyfish, dump file in
    `jellyfish dump -c` format, load_kmers."""
    from breakmer_b200 import utils
    fn = os.path.join(str(tmp_path), name + ".fa")
    with open(fn, "w", newline="") as f:
        f.write(text)
    dump = utils.run_jellyfish(fn, "jellyfish", k)
    assert utils.load_kmers(dump, {}) == expected
    with open(dump) as f:
        assert all(line.split()[0].isupper() for line in f.read().splitlines())      # dump prints upper case


@pytest.mark.gpu
def test_gpu_pipeline_counts_these_inputs_the_same_way(tmp_path):
    """The batched path (native ingest + one CTA per region, region_kmers.cuh) on inputs built from the same cases: the
    case text is the read AND the soft-clip input of a region with an unrelated reference, so sample-only == the counts."""
    from breakmer_b200 import _lib, batch, ingest
    k = 3
    picked = [c for c in CASES if c[2] == k and c[1].startswith(">")]
    d = str(tmp_path)
    refs, fqs, scs, exp = [], [], [], []
    for name, text, _k, expected in picked:
        seqs = ingest_py.kmer_sequences(text)
        ref = os.path.join(d, name + "_ref.fa")
        with open(ref, "w") as f:
            f.write(">ref\n\n")                              # empty reference: nothing is subtracted
        fq = os.path.join(d, name + ".fastq")
        with open(fq, "w", newline="") as f:
            for i, s in enumerate(seqs):
                f.write("@i:1:1:1:%d/1_0\n%s\n+\n%s\n" % (i + 1, s, "I" * len(s)))
        sc = os.path.join(d, name + "_sc.fa")
        with open(sc, "w", newline="") as f:
            f.write(text)
        refs.append(ref); fqs.append(fq); scs.append(sc); exp.append(expected)
    g = ingest.Ingest(n_threads=2)
    h = _lib.Handle(0)
    try:
        out = batch.run(h, g.files(refs, fqs, scs, k=k, rc_thresh=2))
        for i, expected in enumerate(exp):
            assert out.sample_only(i) == expected, picked[i][0]
    finally:
        h.close()
        g.close()
