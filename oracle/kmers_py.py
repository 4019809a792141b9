"""Oracle restatement of the k-mer stage.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

K1  `run_jellyfish` (utils.py:151-179) shells out to **jellyfish 1.1.11**
    (`kmer_region.config:9`; `count -m k -s 100000000 -t 8 -o`, utils.py:160;
    `dump -c`, utils.py:166).  Jellyfish is a third-party binary that is not in
    the reference tree and not installed here, and the reference holds no test
    that pins its output: **parity unpinned** for this row.  The published
    behaviour restated here: forward-strand windows only (no `-C`, Q1); a
    window containing any character outside ACGT/acgt is skipped; windows never
    span records; lower case is folded to upper case; `dump -c` prints
    "<MER> <count>" one per line with exact counts, in hash order (order is
    irrelevant to the caller, which loads it into a dict).
K2  `load_kmers` (utils.py:287-297): accumulate "<mer> <count>" lines of one or
    more dump files into a dict.
K3  `compare_kmers` (sv_processor.py:609-632): sample_only = (case & case_sc)
    - ref, value = case count (Q2, Q3, Q4).
K4  normal-sample subtraction (absent from the reference):
    sample_only -= set(k-mers of the normal sample's reads for the region).
"""
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
_VALID = set("ACGT")


def count_kmers(seqs, k, into=None):
    """Strand-specific k-mer occurrence counts over an iterable of record
    sequences (jellyfish count + dump -c + load_kmers, rows K1+K2)."""
    counts = {} if into is None else into
    for s in seqs:
        s = s.upper()
        n = len(s)
        run = 0                      # length of the current run of valid bases
        for i in range(n):
            if s[i] in _VALID:
                run += 1
            else:
                run = 0
            if run >= k:
                mer = s[i - k + 1:i + 1]
                counts[mer] = counts.get(mer, 0) + 1
    return counts


def revcomp(s):
    # Biopython reverse_complement of an ACGTN string (utils.py:369)
    return "".join(_COMP.get(c, "N") for c in reversed(s))


def dump_lines(counts):
    """The text `jellyfish dump -c` would write (order: sorted, see above)."""
    return ["%s %d" % (m, counts[m]) for m in sorted(counts)]


def sample_only(ref_fwd, read_seqs, sc_seqs, k, normal_seqs=None):
    """Rows K1-K4 for one region.  Returns (ref, case, case_sc, case_only)."""
    ref = count_kmers([ref_fwd], k)
    count_kmers([revcomp(ref_fwd)], k, into=ref)            # sv_processor.py:613-615 (Q2)
    case = count_kmers(read_seqs, k)                         # :618 (Q3)
    case_sc = count_kmers(sc_seqs, k)                        # :620
    sc_mers = set(case) & set(case_sc)                       # :621
    only = sc_mers - set(ref)                                # :622
    if normal_seqs is not None:
        only -= set(count_kmers(normal_seqs, k))             # K4
    case_only = {m: case[m] for m in only}                   # :630-631
    return ref, case, case_sc, case_only
