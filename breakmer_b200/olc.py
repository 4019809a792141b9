"""Drop-in for the reference's olc.py (overlap aligner), running on the GPU.

`nw(seq1, seq2)` keeps the reference signature and return value
(/root/reference/olc.py:40-107): (align1, align2, prej, j, prei, i, max_i).
The scoring constants are the reference's (olc.py:18-20); they are compiled into
the kernel (breakmer_b200/csrc/nw.cuh) and exposed here for information only.
"""
from . import _lib, get_handle

match_award = 1
mismatch_penalty = -2
gap_penalty = -2


def nw_batch(pairs, want_aln=True, device=0):
    """[(seq1, seq2), ...] -> list of the 7-tuples olc.nw returns (align strings are
    '' when want_aln is False).  One warp per pair, one launch for the batch."""
    if not pairs:
        return []
    seqs, pa, pb = [], [], []
    for a, b in pairs:
        if len(a) == 0 or len(b) == 0:
            # olc.py:86-87 reads loop variables that were never bound
            raise NameError("nw: empty sequence (the reference raises NameError)")
        pa.append(len(seqs)); seqs.append(a)
        pb.append(len(seqs)); seqs.append(b)
    try:
        out, alns = get_handle(device).nw_batch(seqs, pa, pb, want_aln=want_aln)
    except _lib.BreakmerError as e:
        if e.code == _lib.BK_ERR_EMPTY_SEQ:
            raise NameError(str(e))
        raise
    res = []
    for i in range(len(pairs)):
        a1, a2 = alns[i] if want_aln else ("", "")
        o = out[i]
        res.append((a1, a2, int(o[0]), int(o[1]), int(o[2]), int(o[3]), int(o[4])))
    return res


def nw(seq1, seq2):
    return nw_batch([(seq1, seq2)])[0]
