"""Oracle restatement of the reference overlap aligner ``olc.nw``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/olc.py:18-107:
  * scores +1 match / -2 mismatch / -2 gap                      (olc.py:18-20)
  * row 0 and column 0 of the score table are 0 (free leading gaps on both
    sequences)                                                   (olc.py:47-52)
  * pointer[i][0] = 1, pointer[0][j] = 2                         (olc.py:55-59)
  * cell = max(left, up, diag); pointer priority diag(3) > up(2) > left(1),
    where "up" is score[i][j-1] (consumes seq1) and "left" is score[i-1][j]
    (consumes seq2)                                              (olc.py:62-74)
  * end cell: last column, ``>=`` scan so the LARGEST row wins ties
                                                                 (olc.py:79-83)
  * do-while traceback until i == 0 or j == 0                    (olc.py:90-105)
  * returns (align1, align2, prej, j, prei, i, max_i)            (olc.py:107)

Pinned against the reference itself by tests/golden/nw_golden.json (generated
by oracle/make_golden.py through oracle/ref_shim.py).

``nw`` is the literal pure-Python form (this is also what the CPU baseline
times, because the reference runs exactly this loop in CPython).
``nw_fast`` returns the same 7-tuple through the C restatement in
oracle/c/oracle.c when that library has been built; tests check the two agree.
"""
import ctypes
import os

MATCH = 1
MISMATCH = -2
GAP = -2


def nw(seq1, seq2):
    m = len(seq1)
    n = len(seq2)
    if m == 0 or n == 0:
        # olc.py:86-87 reads the loop variables i/j, which were never bound
        raise NameError("nw: empty sequence (the reference raises NameError)")
    width = m + 1
    score = [[0] * width for _ in range(n + 1)]
    ptr = [[0] * width for _ in range(n + 1)]
    for i in range(n + 1):
        ptr[i][0] = 1
    for j in range(width):
        ptr[0][j] = 2
    for i in range(1, n + 1):
        b = seq2[i - 1]
        prev = score[i - 1]
        cur = score[i]
        prow = ptr[i]
        for j in range(1, width):
            diag = prev[j - 1] + (MATCH if seq1[j - 1] == b else MISMATCH)
            up = cur[j - 1] + GAP
            left = prev[j] + GAP
            best = left
            if up > best:
                best = up
            if diag > best:
                best = diag
            cur[j] = best
            if best == diag:
                prow[j] = 3
            elif best == up:
                prow[j] = 2
            else:
                prow[j] = 1
    max_i = -200
    i = 0
    for ii in range(n + 1):
        if score[ii][m] >= max_i:
            max_i = score[ii][m]
            i = ii
    j = m
    prei, prej = i, j
    a1 = []
    a2 = []
    while True:
        p = ptr[i][j]
        if p == 3:
            a1.append(seq1[j - 1])
            a2.append(seq2[i - 1])
            i -= 1
            j -= 1
        elif p == 2:
            a2.append('-')
            a1.append(seq1[j - 1])
            j -= 1
        else:
            a2.append(seq2[i - 1])
            a1.append('-')
            i -= 1
        if i == 0 or j == 0:
            break
    return ("".join(reversed(a1)), "".join(reversed(a2)), prej, j, prei, i, max_i)


# --------------------------------------------------------------------------
# C restatement (oracle/c/oracle.c), loaded lazily.
# --------------------------------------------------------------------------
_C = None
_C_TRIED = False


def c_lib():
    global _C, _C_TRIED
    if _C_TRIED:
        return _C
    _C_TRIED = True
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c", "liboracle_c.so")
    if os.path.isfile(path):
        lib = ctypes.CDLL(path)
        lib.oracle_nw.restype = ctypes.c_int
        lib.oracle_nw.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_char_p]
        _C = lib
    return _C


def nw_fast(seq1, seq2):
    lib = c_lib()
    if lib is None:
        return nw(seq1, seq2)
    m, n = len(seq1), len(seq2)
    if m == 0 or n == 0:
        raise NameError("nw: empty sequence (the reference raises NameError)")
    out = (ctypes.c_int * 6)()
    a1 = ctypes.create_string_buffer(m + n + 2)
    a2 = ctypes.create_string_buffer(m + n + 2)
    rc = lib.oracle_nw(seq1.encode(), m, seq2.encode(), n, out, a1, a2)
    if rc != 0:
        raise MemoryError("oracle_nw failed")
    alen = out[5]
    return (a1.raw[:alen].decode(), a2.raw[:alen].decode(), out[0], out[1], out[2], out[3], out[4])
