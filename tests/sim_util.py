"""Driver for tests/sim (host-compiled single-lane build of the device assembler's
control logic; a debugging aid for the container without a GPU -- see
tests/sim/sim_assemble.cpp).  Not used by the product."""
import ctypes
import os
import subprocess

import numpy as np

from oracle import assembler_py

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_DIR = os.path.join(HERE, "sim")
ORDER_NAMES = ["for", "rev", "mid"]
_BASES = "ACGT"


CSRC = os.path.join(HERE, "..", "breakmer_b200", "csrc")
GEN_DIR = os.path.join(SIM_DIR, "_gen")


def _csrc_files():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]


def generate_simt_sources():
    """tests/sim/_gen: the product sources with their launch statements rewritten for the emulator (gen_simt_sources.py)"""
    import importlib.util
    gen = os.path.join(SIM_DIR, "gen_simt_sources.py")
    stamp = os.path.join(GEN_DIR, "all.cpp")
    if not os.path.isfile(stamp) or any(os.path.getmtime(d) > os.path.getmtime(stamp) for d in _csrc_files() + [gen]):
        spec = importlib.util.spec_from_file_location("gen_simt_sources", gen)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.main(GEN_DIR)
    return GEN_DIR


def build_simt(so_name, src_name, extra=()):
    """an emulator build (tests/sim/simt_host.h + cuda_runtime.h shim) of one harness under tests/sim"""
    gen_dir = generate_simt_sources()
    so = os.path.join(SIM_DIR, so_name)
    src = os.path.join(gen_dir, src_name) if src_name == "all.cpp" else os.path.join(SIM_DIR, src_name)
    deps = [src, os.path.join(SIM_DIR, "simt_host.h"), os.path.join(SIM_DIR, "cuda_runtime.h"), os.path.join(gen_dir, "all.cpp")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        flags = ["-O1", "-g"] if os.environ.get("SIMT_DEBUG") else ["-O2"]
        subprocess.check_call(["g++", "-std=c++17", "-fPIC", "-shared", "-w", "-DBK_SIMT"] + flags + list(extra) +
                              ["-I", SIM_DIR, "-I", gen_dir, "-o", so, src, "-lpthread"])
    return so


def build_tsan_driver():
    """tests/sim/simt_tsan_driver: the whole emulated library in one executable built with -fsanitize=thread; every
    emulated GPU thread is a TSan fiber and the kernels' barriers are the only happens-before edges"""
    gen_dir = generate_simt_sources()
    exe = os.path.join(SIM_DIR, "simt_tsan_driver")
    src = os.path.join(SIM_DIR, "simt_tsan_driver.cpp")
    deps = [src, os.path.join(SIM_DIR, "simt_host.h"), os.path.join(SIM_DIR, "cuda_runtime.h"), os.path.join(gen_dir, "all.cpp")]
    if not os.path.isfile(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-w", "-DBK_SIMT", "-DSIMT_TSAN", "-fsanitize=thread",
                               "-I", SIM_DIR, "-I", gen_dir, "-o", exe, src, "-lpthread"])
    return exe


def build_asan_driver():
    """tests/sim/simt_asan_driver: the same executable under AddressSanitizer + UBSan, with every arena allocation of the
    library turned into an exact-size malloc (-DSIMT_ARENA_MALLOC, patched in by gen_simt_sources.py), so that a kernel
    touching one element past ANY device or pinned array is reported with its source line"""
    gen_dir = generate_simt_sources()
    exe = os.path.join(SIM_DIR, "simt_asan_driver")
    src = os.path.join(SIM_DIR, "simt_tsan_driver.cpp")
    deps = [src, os.path.join(SIM_DIR, "simt_host.h"), os.path.join(SIM_DIR, "cuda_runtime.h"), os.path.join(gen_dir, "all.cpp")]
    if not os.path.isfile(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-w", "-DBK_SIMT", "-DSIMT_ARENA_MALLOC", "-fsanitize=address,undefined",
                               "-I", SIM_DIR, "-I", gen_dir, "-o", exe, src, "-lpthread"])
    return exe


def build(asan=False, simt=False):
    """libsim.so: single-lane build of the control logic; libsimt_asm.so (simt=True): assemble_kernel itself, W warps
    of 32 lanes, on the fiber emulator of tests/sim/simt_host.h"""
    if simt:
        return build_simt("libsimt_asm.so", "sim_assemble.cpp")
    name = "libsim_asan.so" if asan else "libsim.so"
    so = os.path.join(SIM_DIR, name)
    src = os.path.join(SIM_DIR, "sim_assemble.cpp")
    deps = [src, os.path.join(SIM_DIR, "simt_host.h")] + \
           [os.path.join(HERE, "..", "breakmer_b200", "csrc", f) for f in ("assemble.cuh", "nw.cuh", "common.cuh")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        flags = ["-O1", "-g", "-fsanitize=address,undefined"] if asan else ["-O2"]
        subprocess.check_call(["g++", "-std=c++17", "-fPIC", "-shared"] + flags + ["-o", so, src])
    return so


def code_to_mer(code, k):
    code = int(code)
    return "".join(_BASES[(code >> (2 * (k - 1 - i))) & 3] for i in range(k))


def mer_to_code(mer):
    v = 0
    for c in mer:
        v = (v << 2) | "ACGT".index(c)
    return v


def decode_contigs(n_ctg, desc, o_seq, o_locs, o_io, o_ot, o_reads, o_mer, o_pos, o_meta, k, rec_ids):
    """Device/sim output arena -> canonical contig records (oracle.contig_record shape).
    desc rows: region, ordinal, seq_off, seq_len, cnt_off, cnt_len, reads_off, n_reads, kmers_off, n_kmers."""
    rows = sorted((tuple(int(v) for v in desc[i * 10:(i + 1) * 10]) for i in range(n_ctg)), key=lambda r: (r[0], r[1]))
    out = []
    for (_reg, _ordn, so, sl, co, cl, ro, nr, ko, nk) in rows:
        kmers = []
        for e in range(ko, ko + nk):
            meta = int(o_meta[e])
            kmers.append([code_to_mer(o_mer[e], k), int(o_pos[e]), meta & 1, meta >> 3, ORDER_NAMES[(meta >> 1) & 3]])
        out.append({
            "seq": bytes(o_seq[so:so + sl]).decode(),
            "indel_only": [int(v) for v in o_io[co:co + cl]],
            "others": [int(v) for v in o_ot[co:co + cl]],
            "reads": sorted(rec_ids[int(r)] for r in o_reads[ro:ro + nr]),
            "kmers": kmers,
            "kmer_locs": [int(v) for v in o_locs[so:so + sl]],
        })
    return out


def sim_init_assembly(mers, records, k, rc_thresh, read_len, asan=False, cap=1 << 23, spec_w=4, simt=False, score_table=True):
    lib = ctypes.CDLL(build(asan, simt))
    if simt:
        lib.sim_use_score_table(ctypes.c_int(1 if score_table else 0))
    uniq = assembler_py.group_reads(records)
    seqs = [u.seq.encode() for u in uniq]
    roff = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=roff[1:])
    rb = np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8).copy()
    mult = np.array([u.nreads for u in uniq] + [0], dtype=np.uint32)
    io = np.array([1 if u.indel_only else 0 for u in uniq] + [0], dtype=np.uint8)
    items = sorted((mer_to_code(m), c) for m, c in mers.items())
    mc = np.array([m for m, _ in items] + [0], dtype=np.uint64)
    cc = np.array([c for _, c in items] + [0], dtype=np.uint32)
    o_seq = np.zeros(cap, np.uint8); o_locs = np.zeros(cap, np.int32)
    o_io = np.zeros(cap, np.int32); o_ot = np.zeros(cap, np.int32); o_reads = np.zeros(cap, np.int32)
    o_mer = np.zeros(cap, np.uint64); o_pos = np.zeros(cap, np.int32); o_meta = np.zeros(cap, np.int32)
    desc = np.zeros(cap, np.int64)
    n_ctg = ctypes.c_int64()
    stats = np.zeros(4, np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.sim_assemble_region.restype = ctypes.c_int
    rc = lib.sim_assemble_region(p(rb), p(roff), ctypes.c_int(len(uniq)), p(mult), p(io), p(mc), p(cc),
                                 ctypes.c_int(len(items)), ctypes.c_int(k), ctypes.c_int(rc_thresh), ctypes.c_int(read_len),
                                 ctypes.c_int(spec_w), ctypes.c_int64(cap // 10), p(o_seq), p(o_locs), p(o_io), p(o_ot), p(o_reads),
                                 p(o_mer), p(o_pos), p(o_meta), p(desc), ctypes.byref(n_ctg), p(stats))
    if rc != 0:
        raise RuntimeError("sim status %d" % rc)
    rec_ids = [u.rep_id for u in uniq]
    out = decode_contigs(n_ctg.value, desc, o_seq, o_locs, o_io, o_ot, o_reads, o_mer, o_pos, o_meta, k, rec_ids)
    return out, {"check_align": int(stats[0]), "cells": int(stats[1]), "find_reads": int(stats[2]), "seeds": int(stats[3])}


if __name__ == "__main__":
    so = build_simt("libbreakmer_simt_TESTONLY.so", "all.cpp")
    print("built", so)
    print("the GPU tests on the emulator:  BK_LIB=%s python -m pytest tests -m gpu -q --deselect tests/test_bench_contract.py" % so)
