"""breakmer_b200 -- B200-native implementation of BreaKmer's per-target k-mer
assembly hot path, behind the reference's own Python API.

Reference call site                       drop-in here
----------------------------------------  -----------------------------------------
olc.nw(seq1, seq2)                        breakmer_b200.olc.nw
utils.run_jellyfish / utils.load_kmers    breakmer_b200.utils.run_jellyfish / load_kmers
sv_assembly.init_assembly(...)            breakmer_b200.sv_assembly.init_assembly
target.compare_kmers()                    breakmer_b200.sv_processor.compare_kmers(target)
(region loop, sv_processor.py:185-201)    breakmer_b200.sv_processor.compare_kmers_batch(targets)

All of them run hand-written sm_100a CUDA through the C ABI in
include/breakmer_b200.h (ctypes binding: breakmer_b200._lib).  There is no CPU
fallback: importing this package is free, but the first call needs the built
library and a CUDA device and raises otherwise.
"""
import threading

__all__ = ["get_handle", "close_handles"]

_handles = {}
_lock = threading.Lock()


def get_handle(device=0):
    """The calling THREAD's handle (one CUDA device + one stream) for `device`.  The C ABI requires that calls on one
    handle never overlap (every call reuses the handle's arenas, and results live there until the next call), and ctypes
    releases the GIL during a call, so handles are never shared between threads."""
    from . import _lib
    key = (device, threading.get_ident())
    with _lock:
        h = _handles.get(key)
        if h is None:
            h = _lib.Handle(device)
            _handles[key] = h
        return h


def close_handles():
    with _lock:
        for h in _handles.values():
            h.close()
        _handles.clear()
