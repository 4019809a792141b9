"""The WHOLE library on the host SIMT emulator: breakmer_b200/csrc compiled by g++ with tests/sim/cuda_runtime.h standing
in for the CUDA runtime ("device" memory = host memory, copies = memcpy) and every kernel launch executed by the fiber
emulator of tests/sim/simt_host.h (tests/sim/gen_simt_sources.py rewrites the launch statements, nothing else).  The
result is a test-only shared library with the product's C ABI; the GPU tests of this repo run against it unchanged through
the BK_LIB override of breakmer_b200/_lib.py -- host pipeline, every kernel, the Python drop-ins -- in a container without
a GPU.  It is a check of the SOURCE (logic, indexing, races at warp / block level); parity of the compiled sm_100a
library is what `pytest -m gpu` on the B200 establishes.  Nothing in breakmer_b200/ knows about this library.

The CPU suite runs a subset sized for about a minute; the complete GPU suite passes on the emulator as well
(`python tests/sim_util.py` builds the library and prints the command; test_gpu_pipeline.py takes 5 minutes,
test_gpu_full_configs.py about an hour)."""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import pytest

import sim_util
from conftest import ROOT

SUBSETS = [
    (["tests/test_gpu_kmers.py", "tests/test_k1_jellyfish_cases.py"], None, 27),
    (["tests/test_gpu_pipeline.py"],
     "other_k or empty_and_ragged or given_mers or capacity_is_reported or odd_inputs", 7),
    (["tests/test_gpu_api.py", "tests/test_handoff.py", "tests/test_ingest.py"],
     "not sharded and not applying_thread and not ingested_batch_equals", 8),
    (["tests/test_gpu_nw.py", "tests/test_gpu_redundancy.py"], "golden or incremental or signature or above_4095", 5),
]


@pytest.fixture(scope="module")
def library():
    return sim_util.build_simt("libbreakmer_simt_TESTONLY.so", "all.cpp")


def _run(library, files, expr):
    env = dict(os.environ, BK_LIB=library)
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + files + (["-k", expr] if expr else [])
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    m = re.search(r"(\d+) passed", out.stdout)
    return out.returncode, int(m.group(1)) if m else 0, out.stdout[-1500:] + out.stderr[-500:]


def test_gpu_tests_pass_on_the_emulated_library(library):
    with ThreadPoolExecutor(max_workers=len(SUBSETS)) as ex:
        results = list(ex.map(lambda s: _run(library, s[0], s[1]), SUBSETS))
    for (files, expr, at_least), (rc, n_passed, tail) in zip(SUBSETS, results):
        assert rc == 0, (files, tail)
        assert n_passed >= at_least, (files, expr, n_passed, tail)
