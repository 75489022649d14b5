"""Pins bh_ic_refdisk (csrc/bh_ic.cpp) — the input of the headline benchmark, BASELINE.json configs[1] — against the
UNMODIFIED reference's own initial conditions, nbody_v5_bench.cu:294-308.

oracle/ref_wrap.cu:ref_ic runs the reference's main() up to its seven uploads (bench:329-335) and hands back the
host arrays; tests/golden/reference_ic_sha256.json holds their digests (written by
`python tests/golden/make_golden.py reference_ic`), so the pin also holds where /root/reference and oracle/_ref are
absent.  Byte equality, every array, every body."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(O.ROOT, "tests", "golden", "reference_ic_sha256.json")
REF = os.path.join(O.ROOT, "oracle", "_ref", "libref_step.so")


def _golden():
    with open(GOLD) as f:
        return json.load(f)


@pytest.mark.parametrize("n", [1000, 16384, 500_000, 1_000_000])
def test_refdisk_generator_matches_the_reference_digest(bh, n):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(O.ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    assert mg.ic_digest(bh.ic_refdisk(n, 42)) == _golden()["sha256"][str(n)]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("n", [1, 2, 777, 500_000])
def test_refdisk_generator_equals_the_reference_main_bit_for_bit(bh, n):
    import ctypes as C

    L = C.CDLL(REF)
    want = [np.zeros(n, np.float32) for _ in range(7)]
    assert L.ref_ic(n, *[x.ctypes.data_as(C.c_void_p) for x in want]) == 0
    got = bh.ic_refdisk(n, 42)
    for name, g, w in zip(["posX", "posY", "posZ", "velX", "velY", "velZ", "mass"], got, want):
        assert g.dtype == np.float32 and g.tobytes() == w.tobytes(), name
    # and the digest file was made from the same generator
    if str(n) in _golden()["sha256"]:
        import hashlib

        assert hashlib.sha256(b"".join(w.tobytes() for w in want)).hexdigest() == _golden()["sha256"][str(n)]
