"""CPU-only: the C-ABI library loads, exports every symbol include/bioem_b200.h declares, and
its host-side input preparation agrees bit for bit with the oracle's restatement of the
reference (no device compute is called here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bioem_b200 import api
from bioem_b200.cases import CASES, build_case
from oracle import pyoracle

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bioem_b200.h")).read()
    declared = set(re.findall(r"\b(bioem_b200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"bioem_b200_context"}
    L = api.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    assert L.bioem_b200_version() >= 100


def test_supported_sizes():
    L = api.lib()
    for n in (32, 36, 64, 128, 224, 360):
        assert L.bioem_b200_supported_size(n) == 1
    for n in (31, 100, 225, 1024):
        assert L.bioem_b200_supported_size(n) == 0


def test_create_fails_loudly_without_device_or_with_bad_config():
    L = api.lib()
    cfg = api.Config(100, 4, 1, 0, 0, 1, 0, 0, 1.0, 1e4, 1.0, 1, 1, 1, 1, 0)
    h = C.c_void_p()
    assert L.bioem_b200_create(C.byref(cfg), 0, C.byref(h)) == 1  # unsupported size
    assert b"NUMBER_PIXELS" in L.bioem_b200_last_error()
    cfg = api.Config(64, 5, 2, 0, 0, 1, 0, 0, 1.0, 4096, 1.0, 1, 1, 1, 1, 0)
    assert L.bioem_b200_create(C.byref(cfg), 0, C.byref(h)) == 1  # maxD % G != 0 (quirk Q3)
    if L.bioem_b200_device_count() == 0:
        cfg = api.Config(64, 4, 1, 0, 0, 1, 0, 0, 1.0, 4096, 1.0, 1, 1, 1, 1, 0)
        rc = L.bioem_b200_create(C.byref(cfg), 0, C.byref(h))
        assert rc == 2 and b"no CPU fallback" in L.bioem_b200_last_error()


@pytest.mark.parametrize("name", ["toy32", "toy36g2", "toy64", "cfg1", "cfg2_slice"])
def test_host_preparation_matches_oracle(name):
    cd = build_case(name)
    hi, parts = api.inputs_for_case(cd)
    P = pyoracle.Prepared(cd.case, cd.model, cd.quats, cd.particles)
    for f, _ in api.Config._fields_:
        a, b = getattr(hi.cfg, f), getattr(P.cfg, f)
        assert a == b, (f, a, b)
    assert hi.C == P.C and hi.O == P.O
    np.testing.assert_array_equal(hi.CtfParam[:, :3], P.CtfParam)
    np.testing.assert_array_equal(hi.refCTF, P.refCTF)
    assert hi.NormDen == P.NormDen
    np.testing.assert_array_equal(hi.points["pos"], P.pts[:, :3])
    np.testing.assert_array_equal(parts, P.maps)


def test_ctf_mirror_row_quirk_q1():
    """rows r <= N/2-2 use |k|=r, rows N/2-1 and N/2 use N/2, rows r > N/2 use N-1-r."""
    n = 16
    ref = np.zeros((1, n * (n // 2 + 1), 2), np.float32)
    par = np.zeros((1, 4), np.float32)
    g = np.zeros(3, np.float32)
    api.lib().bioem_b200_host_ctf_table(n, 1.5, 0, 0.1, 0.1, 1, 2.0, 2.0, 1, 50.0, 50.0, 1, api._fp(ref),
                                        api._fp(par), api._fp(g))
    t = ref[0, :, 0].reshape(n, n // 2 + 1)
    assert np.array_equal(t[n // 2 - 1], t[n // 2])
    for r in range(n // 2 + 1, n):
        assert np.array_equal(t[r], t[n - 1 - r])
    assert (ref[..., 1] == 0).all() and t[0, 0] == 1.0


def test_final_logprob_and_merge_host():
    cd = build_case("toy32")
    hi, _ = api.inputs_for_case(cd)
    P = pyoracle.Prepared(cd.case, cd.model, cd.quats, cd.particles)
    assert hi.final_logprob(3.5, -1234.5) == P.final_logprob(3.5, -1234.5)
    # split an oracle run into two orientation blocks and merge: equals the single run
    full = P.run()["prob"]
    a = P.run(0, 11)["prob"]
    b = P.run(11, P.O)["prob"]
    merged = api.merge_host(np.stack([a, b]))
    np.testing.assert_allclose(merged["Total"], full["Total"], rtol=1e-12)
    for k in ("Constoadd", "cent_x", "cent_y", "orient", "conv", "norm", "mu"):
        np.testing.assert_array_equal(merged[k], full[k])
