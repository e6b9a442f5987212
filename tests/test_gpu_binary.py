"""GPU tests of the drop-in binary bioEM_b200 (reference command line and file formats on top of
the C ABI): run it on the synthetic case files and compare the text outputs with the committed
outputs of the unmodified reference (tests/golden), line by line."""
import os
import subprocess

import numpy as np
import pytest

from bioem_b200 import api
from bioem_b200.cases import build_case, reference_cli
from bioem_b200.outputs import parse_output_probabilities

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EXE = os.path.join(ROOT, "bioem_b200", "bin", "bioEM_b200")
LOGP_ATOL = {26: 5e-3, 32: 5e-3, 33: 5e-3, 35: 5e-3, 64: 2e-2, 128: 5e-2, 224: 0.3, 360: 1.0}


def _run(name, tmp_path, env=None, **overrides):
    if api.lib().bioem_b200_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "bioem_b200", "csrc", "host")])
    cd = build_case(name, str(tmp_path), **overrides)
    r = subprocess.run([EXE] + reference_cli(cd), cwd=tmp_path, capture_output=True, text=True,
                       env={**os.environ, **(env or {})})
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    return cd


@pytest.mark.parametrize("name", ["toy32", "toy32psf", "toy32opts", "toy32euler", "toy32pts", "toy32clip", "toy32amp", "toy32g2odd", "toy32g3", "toy33", "toy35g2", "toy26",
                                  "toy64", "cfg1", "cfg2_slice", "cfg4_voxel_slice"])
def test_binary_output_probabilities_match_reference_golden(name, tmp_path, golden_dir):
    cd = _run(name, tmp_path)
    got_path = tmp_path / "Output_Probabilities"
    ref_path = os.path.join(golden_dir, name, "Output_Probabilities")
    got_lines = open(got_path, encoding="utf-8").read().split("\n")
    ref_lines = open(ref_path, encoding="utf-8").read().split("\n")
    assert len(got_lines) == len(ref_lines)
    # header: byte-identical
    hdr = next(i for i, ln in enumerate(ref_lines) if ln.startswith("RefMap:"))
    assert got_lines[:hdr] == ref_lines[:hdr]
    # every data line has the same tokens in the same places (units, brackets, trailing blank)
    for g, r in zip(got_lines[hdr:], ref_lines[hdr:]):
        gt, rt = g.split(" "), r.split(" ")
        assert len(gt) == len(rt), (g, r)
        for a, b in zip(gt, rt):
            if not _is_number(b):
                assert a == b, (g, r)
    quat = not cd.case.euler_grid  # Euler grids print alpha / beta / gamma (three angle columns)
    got, ref = parse_output_probabilities(str(got_path), quat), parse_output_probabilities(ref_path, quat)
    n = cd.case.n_pixels
    same = 0
    for m in range(len(ref["logp"])):
        assert abs(got["logp"][m] - ref["logp"][m]) <= LOGP_ATOL[n] + 1e-4
        assert abs(got["logp"][m] - ref["logp"][m]) <= 1e-4 * abs(ref["logp"][m])
        if (got["cent_x"][m] == ref["cent_x"][m] and got["cent_y"][m] == ref["cent_y"][m]
                and np.allclose(got["angles"][m], ref["angles"][m], atol=1.1e-4)
                and abs(got["defocus"][m] - ref["defocus"][m]) < 1.1e-4 and abs(got["env"][m] - ref["env"][m]) < 1.1e-4):
            same += 1
            assert abs(got["norm"][m] - ref["norm"][m]) <= 1e-3 * abs(ref["norm"][m]) + 2e-4
            assert abs(got["mu"][m] - ref["mu"][m]) <= 1e-3 * abs(ref["mu"][m]) + 2e-4
    # near-ties are analysed against the oracle in test_gpu_parity.py; here most images must agree
    assert same >= len(ref["logp"]) - max(1, len(ref["logp"]) // 3)


def _is_number(tok):
    try:
        float(tok)
        return True
    except ValueError:
        return False


@pytest.mark.parametrize("name", ["toy32", "toy32opts", "toy32euler"])
def test_binary_ang_prob_matches_reference_golden(name, tmp_path, golden_dir):
    """WRITE_PROB_ANGLES: same header, same number of rows per image, same top orientation and
    log-probabilities within tolerance (the order further down the list may swap on near-ties).
    toy32opts adds PRIOR_ANGLES: the rows are ordered without the prior, which is printed and added last."""
    _run(name, tmp_path)
    got = open(tmp_path / "ANG_PROB").read().split("\n")
    ref = open(os.path.join(golden_dir, name, "ANG_PROB")).read().split("\n")
    assert got[:3] == ref[:3] and len(got) == len(ref)
    g = np.array([[float(x) for x in ln.replace("Separated:", "").split()] for ln in got[3:] if ln.strip()])
    r = np.array([[float(x) for x in ln.replace("Separated:", "").split()] for ln in ref[3:] if ln.strip()])
    assert g.shape == r.shape
    na = 3 if name == "toy32euler" else 4  # angle columns: alpha beta gamma, or q1..q4
    lp = 1 + na                              # then logP, three "Separated:" terms and, with PRIOR_ANGLES, the prior
    for m in np.unique(r[:, 0]):
        gm, rm = g[g[:, 0] == m], r[r[:, 0] == m]
        has_prior = gm.shape[1] > lp + 4
        key = gm[:, lp] - (gm[:, lp + 4] if has_prior else 0.0)
        assert np.all(np.diff(key) <= 1e-3)  # descending log-probability (before the angle prior, 4 decimals)
        if has_prior:
            np.testing.assert_array_equal(gm[:, lp + 4], rm[:, lp + 4])
        np.testing.assert_allclose(gm[:, lp], rm[:, lp], atol=5e-3 + 1e-4)
        # same set of orientations up to swaps between near-equal entries
        gs = {tuple(np.round(x, 3)) for x in gm[:, 1:lp]}
        rs = {tuple(np.round(x, 3)) for x in rm[:, 1:lp]}
        assert len(gs & rs) >= len(rs) - 1


def test_binary_multi_gpu_split_equals_single(tmp_path):
    """--Gpus / BIOEM_B200_GPUS: the orientation grid split into blocks, one handle per block, merged ON THE DEVICE
    (bioem_b200_merge_peers / _merge_top_angles_peers).  On a box with fewer GPUs than blocks the blocks share the
    GPUs (BIOEM_B200_OVERSUBSCRIBE=1), so the multi-GPU code path is exercised on every box."""
    one = tmp_path / "one"
    one.mkdir()
    _run("toy64", one, env={"BIOEM_B200_GPUS": "1"})
    a = parse_output_probabilities(str(one / "Output_Probabilities"))
    for nblocks in (2, 3):
        two = tmp_path / f"b{nblocks}"
        two.mkdir()
        _run("toy64", two, env={"BIOEM_B200_GPUS": str(nblocks), "BIOEM_B200_OVERSUBSCRIBE": "1"})
        b = parse_output_probabilities(str(two / "Output_Probabilities"))
        np.testing.assert_allclose(a["logp"], b["logp"], atol=2e-4)
        for k in ("cent_x", "cent_y", "angles", "env", "defocus"):
            np.testing.assert_array_equal(a[k], b[k])
        # WRITE_PROB_ANGLES: every block keeps its most probable orientations, merged on the device
        o1, o2 = tmp_path / f"ang1_{nblocks}", tmp_path / f"ang{nblocks}"
        o1.mkdir()
        o2.mkdir()
        _run("toy32", o1, env={"BIOEM_B200_GPUS": "1"})
        _run("toy32", o2, env={"BIOEM_B200_GPUS": str(nblocks), "BIOEM_B200_OVERSUBSCRIBE": "1"})
        assert open(o1 / "ANG_PROB").read() == open(o2 / "ANG_PROB").read()
        assert open(o1 / "Output_Probabilities").read() == open(o2 / "Output_Probabilities").read()


def test_binary_warns_about_model_points_out_of_frame(tmp_path):
    """reference bioem.cpp:1724-1734,1756-1780: "point out of image size" once per projection that loses points"""
    if api.lib().bioem_b200_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    cd = build_case("toy32clip", str(tmp_path))
    r = subprocess.run([EXE] + reference_cli(cd), cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0
    warn = [ln for ln in r.stdout.splitlines() if "point out of image size" in ln]
    assert 1 <= len(warn) <= cd.case.n_orient
    cd = build_case("toy32", str(tmp_path / "ok"))
    r = subprocess.run([EXE] + reference_cli(cd), cwd=tmp_path / "ok", capture_output=True, text=True)
    assert r.returncode == 0 and "point out of image size" not in r.stdout


REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "bioEM_ref_cuda")


@pytest.mark.parametrize("name,overrides", [("toy64", {}), ("cfg2_slice", {}),
                                            # the headline shape: all 1000 particles of cfg2 (224 x 224, 32 CTFs, window 81 x 81)
                                            ("cfg2", dict(n_particles=1000, n_orient=4))])
def test_binary_agrees_with_reference_cuda_path_on_this_gpu(name, overrides, tmp_path):
    """The reference's own CUDA path (bioem_cuda.cu + cuFFT, rebuilt for sm_100a by oracle/Makefile) run on
    the same box and the same files: same maximizing orientation / CTF / displacement, log P within 1e-4
    relative.  The checker is executed, never linked: skipped when it was not built."""
    if not os.path.exists(REF_CUDA):
        pytest.skip("oracle/_ref/bioEM_ref_cuda not built")
    cd = _run(name, tmp_path, **overrides)
    os.rename(tmp_path / "Output_Probabilities", tmp_path / "ours")
    r = subprocess.run([REF_CUDA] + reference_cli(cd), cwd=tmp_path, capture_output=True, text=True, timeout=600,
                       env={**os.environ, "GPU": "1", "GPUWORKLOAD": "100", "GPUDEVICE": "0"})
    if r.returncode != 0:
        pytest.skip("reference CUDA build does not run on this box: " + (r.stdout + r.stderr)[-300:])
    got, ref = parse_output_probabilities(str(tmp_path / "ours")), parse_output_probabilities(
        str(tmp_path / "Output_Probabilities"))
    same = 0
    for m in range(len(ref["logp"])):
        assert abs(got["logp"][m] - ref["logp"][m]) <= 1e-4 * abs(ref["logp"][m])
        same += (got["cent_x"][m] == ref["cent_x"][m] and got["cent_y"][m] == ref["cent_y"][m]
                 and np.allclose(got["angles"][m], ref["angles"][m], atol=1.1e-4)
                 and abs(got["defocus"][m] - ref["defocus"][m]) < 1.1e-4 and abs(got["env"][m] - ref["env"][m]) < 1.1e-4)
    # cuFFT + --use_fast_math round differently from FFTW: a near-tie may flip for one image
    assert same >= len(ref["logp"]) - max(1, len(ref["logp"]) // 3)
