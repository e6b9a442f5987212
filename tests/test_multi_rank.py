"""CPU, world_size 2, gloo: the orientation sharding and the merge of per-rank results that
bench.py / the multi-GPU run use (reference MPI split bioem.cpp:748-753, reduction :909-1044).
Each rank evaluates its block of orientations with the CPU oracle (the checker — there is no GPU
here), the per-image records travel through one all_gather, and the library's host merge
(bioem_b200_merge_host, same rule as the device merge kernel) must reproduce the single-process
result: identical arg-max records, lowest rank wins ties, log-sum-exp equal."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bioem_b200 import api
    from bioem_b200.cases import build_case
    from oracle import pyoracle
    cd = build_case("toy32")
    P = pyoracle.Prepared(cd.case, cd.model, cd.quats, cd.particles)
    o_lo, o_hi = rank * P.O // world, (rank + 1) * P.O // world  # same split as bench.py
    mine = P.run(o_lo, o_hi)["prob"]
    assert mine.dtype == api.PROB_MAP_DTYPE
    t = torch.from_numpy(mine.view(np.uint8).copy())
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    parts = np.stack([g.numpy().view(api.PROB_MAP_DTYPE) for g in gathered])
    merged = api.merge_host(parts)
    if rank == 0:
        np.save(out_path, merged)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_oracle_blocks_merge_to_the_single_process_result(tmp_path):
    sys.path.insert(0, ROOT)
    from bioem_b200.cases import build_case
    from oracle import pyoracle
    out = str(tmp_path / "merged.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    merged = np.load(out)
    cd = build_case("toy32")
    P = pyoracle.Prepared(cd.case, cd.model, cd.quats, cd.particles)
    full = P.run()["prob"]
    for k in ("Constoadd", "cent_x", "cent_y", "orient", "conv", "norm", "mu"):
        np.testing.assert_array_equal(merged[k], full[k])
    np.testing.assert_allclose(merged["Total"], full["Total"], rtol=1e-12)


def test_merge_prefers_the_lowest_rank_on_ties():
    sys.path.insert(0, ROOT)
    from bioem_b200 import api
    a = np.zeros((3, 2), dtype=api.PROB_MAP_DTYPE)
    a["Constoadd"] = [[-10.0, -5.0], [-10.0, -7.0], [-12.0, -5.0]]
    a["Total"] = [[1.0, 2.0], [3.0, 1.0], [5.0, 4.0]]
    a["orient"] = [[0, 1], [10, 11], [20, 21]]
    m = api.merge_host(a)
    # image 0: ranks 0 and 1 tie at -10 -> rank 0's record; image 1: ranks 0 and 2 tie at -5 -> rank 0
    assert list(m["orient"]) == [0, 1]
    np.testing.assert_allclose(m["Total"][0], 1.0 + 3.0 + 5.0 * np.exp(-2.0))
    np.testing.assert_allclose(m["Total"][1], 2.0 + 1.0 * np.exp(-2.0) + 4.0)


# ---------------------------------------------------------------------------------------------
# Size-independent properties of the rank merge (bioem_b200_merge_host, the host twin of
# merge_partials_kernel): merging in blocks equals merging at once, log-sum-exp is preserved, and
# the arg-max record comes from the lowest rank among equal maxima.
from hypothesis import given, settings, strategies as st  # noqa: E402

from bioem_b200 import api as _api  # noqa: E402


def _random_parts(rng, n_ranks, m):
    p = np.zeros((n_ranks, m), dtype=_api.PROB_MAP_DTYPE)
    p["Constoadd"] = np.round(rng.uniform(-80000.0, -60000.0, size=(n_ranks, m)), 3)
    p["Total"] = rng.uniform(1.0, 50.0, size=(n_ranks, m))
    p["cent_x"] = rng.integers(-40, 41, size=(n_ranks, m))
    p["cent_y"] = rng.integers(-40, 41, size=(n_ranks, m))
    p["orient"] = np.arange(n_ranks)[:, None] * 1000 + rng.integers(0, 1000, size=(n_ranks, m))
    p["conv"] = rng.integers(0, 32, size=(n_ranks, m))
    p["norm"] = rng.uniform(0.5, 2.0, size=(n_ranks, m)).astype(np.float32)
    p["mu"] = rng.normal(size=(n_ranks, m)).astype(np.float32)
    return p


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n_ranks=st.integers(1, 8), m=st.integers(1, 17), cut=st.integers(0, 8),
       tie=st.booleans())
def test_merge_host_properties(seed, n_ranks, m, cut, tie):
    rng = np.random.default_rng(seed)
    p = _random_parts(rng, n_ranks, m)
    if tie and n_ranks > 1:  # two ranks share the maximum of image 0
        p["Constoadd"][n_ranks - 1, 0] = p["Constoadd"][0, 0] = p["Constoadd"][:, 0].max() + 1.0
    all_at_once = _api.merge_host(p)
    # log-sum-exp is preserved
    cmax = p["Constoadd"].max(axis=0)
    want = np.log((p["Total"] * np.exp(p["Constoadd"] - cmax)).sum(axis=0)) + cmax
    got = np.log(all_at_once["Total"]) + all_at_once["Constoadd"]
    np.testing.assert_allclose(got, want, rtol=1e-13)
    np.testing.assert_array_equal(all_at_once["Constoadd"], cmax)
    # the record of the lowest rank holding the maximum
    owner = (p["Constoadd"] == cmax).argmax(axis=0)
    for k in ("cent_x", "cent_y", "orient", "conv", "norm", "mu"):
        np.testing.assert_array_equal(all_at_once[k], p[k][owner, np.arange(m)])
    # merging [0, cut) and [cut, n) first, then the two results, gives the same
    cut = min(max(cut, 1), n_ranks - 1) if n_ranks > 1 else 0
    if n_ranks > 1:
        two = np.stack([_api.merge_host(p[:cut]), _api.merge_host(p[cut:])])
        blocks = _api.merge_host(two)
        np.testing.assert_allclose(np.log(blocks["Total"]) + blocks["Constoadd"], got, rtol=1e-13)
        for k in ("Constoadd", "cent_x", "cent_y", "orient", "conv", "norm", "mu"):
            np.testing.assert_array_equal(blocks[k], all_at_once[k])
