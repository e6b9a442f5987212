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
