"""N > 1 path on CPU: world_size-2 gloo processes shard a query batch, search their slice (the CPU oracle stands in
for the per-GPU search here — this test is about the sharding / gather logic) and gather on rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seismic_b200.distributed import gather_results, pack_results, shard_bounds, sharded_batch_search, unpack_results


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 10, 10000, 10001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_pack_roundtrip():
    ids = torch.tensor([[1, 2 ** 40 + 5, -1]], dtype=torch.int64)
    sc = torch.tensor([[1.5, -0.0, float("-inf")]])
    cnt = torch.tensor([2], dtype=torch.int32)
    a, b, c = unpack_results(pack_results(ids, sc, cnt), 3)
    assert torch.equal(a, ids) and torch.equal(b.view(torch.int32), sc.view(torch.int32)) and torch.equal(c, cnt)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_docs, n_queries, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from seismic_b200 import Dataset, HostIndex
        cfg = Dataset.synth_config(n_docs, dim=2000)
        index = HostIndex.build(Dataset.synth_documents(cfg), n_postings=300)     # every rank holds a replica
        q = Dataset.synth_queries(cfg, n_queries)

        def search(o, c, v):
            return oracle.batch_search(index.view, o, c, v, 10, 3, 0.8)[:3]
        res = sharded_batch_search(search, q.offsets, q.comps, q.values, 10)
        if rank == 0:
            full = oracle.batch_search(index.view, q.offsets, q.comps, q.values, 10, 3, 0.8)
            ok = all(np.array_equal(a, b) for a, b in zip(res, full[:3]))
            np.save(out_path, np.array([int(ok), len(res[0])]))
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_queries", [(2, 101), (3, 64)])
def test_sharded_search_equals_single_process(tmp_path, world, n_queries):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(world, _free_port(), 4000, n_queries, out), nprocs=world, join=True)
    ok, n = np.load(out)
    assert ok == 1 and n == n_queries
