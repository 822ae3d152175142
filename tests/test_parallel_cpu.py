"""Host-side sharding logic of the multi-GPU path, exercised with world_size=2 on the gloo backend (CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rel_pose_b200 import parallel as P


def test_shard_ranges_partition_every_size():
    for n in (0, 1, 2, 7, 64, 4096, 4099):
        for w in (1, 2, 3, 4, 8):
            spans = [P.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_range(4, 2, 2)


def test_micro_batches_cover_shard():
    assert P.micro_batches(3, 3, 4) == []
    assert P.micro_batches(0, 10, 4) == [(0, 4), (4, 8), (8, 10)]
    lo, hi = P.shard_range(4096, 8, 5)
    mb = P.micro_batches(lo, hi, 64)
    assert mb[0][0] == lo and mb[-1][1] == hi and len(mb) == 8
    with pytest.raises(ValueError):
        P.micro_batches(0, 4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_forward(pair_ids):
    """Stand-in for the per-pair forward: a deterministic function of the pair index only, so that
    sharded == unsharded proves the partition/gather logic (the CUDA forward itself is covered on the GPU)."""
    x = pair_ids.double()[:, None, None]
    j = torch.arange(7, dtype=torch.float64)[None, None, :]
    v = torch.arange(2, dtype=torch.float64)[None, :, None]
    return torch.sin(x * 0.37 + j * 1.3 + v * 0.11).float()


def _worker(rank, world, port, n_pairs, micro, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = P.shard_range(n_pairs, world, rank)
        outs = [_fake_forward(torch.arange(a, b)) for a, b in P.micro_batches(lo, hi, micro)]
        local = torch.cat(outs, 0) if outs else torch.zeros((0, 2, 7))
        full = P.gather_poses(local, n_pairs)
        # barrier + max-over-ranks reduction of a timing value, as bench.py does
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, full, float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs,micro", [(10, 3), (7, 2), (1, 4)])
def test_gloo_world2_sharded_equals_unsharded(n_pairs, micro):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, micro, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=420) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _fake_forward(torch.arange(n_pairs))
    for rank, full, tmax in res:
        assert full.shape == (n_pairs, 2, 7)
        assert torch.equal(full, ref)
        assert tmax == 2.0


def test_gather_without_process_group_is_identity():
    x = torch.randn(3, 2, 7)
    assert P.gather_poses(x, 3) is x


def test_streamed_inference_refuses_cpu_model():
    m = torch.nn.Linear(2, 2)
    with pytest.raises(RuntimeError):
        P.StreamedInference(m)
