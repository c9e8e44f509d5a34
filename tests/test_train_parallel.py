"""Gradient bucketing / allreduce of the data-parallel training step on CPU (gloo, world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spoofsv_b200.train import OverlappedGradReducer, allreduce_gradients, plan_buckets, shard_batch, shard_weight


def test_plan_buckets_preserves_order_and_limits():
    sizes = [5, 3, 9, 1, 1, 20, 2]
    b = plan_buckets(sizes, 10)
    assert [i for g in b for i in g] == list(range(len(sizes)))
    assert b == [[0, 1], [2, 3], [4], [5], [6]]
    assert all(sum(sizes[i] for i in g) <= 10 or len(g) == 1 for g in b)
    assert plan_buckets([], 4) == []
    with pytest.raises(ValueError):
        plan_buckets([1], 0)


def test_shard_batch_covers_everything():
    for n, w in [(32, 8), (33, 4), (5, 8), (0, 2)]:
        parts = [shard_batch(n, w, r) for r in range(w)]
        assert parts[0].start == 0 and parts[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
        assert max(p.stop - p.start for p in parts) - min(p.stop - p.start for p in parts) <= 1


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                   # same model on both ranks
        model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        frozen = torch.nn.Parameter(torch.ones(4), requires_grad=False)
        unused = torch.nn.Parameter(torch.ones(6))             # never receives a gradient
        params = list(model.parameters()) + [frozen, unused]
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 7, generator=g), torch.randn(8, 3, generator=g)
        sl = shard_batch(8, world, rank)
        loss = ((model(x[sl]) - y[sl]) ** 2).sum() / 8         # mean over the GLOBAL batch, split by rank
        loss.backward()
        n = allreduce_gradients(params, bucket_mb=1e-4, average=False)     # tiny buckets: several in flight
        ref = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        ref.load_state_dict(model.state_dict())
        (((ref(x) - y) ** 2).sum() / 8).backward()
        err = max(float((a.grad - b.grad).abs().max()) for a, b in zip(model.parameters(), ref.parameters()))
        q.put((rank, n, err, float(unused.grad.abs().max()), frozen.grad is None))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, err, unused_max, frozen_none in out:
        assert n == 7 * 5 + 5 + 5 * 3 + 3 + 6           # every trainable element, the frozen one excluded
        assert err <= 1e-6                              # summed shard gradients == gradient of the global batch
        assert unused_max == 0.0 and frozen_none


def _worker_uneven(rank: int, world: int, port: int, q):
    """A global batch of 5 over 2 ranks (shards of 3 and 2): per-rank MEAN losses weighted by shard_weight, averaged by
    allreduce_gradients, must equal the gradient of the mean over the global batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(5, 7, generator=g), torch.randn(5, 3, generator=g)
        sl = shard_batch(5, world, rank)
        w = shard_weight(5, world, rank)
        (((model(x[sl]) - y[sl]) ** 2).mean() * w).backward()
        allreduce_gradients(model.parameters(), average=True)
        ref = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        ref.load_state_dict(model.state_dict())
        ((ref(x) - y) ** 2).mean().backward()
        err = max(float((a.grad - b.grad).abs().max()) for a, b in zip(model.parameters(), ref.parameters()))
        q.put((rank, w, err))
    finally:
        dist.destroy_process_group()


def test_uneven_shards_are_weighted_to_the_global_mean_gloo():
    assert shard_weight(32, 8, 3) == 1.0 and abs(sum(shard_weight(5, 2, r) for r in range(2)) - 2.0) < 1e-12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_uneven, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [round(w, 6) for _, w, _ in out] == [1.2, 0.8]
    assert all(err <= 1e-6 for _, _, err in out)


def _worker_overlap(rank: int, world: int, port: int, q):
    """The hook-driven reducer (bucket allreduces launched during backward) gives the same gradients as the global
    batch; a parameter that gets no gradient contributes zeros; a second step re-arms cleanly."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(7, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
        unused = torch.nn.Parameter(torch.ones(6))
        params = list(model.parameters()) + [unused]
        red = OverlappedGradReducer(params, bucket_mb=1e-4)            # tiny buckets: several launches from the hooks
        assert len(red.buckets) >= 3
        g = torch.Generator().manual_seed(1)
        errs = []
        for step in range(2):
            x, y = torch.randn(8, 7, generator=g), torch.randn(8, 3, generator=g)
            for p in params:
                p.grad = None
            sl = shard_batch(8, world, rank)
            red.begin()
            ((model(x[sl]) - y[sl]) ** 2).mean().backward()
            launched_in_backward = sum(red.launched)
            n = red.finish(average=True)
            ref = torch.nn.Sequential(torch.nn.Linear(7, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
            ref.load_state_dict(model.state_dict())
            ((ref(x) - y) ** 2).mean().backward()
            errs.append(max(float((a.grad - b.grad).abs().max()) for a, b in zip(model.parameters(), ref.parameters())))
            assert launched_in_backward >= 2 and n == sum(p.numel() for p in params)
            assert float(unused.grad.abs().max()) == 0.0
        red.close()
        q.put((rank, max(errs)))
    finally:
        dist.destroy_process_group()


def test_overlapped_reducer_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlap, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(err <= 1e-6 for _, err in out)


def test_allreduce_is_a_noop_without_a_process_group():
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert allreduce_gradients([p]) == 0 and float(p.grad[0]) == 2.0


def test_discriminators_keep_the_reference_state_dict_layout():
    """models/discriminator.py:6-80: parameter names and shapes (checkpoints carry `disc_state_dict`)."""
    from spoofsv_b200.train import guided_attention_mat, linDisc, melDisc
    m, l = melDisc(80, 128), linDisc(513, 128)
    keys = list(m.state_dict().keys())
    assert keys[:4] == ["conv1.weight", "conv1.bias", "ln1.weight", "ln1.bias"]
    assert keys[4:10] == ["hc.conv.weight", "hc.conv.bias", "hc.ln1.weight", "hc.ln1.bias", "hc.ln2.weight", "hc.ln2.bias"]
    assert list(l.state_dict().keys()) == keys
    assert tuple(m.state_dict()["hc.conv.weight"].shape) == (256, 128, 3) and tuple(m.state_dict()["conv4.weight"].shape) == (4, 16, 1)
    assert tuple(l.state_dict()["conv4.weight"].shape) == (8, 16, 1) and tuple(l.state_dict()["conv1.weight"].shape) == (128, 513, 1)
    assert sum(p.numel() for p in m.parameters()) == 119233          # elements all-reduced on a D step
    x = torch.rand(2, 80, 40)
    assert tuple(m.eval()(x).shape) == (2, 1, 1)
    w = guided_attention_mat(186, 325)
    assert tuple(w.shape) == (186, 325) and float(w[0, 0]) == 0.0 and 0.99 < float(w[0, 324]) <= 1.0
