"""Host-side multi-GPU logic on CPU: world_size-2 `gloo` process groups (the N > 1 path of SURVEY.md section 8e).
The CUDA model cannot run here (no CPU fallback), so the data-parallel trainer is exercised with a small
stand-in module: what is tested is the sharding, the bucketed gradient all-reduce, the loss reduction and the
acceptance-statistics collectives."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from timewarp_b200 import distributed as twd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    r, w = twd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return [ret[r] for r in range(world)]


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            spans = [twd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        twd.shard_range(4, 2, 2)


def test_bucket_plan_splits_on_capacity():
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (10, 20, 30, 5)]
    assert twd.GradientBuckets(ps, bucket_bytes=1 << 20).num_buckets == 1
    b = twd.GradientBuckets(ps, bucket_bytes=30 * 4)
    assert [sum(n for _, _, n in bk) for bk in b.plan] == [30, 30, 5]


# ---------------------------------------------------------------- world_size 2
def _grad_allreduce(rank, world):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    twd.broadcast_parameters(model.parameters())
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
    xs, ys = twd.shard_batch([X, Y], rank, world)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    trainer = twd.DataParallelTrainer(model, opt, bucket_bytes=64, loss_fn=lambda m, b: ((m(b["x"]) - b["y"]) ** 2).mean())
    assert trainer.buckets.num_buckets > 1
    loss = trainer.step({"x": xs, "y": ys})
    return float(loss), [p.detach().clone() for p in model.parameters()], [p.grad.clone() for p in model.parameters()]


def test_data_parallel_step_equals_full_batch_step():
    out = _run(_grad_allreduce)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
    loss = ((model(X) - Y) ** 2).mean()
    loss.backward()
    grads = [p.grad.clone() for p in model.parameters()]
    torch.optim.SGD(model.parameters(), lr=0.1).step()
    for rank_out in out:
        assert abs(rank_out[0] - float(loss)) < 1e-6
        for a, b in zip(rank_out[2], grads):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)
        for a, b in zip(rank_out[1], model.parameters()):
            torch.testing.assert_close(a, b.detach(), rtol=1e-5, atol=1e-7)
    for a, b in zip(out[0][1], out[1][1]):  # replicas stay bit-identical
        assert torch.equal(a, b)


def _stats(rank, world):
    n = 7
    lo, hi = twd.shard_range(n, rank, world)
    acc = (torch.arange(lo, hi) % 3 == 0).float()
    counts = [twd.shard_range(n, r, world)[1] - twd.shard_range(n, r, world)[0] for r in range(world)]
    ragged = twd.allgather_chain_stats(acc, counts)
    even = twd.allgather_chain_stats(torch.full((2,), float(rank)))
    # single-chain mode: 10 proposals sharded 5 + 5; rank 0 accepts none, rank 1 accepts local index 2 -> global 7
    first = twd.first_accepted_global(torch.tensor([-1 if rank == 0 else 2]), offset=5 * rank, none_value=10)
    none = twd.first_accepted_global(torch.tensor([-1]), offset=5 * rank, none_value=10)
    return ragged, even, int(first), int(none), twd.rank_seed(100, rank)


def test_acceptance_collectives():
    out = _run(_stats)
    want = (torch.arange(7) % 3 == 0).float()
    for r, (ragged, even, first, none, seed) in enumerate(out):
        assert torch.equal(ragged, want)
        assert torch.equal(even, torch.tensor([0.0, 0.0, 1.0, 1.0]))
        assert first == 7 and none == 10 and seed == 100 + r


def test_single_process_is_a_no_op():
    assert twd.world_info() == (0, 1)
    t = torch.tensor([1.0, 0.0])
    assert twd.allgather_chain_stats(t) is t
    assert float(twd.all_reduce_loss(torch.tensor(3.0))) == 3.0
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    twd.GradientBuckets([p]).all_reduce()
    assert torch.equal(p.grad, torch.full((3,), 2.0))
    assert abs(float(twd.clip_grad_norm([p], 1.0)) - 12 ** 0.5) < 1e-6 and abs(float(p.grad.norm()) - 1.0) < 1e-4


class _FlatGradModel(torch.nn.Module):
    """Stand-in for the CUDA model's contract: after backward every gradient is a view of ONE flat buffer published as
    `_last_flat_grad`; `unused` never receives a gradient (like the log_lengthscales a pass does not read)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Parameter(torch.randn(3, 2))
        self.b = torch.nn.Parameter(torch.randn(2))
        self.unused = torch.nn.Parameter(torch.zeros(4))
        self._last_flat_grad = None

    def forward(self, x, y):
        return (((x @ self.a + self.b) - y) ** 2).mean()


def _flat_allreduce(rank, world):
    model = _FlatGradModel()
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(8, 3, generator=g), torch.randn(8, 2, generator=g)
    lo, hi = twd.shard_range(8, rank, world)

    def loss_fn(m, batch):  # autograd computes the gradients, then they are re-homed into one flat buffer
        loss = m(batch["x"], batch["y"])
        ga, gb = torch.autograd.grad(loss, [m.a, m.b])
        flat = torch.cat([ga.reshape(-1), gb.reshape(-1)])
        m.a.grad, m.b.grad = flat[:6].view(3, 2), flat[6:].view(2)
        m._last_flat_grad = flat
        return loss.detach().requires_grad_(True) * 1.0  # (its own backward is a no-op for the parameters)

    trainer = twd.DataParallelTrainer(model, torch.optim.SGD([model.a, model.b], lr=0.1), loss_fn=loss_fn)
    calls = []
    orig = trainer.buckets.all_reduce
    trainer.buckets.all_reduce = lambda *a, **k: (calls.append(1), orig(*a, **k))
    trainer.step({"x": X[lo:hi], "y": Y[lo:hi]})
    return [model.a.grad.clone(), model.b.grad.clone()], len(calls), model.unused.grad is None


def test_flat_gradient_buffer_is_reduced_in_place():
    """One collective over the flat buffer (no bucket copies), also when some parameters have no gradient on any rank."""
    out = _run(_flat_allreduce)
    model = _FlatGradModel()
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(8, 3, generator=g), torch.randn(8, 2, generator=g)
    model(X, Y).backward()
    for grads, bucket_calls, unused_none in out:
        assert bucket_calls == 0 and unused_none
        torch.testing.assert_close(grads[0], model.a.grad, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(grads[1], model.b.grad, rtol=1e-5, atol=1e-7)
