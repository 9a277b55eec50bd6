"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests).  The path shards with NO data-path collective:

* MH / exploration: chains (or, in the reference's single-chain mode, proposals) are partitioned
  contiguously over the ranks, per-rank generators are seeded `seed + rank`, weights are replicated.
  The only collectives are an all-gather of per-chain acceptance statistics per reporting interval and,
  in single-chain mode, a MIN all-reduce of the first accepted proposal index (4 bytes per iteration,
  utils/evaluation_utils.py:675-689 picks the FIRST accepted proposal).
* NLL training: batch-sharded data parallel.  The reference drives this with DeepSpeed ZeRO-1
  (train_deepspeed.py:99-120): gradients are averaged over the data-parallel group, the loss value is
  divided by the world size and all-reduced for logging (train_deepspeed.py:186-188).  Here: ONE bucketed
  all-reduce(sum) over the flattened gradients followed by a 1/world scale, then the local optimizer step
  on replicated parameters (36 M parameters = 144 MB fp32: ZeRO partitioning buys nothing on 180 GB GPUs).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def init_from_env(backend: Optional[str] = None, device: Optional[torch.device] = None) -> Tuple[int, int]:
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun's
    environment).  Returns (rank, world).  A single process (WORLD_SIZE unset or 1) needs no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if (device is not None and device.type == "cuda") else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def world_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of n units: ranks < n % world get one extra unit."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"invalid rank {rank} / world {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[Tensor], rank: int, world: int, dim: int = 0) -> List[Tensor]:
    """This rank's contiguous slice of every tensor along `dim` (same partition as shard_range)."""
    out = []
    for t in tensors:
        lo, hi = shard_range(t.shape[dim], rank, world)
        out.append(t.narrow(dim, lo, hi - lo))
    return out


def rank_seed(seed: int, rank: int) -> int:
    return int(seed) + int(rank)


# ------------------------------------------------------------------------------------------ sampling
def allgather_chain_stats(stat: Tensor, counts: Optional[Sequence[int]] = None) -> Tensor:
    """Concatenate a per-chain statistic (acceptance indicator / rate) over the ranks in rank order.
    `counts[r]` = chains on rank r when the shards are ragged (shard_range); equal shards otherwise."""
    rank, world = world_info()
    if world == 1:
        return stat
    stat = stat.contiguous()
    if counts is None:
        out = [torch.empty_like(stat) for _ in range(world)]
        dist.all_gather(out, stat)
        return torch.cat(out)
    width = max(counts)
    padded = stat.new_zeros((width,) + tuple(stat.shape[1:]))
    padded[: stat.shape[0]] = stat
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.cat([o[:c] for o, c in zip(out, counts)])


def first_accepted_global(first_local: Tensor, offset: int, none_value: int) -> Tensor:
    """Single-chain mode with the S proposals sharded over ranks: `first_local` is the index of the first
    accepted proposal in this rank's shard (or < 0 if none).  Returns the global index of the first accepted
    proposal over all ranks, `none_value` if no rank accepted (one MIN all-reduce of a single integer)."""
    g = torch.where(first_local >= 0, first_local + offset, torch.full_like(first_local, none_value))
    rank, world = world_info()
    if world > 1:
        dist.all_reduce(g, op=dist.ReduceOp.MIN)
    return g


# ------------------------------------------------------------------------------------------ training
def broadcast_parameters(params: Iterable[Tensor], src: int = 0) -> None:
    """Replicate rank `src`'s parameters (call once after construction / checkpoint load)."""
    rank, world = world_info()
    if world == 1:
        return
    for p in params:
        dist.broadcast(p.data, src)


class GradientBuckets:
    """Flat fp32 buckets over the gradients of `params` (fixed order = parameter order).  all_reduce()
    copies the gradients in, runs ONE all-reduce(sum) per bucket, scales by 1/world and copies back.
    With bucket_bytes >= the model size (default 256 MB > 144 MB) that is a single collective per step."""

    def __init__(self, params: Iterable[Tensor], bucket_bytes: int = 256 << 20):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        cap = max(int(bucket_bytes) // 4, 1)
        self.plan: List[List[Tuple[Tensor, int, int]]] = [[]]
        fill = 0
        for p in self.params:
            n = p.numel()
            if fill > 0 and fill + n > cap:
                self.plan.append([])
                fill = 0
            self.plan[-1].append((p, fill, n))
            fill += n
        self.buffers = [torch.zeros(sum(n for _, _, n in b), dtype=torch.float32, device=dev) for b in self.plan]

    @property
    def num_buckets(self) -> int:
        return len(self.buffers)

    @torch.no_grad()
    def all_reduce(self, average: bool = True) -> None:
        rank, world = world_info()
        handles = []
        for buf, bucket in zip(self.buffers, self.plan):
            for p, off, n in bucket:
                if p.grad is None:
                    buf[off:off + n].zero_()
                else:
                    buf[off:off + n].copy_(p.grad.reshape(-1))
            if world > 1:
                handles.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True))
        for h in handles:
            h.wait()
        for buf, bucket in zip(self.buffers, self.plan):
            if average and world > 1:
                buf.mul_(1.0 / world)
            for p, off, n in bucket:
                if p.grad is None:
                    p.grad = buf[off:off + n].reshape(p.shape).clone()
                else:
                    p.grad.copy_(buf[off:off + n].reshape(p.shape))


def all_reduce_loss(loss_value: Tensor) -> Tensor:
    """train_deepspeed.py:186-188: divide by the data-parallel world size, then all-reduce (sum)."""
    rank, world = world_info()
    out = loss_value.detach().clone()
    if world > 1:
        out.div_(world)
        dist.all_reduce(out)
    return out


def clip_grad_norm(params: Iterable[Tensor], max_norm: float) -> Tensor:
    """Global L2 clipping on the (already averaged) gradients -- DeepSpeed's `gradient_clipping`
    (train_deepspeed.py:116-117; 0.0 / None disables it)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:  # nothing to clip (e.g. a step before any backward)
        return torch.zeros(())
    total = torch.sqrt(sum((g.detach().float() ** 2).sum() for g in grads))
    if max_norm and max_norm > 0:
        scale = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        for g in grads:
            g.mul_(scale)
    return total


class DataParallelTrainer:
    """One NLL training step of the batch-sharded data-parallel job: local forward (loss of
    density_model_base.py:14-47 on this rank's shard) + hand-written backward, one bucketed gradient
    all-reduce, optional clipping, local optimizer step.  `loss_fn(model, batch) -> scalar` defaults to
    calling the model with the batch's keyword tensors (losses.py:346-356)."""

    def __init__(self, model, optimizer, clip_grad_norm_value: Optional[float] = None, bucket_bytes: int = 256 << 20, loss_fn=None):
        self.model, self.optimizer = model, optimizer
        self.clip = clip_grad_norm_value
        self.buckets = GradientBuckets(model.parameters(), bucket_bytes)
        self.loss_fn = loss_fn or (lambda m, batch: m(**batch))
        # optional device timing of the gradient collective (bench.py): CUDA event pairs recorded on the current stream around
        # the all-reduce (the NCCL stream is joined to it on both sides), one pair per step; see `collective_ms`
        self.time_collective = False
        self._collective_events: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []
        self.collective_bytes = 0

    def collective_ms(self, reset: bool = True) -> float:
        """Mean device time of the gradient all-reduce (+ the 1/world scale) over the steps since the last reset."""
        if not self._collective_events:
            return 0.0
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._collective_events) / len(self._collective_events)
        if reset:
            self._collective_events = []
        return ms

    @torch.no_grad()
    def _all_reduce_flat(self) -> bool:
        """The CUDA model's backward hands out every gradient as a view of ONE flat buffer: reduce that buffer in place
        (a single collective, no per-parameter copies).  False when the gradients do not live in such a buffer."""
        flat = getattr(self.model, "_last_flat_grad", None)
        if flat is None:
            return False
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        seen = False
        for p in self.model.parameters():
            if not p.requires_grad or p.grad is None:  # (unused parameters, e.g. the log_lengthscales a pass never reads,
                continue                               #  have no gradient on ANY rank: nothing to reduce)
            a = p.grad.data_ptr()  # (address range instead of storage identity: no Python storage object per parameter and step)
            if a < lo or a + p.grad.numel() * p.grad.element_size() > hi:
                return False
            seen = True
        if not seen:
            return False
        rank, world = world_info()
        self.collective_bytes = flat.numel() * flat.element_size()
        if world > 1:
            timed = self.time_collective and flat.is_cuda
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / world)
            if timed:
                e1.record()
                self._collective_events.append((e0, e1))
        return True

    def step(self, local_batch: dict) -> Tensor:
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model, local_batch)
        loss.backward()
        if not self._all_reduce_flat():
            self.buckets.all_reduce(average=True)
        if self.clip:
            clip_grad_norm(self.model.parameters(), self.clip)
        self.optimizer.step()
        return all_reduce_loss(loss)
