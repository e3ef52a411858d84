"""Data-parallel sharding of the image batch: one process per GPU, weights replicated, images are
independent units, and the only collective is the gather of per-image scores and maps to the metric sink
(rank 0), where the reference's metrics_eval (forward_utils.py:233-308) runs its global min-max on all images.

The reference has no distributed code at all (device hard-wired to cuda:0, test_last.py:206); this module is the
B200-native addition. It is backend-agnostic host logic: NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_items over `world` ranks; the first (n_items % world) ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad shard request n={n_items} rank={rank} world={world}")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_predictions(maps: torch.Tensor, scores: torch.Tensor, n_total: Optional[int] = None, dst: int = 0,
                       group=None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Gather per-rank maps [b_r,S,S] and scores [b_r] to `dst` in global image order.

    Ranks may hold different b_r (ragged last shard). Returns (maps [n,S,S], scores [n]) on dst, (None, None)
    elsewhere. Without an initialised process group (single process) it is the identity.
    """
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return maps, scores
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if n_total is None:
        cnt = torch.tensor([maps.shape[0]], device=maps.device, dtype=torch.int64)
        dist.all_reduce(cnt, group=group)
        n_total = int(cnt.item())
    sizes = shard_sizes(n_total, world)
    if maps.shape[0] != sizes[rank] or scores.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {maps.shape[0]} images, expected {sizes[rank]} of {n_total}")
    S = maps.shape[1:]
    if len(set(sizes)) == 1:
        # equal shards: one gather each for maps and scores
        if rank == dst:
            out_m = torch.empty((n_total, *S), dtype=maps.dtype, device=maps.device)
            out_s = torch.empty((n_total,), dtype=scores.dtype, device=scores.device)
            dist.gather(maps.contiguous(), list(out_m.chunk(world)), dst=dst, group=group)
            dist.gather(scores.contiguous(), list(out_s.chunk(world)), dst=dst, group=group)
            return out_m, out_s
        dist.gather(maps.contiguous(), None, dst=dst, group=group)
        dist.gather(scores.contiguous(), None, dst=dst, group=group)
        return None, None
    # ragged shards: pad to the largest, gather, trim on the sink
    mx = max(sizes)
    pm = torch.zeros((mx, *S), dtype=maps.dtype, device=maps.device)
    ps = torch.zeros((mx,), dtype=scores.dtype, device=scores.device)
    pm[: maps.shape[0]] = maps
    ps[: scores.shape[0]] = scores
    if rank == dst:
        gm = [torch.empty_like(pm) for _ in range(world)]
        gs = [torch.empty_like(ps) for _ in range(world)]
        dist.gather(pm, gm, dst=dst, group=group)
        dist.gather(ps, gs, dst=dst, group=group)
        return (torch.cat([g[:n] for g, n in zip(gm, sizes)]), torch.cat([g[:n] for g, n in zip(gs, sizes)]))
    dist.gather(pm, None, dst=dst, group=group)
    dist.gather(ps, None, dst=dst, group=group)
    return None, None


class AsyncGather:
    """Gather of equal-sized micro-batch results to the sink, overlapped with the next micro-batch's kernels
    (BASELINE config 5): ``submit`` enqueues NCCL ``gather``s that wait for the producing stream and then run on the
    communicator's own stream; ``wait`` makes the current stream wait for all of them. On the sink the results land
    directly in their global positions of ``out_maps`` / ``out_scores`` (rank-major, then micro-batch, then image:
    the contiguous split of shard_range), so no re-ordering pass is needed.
    """

    def __init__(self, per_rank: int, micro: int, map_shape, device, dst: int = 0, group=None):
        if per_rank % micro:
            raise ValueError(f"per-rank share {per_rank} is not a multiple of the micro-batch {micro}")
        self.group, self.dst = group, dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.per_rank, self.micro = per_rank, micro
        self.out_maps = self.out_scores = None
        if self.rank == dst:
            self.out_maps = torch.empty((self.world * per_rank, *map_shape), dtype=torch.float32, device=device)
            self.out_scores = torch.empty((self.world * per_rank,), dtype=torch.float32, device=device)
        self._pending = []

    def submit(self, m: int, maps: torch.Tensor, scores: torch.Tensor) -> None:
        """Results of this rank's micro-batch m (images [m*micro, (m+1)*micro) of its shard)."""
        if self.world == 1:
            self.out_maps[m * self.micro:(m + 1) * self.micro].copy_(maps)
            self.out_scores[m * self.micro:(m + 1) * self.micro].copy_(scores)
            return
        lists = (None, None)
        if self.rank == self.dst:
            sl = [slice(r * self.per_rank + m * self.micro, r * self.per_rank + (m + 1) * self.micro)
                  for r in range(self.world)]
            lists = ([self.out_maps[s] for s in sl], [self.out_scores[s] for s in sl])
        maps, scores = maps.contiguous(), scores.contiguous()
        w1 = dist.gather(maps, lists[0], dst=self.dst, group=self.group, async_op=True)
        w2 = dist.gather(scores, lists[1], dst=self.dst, group=self.group, async_op=True)
        self._pending.append((w1, w2, maps, scores))     # the sources stay alive until the collective has run

    def wait(self) -> None:
        for w1, w2, _, _ in self._pending:
            w1.wait()
            w2.wait()
        self._pending.clear()
