"""CPU, world_size 2 (gloo): the data-parallel host logic -- contiguous image shards, gather of
per-image maps and scores to the metric sink in global order, ragged last shard -- without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_ranges_cover_everything_once():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import dp
    for n in (0, 1, 5, 32, 1024, 1027):
        for world in (1, 2, 4, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = dp.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, S, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import aaclip_b200  # noqa: F401
        from aaclip_b200 import dp
        lo, hi = dp.shard_range(n_total, rank, world)
        # image i's "map" and "score" are functions of the GLOBAL index only: any sharding must reproduce them
        idx = torch.arange(lo, hi, dtype=torch.float32)
        maps = idx[:, None, None] + torch.arange(S * S, dtype=torch.float32).reshape(S, S) / (S * S)
        scores = idx * 0.5
        gm, gs = dp.gather_predictions(maps, scores, n_total=n_total, dst=0)
        gm2, gs2 = dp.gather_predictions(maps, scores, n_total=None, dst=0)      # count discovered by all-reduce
        if rank == 0:
            all_idx = torch.arange(n_total, dtype=torch.float32)
            want_m = all_idx[:, None, None] + torch.arange(S * S, dtype=torch.float32).reshape(S, S) / (S * S)
            ok = (torch.equal(gm, want_m) and torch.equal(gs, all_idx * 0.5) and torch.equal(gm2, want_m)
                  and torch.equal(gs2, all_idx * 0.5))
            q.put(("ok" if ok else "mismatch", tuple(gm.shape)))
        else:
            assert gm is None and gs is None
        # a rank holding the wrong number of images is an error on that rank, not a silent mis-ordering
        if rank == 1 and hi - lo > 0:
            try:
                dp.gather_predictions(maps[:-1], scores[:-1], n_total=n_total, dst=0)
                q.put(("no-error", ()))
            except ValueError:
                pass
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 5])      # equal shards, ragged last shard
def test_gather_predictions_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    status, shape = q.get(timeout=5)
    assert status == "ok" and shape == (n_total, 6, 6)
    assert q.empty()


def test_single_process_gather_is_identity():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import dp
    m, s = torch.zeros(3, 4, 4), torch.zeros(3)
    gm, gs = dp.gather_predictions(m, s)
    assert gm is m and gs is s


def _metrics_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import aaclip_b200  # noqa: F401
        from aaclip_b200 import metrics
        # the statistics the device metrics reduce across ranks: (min, max) and the two count histograms
        mm = metrics.reduce_minmax(torch.tensor([-1.0 - rank, 2.0 + 3 * rank]))
        hist = torch.zeros(2, 16, dtype=torch.int32)
        hist[rank, rank + 3] = 5 + rank
        hist[0, 0] = 1
        total = metrics.reduce_histogram(hist)
        # image-level statistics: ragged per-rank vectors, DEFAULT process group with group=None (the case that used
        # to stay shard-local), concatenated in rank order on every rank
        mine = torch.arange(3 + 2 * rank, dtype=torch.float32) + 10 * rank
        everyone = metrics.gather_concat(mine)
        want_all = torch.cat([torch.arange(3, dtype=torch.float32), torch.arange(5, dtype=torch.float32) + 10])
        ok_gather = torch.equal(everyone, want_all)
        flags = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(flags, torch.tensor([float(ok_gather)]))
        if rank == 0:
            want = torch.zeros(2, 16, dtype=torch.int64)
            want[0, 3], want[1, 4], want[0, 0] = 5, 6, 2
            q.put((mm.tolist() == [-2.0, 5.0], torch.equal(total, want) and all(f.item() == 1.0 for f in flags)))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_metric_statistics_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_metrics_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == (True, True)


def _async_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import aaclip_b200  # noqa: F401
        from aaclip_b200 import dp
        per_rank, micro, S = 6, 2, 3                      # 3 micro-batches of 2 images per rank (config-5 shape, tiny)
        g = dp.AsyncGather(per_rank, micro, (S, S), "cpu", dst=0)
        for m in range(per_rank // micro):
            idx = torch.arange(rank * per_rank + m * micro, rank * per_rank + (m + 1) * micro, dtype=torch.float32)
            g.submit(m, idx[:, None, None].expand(micro, S, S).contiguous(), idx * 2)
        g.wait()
        if rank == 0:
            want = torch.arange(world * per_rank, dtype=torch.float32)
            q.put(bool(torch.equal(g.out_scores, want * 2) and torch.equal(g.out_maps[:, 0, 0], want)
                       and torch.equal(g.out_maps[:, 2, 1], want)))
        else:
            assert g.out_maps is None
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_async_microbatch_gather_world2():
    """bench.py's N > 1 path (BASELINE config 5): micro-batch results gathered to the sink in global image order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_async_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
    with pytest.raises(ValueError):
        import aaclip_b200  # noqa: F401
        from aaclip_b200 import dp
        dp.AsyncGather(7, 2, (3, 3), "cpu")
