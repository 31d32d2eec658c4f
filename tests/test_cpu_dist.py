"""world_size-2 gloo test of the multi-GPU host logic (clip sharding, output gather, metric reduce) on CPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from sparsernns_b200.dist import gather_outputs, reduce_metrics, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_clips * 4 * 3, dtype=torch.float32).reshape(n_clips, 4, 3)
    a, b = shard_range(n_clips, rank, world)
    local = full[a:b] * 2.0                       # stands in for the per-rank forward of independent clips
    out = gather_outputs(local, n_clips)
    met = reduce_metrics(torch.tensor([float(local.sum()), float(b - a)]), "sum")
    mx = reduce_metrics(torch.tensor([float(rank)]), "max")
    ok = torch.equal(out, full * 2.0) and float(met[0]) == float((full * 2).sum()) and int(met[1]) == n_clips and int(mx[0]) == world - 1
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 5])
def test_gloo_world_size_2(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_parse_cpulist_and_numa_binding_is_best_effort():
    from sparsernns_b200.dist import bind_to_gpu_numa_node, parse_cpulist

    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == []
    info = bind_to_gpu_numa_node(0)          # no GPU here: reports the reason, never raises
    assert info["bound"] is False and "error" in info
