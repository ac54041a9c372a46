"""world_size-2 gloo test of the multi-GPU host logic on CPU: the setup
broadcast of the packets, stream->rank assignment, and the reductions bench.py
reports with.  Each rank parses its shard with the reference host code +
recorder (record mode, no device) to show ranks work independently."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import support as S
from theora_b200 import sharding
import th_streams as streams

G = np.load(os.path.join(S.GOLDEN_DIR, "streams.npz"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, blob, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = sharding.broadcast_bytes(blob if rank == 0 else b"", 0)
        mine = sharding.stream_assignment(5, world)[rank]
        frames = 0
        if streams.available():
            for _ in mine:
                _, works, _ = streams.capture_stream_work(got, streams.BACKEND_RECORD)
                frames += len(works)
        else:
            frames = 6 * len(mine)
        sharding.quiet_barrier("after_parse")  # store-based rendezvous (no spinning), as bench.py uses around e2e
        total = sharding.sum_over_ranks(frames)
        worst = sharding.max_over_ranks(1.0 + rank)
        q.put((rank, got == blob, mine, frames, total, worst))
    finally:
        dist.destroy_process_group()


def test_stream_assignment_is_a_partition():
    for n in (1, 5, 8, 13):
        for w in (1, 2, 4, 8):
            parts = sharding.stream_assignment(n, w)
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_ranks_share_only_the_setup_broadcast():
    blob = G["s64_q32_kf4_blob"].tobytes()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, blob, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "broadcast corrupted the packets"
    assert res[0][2] == [0, 2, 4] and res[1][2] == [1, 3]
    assert res[0][3] == 18 and res[1][3] == 12  # 6 frames per stream
    assert res[0][4] == res[1][4] == 30.0
    assert res[0][5] == res[1][5] == 2.0
