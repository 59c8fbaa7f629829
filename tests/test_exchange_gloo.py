"""CPU, world_size 2 (gloo): the keyframe-record all-gather and its routing -- the N>1 path of bench.py without GPUs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from omniswarm_b200 import lib, swarm


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec = lib.KeyframeRecord()
    rec.drone_id, rec.msg_id, rec.n_dirs = rank + 1, 100 + rank, 4
    for d in range(4):
        rec.n_kpts[d] = (d + rank) % 3            # some directions empty
    np.ctypeslib.as_array(rec.global_desc[1])[:] = float(rank + 1)
    mine = torch.frombuffer(bytearray(bytes(rec)), dtype=torch.uint8)
    gathered = swarm.exchange_records(mine)
    routes = swarm.routing(gathered, self_id=rank + 1)
    other = lib.KeyframeRecord.from_buffer_copy(swarm.record_view(gathered, 1 - rank).numpy().tobytes())
    q.put((rank, routes, float(np.ctypeslib.as_array(other.global_desc[1])[7]), gathered.numel()))
    dist.destroy_process_group()


def test_record_all_gather_and_routing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in range(2)])
    [p.join(timeout=60) for p in ps]
    for rank, routes, other_val, n in res:
        assert n == 2 * lib.RECORD_BYTES
        assert other_val == float(2 - rank)                       # the peer's payload arrived intact
        assert [r[3] for r in routes] == (["local", "remote"] if rank == 0 else ["remote", "local"])
        assert [r[1] for r in routes] == [1, 2] and [r[2] for r in routes] == [100, 101]
        assert routes[0][4] == [1, 2] and routes[1][4] == [0, 1, 3]   # only directions with landmark_num > 0 are added


def test_single_rank_exchange_is_identity():
    rec = torch.arange(lib.RECORD_BYTES, dtype=torch.int64).to(torch.uint8)
    assert torch.equal(swarm.exchange_records(rec), rec)
