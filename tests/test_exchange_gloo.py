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


# ---------------------------------------------------------------------------------------------------------------
# row-sharded database search (SURVEY.md 8e alternative): the exchange logic on gloo, shard scan / merge replaced by
# the CPU oracle (the CUDA steps themselves are checked in tests/test_gpu_match.py)
# ---------------------------------------------------------------------------------------------------------------
class _HostShardedIndex(swarm.RowShardedIndex):
    def _open(self, shard):
        from oracle import frontend_ref as fr
        self._index = fr.IndexFlatIP(self.dim)
        self._index.add(shard)

    def _local_search(self, queries, k):
        D, I = self._index.search(queries.numpy(), k)
        return torch.from_numpy(D), torch.from_numpy(I)

    def _merge(self, cand_scores, cand_ids, k):
        s = cand_scores.numpy().copy().reshape(-1)
        i = cand_ids.numpy().copy().reshape(cand_scores.shape[0], -1)
        i = np.where(i >= 0, i + np.asarray(self.starts, np.int64)[:, None], -1).reshape(-1)
        keep = i >= 0
        s, i = s[keep], i[keep]
        order = np.lexsort((i, -s.astype(np.float64)))[:k]           # score descending, ties by ascending global id
        out_s = np.full(k, -np.inf, np.float32); out_i = np.full(k, -1, np.int64)
        out_s[:len(order)] = s[order]; out_i[:len(order)] = i[order]
        return torch.from_numpy(out_s), torch.from_numpy(out_i)

    def close(self):
        self._index = None


def _sharded_worker(rank, world, port, q, n_rows, dim, k):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    db = rng.standard_normal((n_rows, dim)).astype(np.float32)
    db[7] = db[3]                                                   # a tie across (possibly) different shards
    a, b = swarm.shard_rows(n_rows, rank, world)
    idx = _HostShardedIndex(db[a:b], n_rows, dim)
    query = db[3 + rank * 4].copy()                                 # rank 0 queries row 3 (tied with 7), rank 1 row 7
    s, i = idx.search(torch.from_numpy(query), k)
    q.put((rank, s.numpy(), i.numpy(), (a, b)))
    dist.destroy_process_group()


def test_row_sharded_search_matches_unsharded_oracle():
    from oracle import frontend_ref as fr
    n_rows, dim, k = 37, 64, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q, n_rows, dim, k)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    rng = np.random.default_rng(5)
    db = rng.standard_normal((n_rows, dim)).astype(np.float32)
    db[7] = db[3]
    full = fr.IndexFlatIP(dim); full.add(db)
    assert [r[3] for r in res] == [(0, 19), (19, 37)]
    for rank, s, i, _ in res:
        D, I = full.search(db[3 + rank * 4][None], k)
        assert np.array_equal(i, I[0]), (rank, i, I[0])
        assert np.array_equal(s, D[0])
        assert list(i[:2]) == [3, 7]                                # the tie resolves by ascending GLOBAL row id


def test_shard_rows_cover_everything():
    for n in (0, 1, 7, 8, 50000):
        for w in (1, 2, 3, 8):
            spans = [swarm.shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
