"""CPU: synthetic generators and host-side plumbing."""
import numpy as np

from omniswarm_b200 import synth, lib


def test_generators_are_deterministic():
    a, b = synth.superpoint_weights(0), synth.superpoint_weights(0)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert synth.flatten_sp_weights(a).size == 1300865 == synth.sp_num_weights()
    assert synth.flatten_nv_weights(synth.netvlad_weights(0)).size == synth.nv_num_weights()
    assert np.array_equal(synth.image(3), synth.image(3))
    img = synth.image(3, zero_bottom_quarter=True)
    assert img[360:].max() == 0 and img[:360].max() > 0


def test_c5_graph_shape():
    g = synth.pose_graph_c5(0)
    assert g["n_nodes"] == 2000 and len(g["ftype"]) == 12000
    assert (g["ftype"] == synth.FACTOR_DISTANCE).sum() == 4000
    assert (g["ia"] != g["ib"]).all() and g["fixed"].sum() == 1
    assert g["payload"].shape == (12000, lib.PAYLOAD_LEN)


def test_descriptor_db_unit_norm():
    db = synth.descriptor_db(100)
    assert np.allclose(np.linalg.norm(db, axis=1), 1.0, atol=1e-5)
    q = synth.noisy_queries(db, np.array([5, 6]))
    assert ((db @ q.T).argmax(0) == [5, 6]).all() and (db[[5, 6]] * q).sum(1).min() > 0.8


def test_solver_chain_plan_recovers_odometry_chains():
    """Host logic of the chain preconditioner (no GPU): the greedy path cover must be a permutation, every link must be
    backed by a factor between the two nodes, links never cross a 16-node boundary, fixed nodes come last, and on the
    swarm graph the odometry chains are recovered (almost every free node is linked to its predecessor)."""
    from omniswarm_b200 import host, synth
    g = synth.pose_graph(5, 40, seed=3)
    rng = np.random.default_rng(0)
    perm = rng.permutation(g["n_nodes"])                       # caller's node ids in arbitrary order
    g2 = dict(g)
    g2["ia"] = perm[g["ia"]].astype(np.int32); g2["ib"] = perm[g["ib"]].astype(np.int32)
    fixed = np.zeros_like(g["fixed"]); fixed[perm[np.nonzero(g["fixed"])[0]]] = 1
    g2["fixed"] = fixed
    order, link = host.PoseGraphSolver.chain_plan(g2)
    n = g["n_nodes"]
    assert sorted(order.tolist()) == list(range(n))
    n_fixed = int(fixed.sum())
    assert fixed[order[n - n_fixed:]].all() and not fixed[order[:n - n_fixed]].any()
    pairs = {(int(a), int(b)) for a, b in zip(g2["ia"], g2["ib"])} | {(int(b), int(a)) for a, b in zip(g2["ia"], g2["ib"])}
    for i in np.nonzero(link)[0]:
        assert i % 16 != 0
        assert (int(order[i - 1]), int(order[i])) in pairs
    free = n - n_fixed
    # 5 chains of 40 (one starts at the fixed node): all but the path starts and the 16-boundaries are linked
    assert link.sum() >= free - 5 - free // 16 - 1
    # the links follow ego-motion edges: consecutive frames of the same drone in the ORIGINAL numbering
    inv = np.argsort(perm)
    orig = inv[order]
    lk = np.nonzero(link)[0]
    same_drone = (orig[lk] % 5) == (orig[lk - 1] % 5)
    assert same_drone.mean() > 0.95 and (np.abs(orig[lk] // 5 - orig[lk - 1] // 5)[same_drone] == 1).all()


def test_bench_database_capacity_covers_every_round():
    """bench.py sizes the keyframe stores from the rounds of the run: per round a drone adds 4 own rows to its local store and
    4 rows per other drone to its remote store, and the host's conservative bound charges every gathered record to both
    stores.  Neither the true counts nor the bound may reach the capacity, for any world size and step count the driver
    may pass (an undersized remote store once failed an 8-rank run in its replay leg)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for world in (1, 2, 4, 8):
        for steps, warmup in ((20, 3), (1, 0), (100, 10), (2, 1)):
            rounds = bench.keyframe_rounds(steps, warmup)
            assert rounds == 2 * (steps + warmup) * bench.KF_PER_STEP + 8 + 5 + 120
            cap = bench.db_capacity(10_000, world, steps, warmup)
            local_rows = 10_000 + 4 * rounds
            remote_rows = 4 * (world - 1) * rounds
            bound = 10_000 + 4 * rounds + 4 * world * rounds          # own ingest + every gathered record
            assert cap > local_rows and cap > remote_rows and cap > bound, (world, steps, cap, bound)
