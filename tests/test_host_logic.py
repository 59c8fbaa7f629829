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
