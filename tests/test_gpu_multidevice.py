"""Several GPUs behind ONE engine handle (kb_finalize_multi, SURVEY 8b "kb_finalize(device_list)"): host-buffer batches are sharded
over the devices inside the C ABI and must give exactly the answers of one device.  Needs two GPUs (gpurun --gpus 2)."""
import numpy as np
import pytest

from klampt_b200 import synth

pytestmark = pytest.mark.gpu


def _two():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")


def test_sharded_batches_equal_one_device(built):
    _two()
    from klampt_b200.engine import Engine
    w = synth.world_c2(2, n_obstacles=60)
    one, two = Engine(w, device=0), Engine(w, device=[0, 1])
    assert two.num_devices() == 2 and one.num_devices() == 1
    Q = synth.sample_configs(w.robot, 150_003, 8)                       # ragged: not a multiple of the shard alignment
    f1, p1 = one.feasible_batch(Q, return_pairs=True)
    f2, p2 = two.feasible_batch(Q, return_pairs=True)
    assert np.array_equal(f1, f2)
    assert np.array_equal(p1[:, 0] >= 0, p2[:, 0] >= 0)                  # which pair is named first is order dependent; that one is named is not
    assert np.array_equal(np.unpackbits(two.feasible_batch_bits(Q), bitorder="little")[:len(Q)], f1)
    assert np.array_equal(two.feasible_batch(Q.astype(np.float32)), one.feasible_batch(Q.astype(np.float32)))
    ok = Q[f1 == 1]
    A, B = ok[:20_000], ok[20_000:40_000]
    v1, n1 = one.edges_visible_batch(A, B, eps=0.02)
    v2, n2 = two.edges_visible_batch(A, B, eps=0.02)
    assert np.array_equal(v1, v2) and np.array_equal(n1, n2)
    assert np.array_equal(np.unpackbits(two.edges_visible_batch_bits(A, B, eps=0.02), bitorder="little")[:len(A)], v1)
    d1 = one.distance_batch_ex(Q[:30_000], upper_bound=0.3, include_self=True)
    d2 = two.distance_batch_ex(Q[:30_000], upper_bound=0.3, include_self=True)
    assert np.array_equal(d1[0], d2[0]) and np.array_equal(d1[1], d2[1])
    np.testing.assert_array_equal(d1[2], d2[2])
    # rays are cast in contiguous blocks on the two devices (every replica runs FK for the configuration itself)
    rng = np.random.default_rng(3)
    src = rng.uniform([-3, -3, 0.2], [3, 3, 3], (60_001, 3))
    rays = np.hstack([src, rng.uniform([-1, -1, 0], [1, 1, 1.5], (60_001, 3)) - src])
    r1, r2 = one.raycast_batch(Q[0], rays), two.raycast_batch(Q[0], rays)
    assert all(np.array_equal(a, b) for a, b in zip(r1, r2)) and (r1[0] >= 0).mean() > 0.3
    # small batches stay on the first device; the threshold is an option
    two.set_option("multi_min", 16)
    assert np.array_equal(two.feasible_batch(Q[:1000]), f1[:1000])
    st = two.stats()
    assert st["configs_checked"] >= len(Q)
    # the work really was split: both devices launched kernels
    two.reset_stats()
    two.feasible_batch(Q)
    assert two.stats()["kernel_launches"] >= 2 * one.stats()["kernel_launches"] / max(1, one.stats()["kernel_launches"]) and two.stats()["configs_checked"] == len(Q)


def test_dynamic_cloud_is_rebuilt_on_every_device(built):
    _two()
    from klampt_b200.engine import Engine
    from klampt_b200.worldspec import GeomSpec
    w = synth.world_c1()
    g = w.add_geom(GeomSpec.dynamic_cloud(50_000, radius=0.0, margin=0.004))
    w.terrains.append(g)
    one, two = Engine(w, device=0), Engine(w, device=[0, 1])
    rng = np.random.default_rng(4)
    pts = rng.uniform([0.45, -0.3, 0.2], [0.8, 0.3, 0.9], size=(30_000, 3))      # a block of points beside the arm
    one.update_pointcloud(g, pts); two.update_pointcloud(g, pts)
    Q = synth.sample_configs(w.robot, 40_000, 3)
    a, b = one.feasible_batch(Q), two.feasible_batch(Q)
    base = Engine(synth.world_c1()).feasible_batch(Q)
    assert np.array_equal(a, b) and 0 < a.sum() < base.sum()                 # the cloud is there on both devices, and it matters
