"""N>1 host path on CPU: two gloo ranks shard a batch, run a stand-in feasibility function on their block and gather.
(The per-rank function is a plain numpy predicate here: the CUDA engine itself needs a GPU; what is under test is the
partition + gather logic every rank runs around it.)"""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    from klampt_b200.shard import ShardedRunner
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    Q = rng.uniform(-1, 1, size=(n, 7))
    A, B = Q, Q[::-1].copy()
    calls = []

    def feas(Qb):
        calls.append(len(Qb))
        return (Qb[:, 1] * Qb[:, 2] > -0.1).astype(np.uint8)

    def vis(Ab, Bb):
        return ((Ab[:, 0] + Bb[:, 0]) > 0).astype(np.uint8)

    run = ShardedRunner(feas, vis)
    f = run.feasible_batch(Q)
    v = run.visible_batch(A, B, block=64)
    np.save(os.path.join(out_dir, "f%d.npy" % rank), f)
    np.save(os.path.join(out_dir, "v%d.npy" % rank), v)
    np.save(os.path.join(out_dir, "c%d.npy" % rank), np.array(calls))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    n, world = 10007, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    Q = rng.uniform(-1, 1, size=(n, 7))
    want_f = (Q[:, 1] * Q[:, 2] > -0.1).astype(np.uint8)
    want_v = ((Q[:, 0] + Q[::-1][:, 0]) > 0).astype(np.uint8)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("f%d.npy" % r)), want_f)      # every rank holds the full result
        assert np.array_equal(np.load(tmp_path / ("v%d.npy" % r)), want_v)
    c0, c1 = np.load(tmp_path / "c0.npy"), np.load(tmp_path / "c1.npy")
    assert c0.sum() + c1.sum() == n and abs(int(c0.sum()) - int(c1.sum())) <= 1  # each rank only touched its own block
