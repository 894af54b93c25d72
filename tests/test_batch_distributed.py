"""Host-side logic of the multi-GPU path (replicas only), exercised on CPU with the gloo
backend and world_size 2: sequence sharding, per-sequence replay threads, and the result gather that NCCL
performs on the GPU box. The "node" here is the CPU oracle's replay (this is a test: the oracle is the
checker), so that the gathered poses can be compared with a single-process run bit for bit."""
import os
import socket

import numpy as np
import pytest

from conftest import SMALL_WORLD, upstream_params


def test_shard_sequences_partition():
    from lego_loam_sr_b200 import batch
    for n, w in ((64, 1), (64, 2), (64, 4), (64, 8), (5, 3), (2, 8)):
        owned = [batch.shard_sequences(n, w, r) for r in range(w)]
        flat = sorted(s for o in owned for s in o)
        assert flat == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    with pytest.raises(ValueError):
        batch.shard_sequences(4, 2, 2)


class _OracleNode:
    def __init__(self, orc):
        self.rp = orc.Replay(upstream_params(orc, iter_count_thres=4), window=50)

    def close(self):
        self.rp.close()


def _feed(node, data, k):
    f = data["frames"][k]
    res, nc, ns = node.rp.step(data["odom"][k], f["corner_scan"], f["corner_scan_ds"], f["surf_last_ds"],
                               f["outlier_last_ds"], f["surf_total_ds"])
    return np.array(res.pose, np.float32), res


def _make_sequences(ids, n_scans):
    from lego_loam_sr_b200 import synth
    small = dict(n_blocks=8, n_poles=50, half_x=25, half_z=35)
    return {s: synth.make_sequence(seed=100 + s, n_scans=n_scans, step=0.7, world_kw=small) for s in ids}


def _worker(rank, world, port, n_seq, n_scans, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import orc
    from lego_loam_sr_b200 import batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = batch.shard_sequences(n_seq, world, rank)
    seqs = _make_sequences(ids, n_scans)
    recs = batch.replay_sequences(seqs, lambda: _OracleNode(orc), _feed, n_threads=2)
    allrec = batch.gather_records(recs, n_seq, n_scans, world, rank, dist=dist)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), allrec)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(orc, tmp_path):
    import torch.multiprocessing as mp
    from lego_loam_sr_b200 import batch
    n_seq, n_scans = 3, 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_seq, n_scans, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1) and not np.isnan(r0).any()
    # single process, world 1
    seqs = _make_sequences(range(n_seq), n_scans)
    recs = batch.replay_sequences(seqs, lambda: _OracleNode(orc), _feed, n_threads=1)
    single = batch.gather_records(recs, n_seq, n_scans, 1, 0)
    assert np.array_equal(single, r0)  # gathered poses identical (bitwise) for world = 1 and 2
    assert (r0[:, :, 0] == np.arange(n_seq)[:, None]).all() and (r0[:, :, 1] == np.arange(n_scans)[None, :]).all()
    assert (r0[:, 1:, batch.RECORD_FIELDS.index("status")] == 0).all()
