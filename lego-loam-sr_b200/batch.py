"""Batched replay of independent sequences (BASELINE config 5): replicas only.

One registration is sequential and stays on one GPU. What shards is the set of independent sequences:
`seq -> rank = seq mod world_size`; inside a rank every sequence has its own MapOptimization mirror
(= its own ll_handle and CUDA stream) driven by its own host thread, so the latency-bound phases of one
sequence (launch gaps, the single-block solve) overlap with the kernels of the others. No data-path
collective exists; torch.distributed (NCCL on GPUs, gloo in CPU tests) only gathers the fixed-size result
records {seq, scan, pose[6], iter_num, sel_num, status, min_lambda, cf_mean}.
"""
import threading

import numpy as np

RECORD_FIELDS = ("seq", "scan", "rx", "ry", "rz", "tx", "ty", "tz", "iter_num", "sel_num", "status", "min_lambda",
                 "cf_mean")
RECORD_WIDTH = len(RECORD_FIELDS)


def shard_sequences(n_sequences, world_size, rank):
    """Sequence ids owned by `rank` (round robin, so that every rank gets the same count +- 1)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, n_sequences, world_size))


def make_record(seq, scan, pose, res):
    return np.array([seq, scan, pose[0], pose[1], pose[2], pose[3], pose[4], pose[5], res.iter_num, res.sel_num,
                     res.status, res.min_lambda, res.cf_mean], np.float64)


def gather_records(local_records, n_sequences, n_scans, world_size, rank, dist=None, device=None):
    """All ranks' records as one [n_sequences, n_scans, RECORD_WIDTH] array (every rank gets it).

    local_records: {seq: [n_scans, RECORD_WIDTH]} for the sequences this rank owns. With world_size 1 (or
    dist None) nothing is communicated. Ranks may own different numbers of sequences, so each rank
    contributes a padded [max_owned, n_scans, W] block.
    """
    out = np.full((n_sequences, n_scans, RECORD_WIDTH), np.nan, np.float64)
    if world_size == 1 or dist is None:
        for s, rec in local_records.items():
            out[s] = rec
        return out
    import torch
    max_owned = (n_sequences + world_size - 1) // world_size
    block = np.full((max_owned, n_scans, RECORD_WIDTH), np.nan, np.float64)
    for k, s in enumerate(shard_sequences(n_sequences, world_size, rank)):
        block[k] = local_records[s]
    t = torch.from_numpy(block)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(parts, t)
    for r in range(world_size):
        arr = parts[r].cpu().numpy()
        for k, s in enumerate(shard_sequences(n_sequences, world_size, r)):
            out[s] = arr[k]
    return out


def replay_sequences(sequences, make_node, feed, n_threads=None, nodes=None):
    """Replay {seq_id: data} concurrently: one node (MapOptimization mirror) and one host thread per sequence.

    make_node() -> node; feed(node, data, k) -> (pose, res) registers scan k. Returns {seq: [n, W] records}.
    ctypes releases the GIL inside every library call, so the threads overlap on the GPU (one CUDA stream
    per handle). `nodes` = {seq: node} created beforehand keeps handle creation (cudaMalloc, which synchronises
    the device) out of a timed region; such nodes are left open for the caller to close.
    """
    ids = sorted(sequences)
    records = {}
    errors = []

    def work(my_ids):
        try:
            for s in my_ids:
                node = nodes[s] if nodes is not None else make_node()
                data = sequences[s]
                n = len(data["frames"])
                rec = np.zeros((n, RECORD_WIDTH), np.float64)
                for k in range(n):
                    pose, res = feed(node, data, k)
                    rec[k] = make_record(s, k, pose, res)
                records[s] = rec
                if nodes is None:
                    node.close()
        except Exception as e:  # surfaced to the caller below
            errors.append(e)

    n_threads = n_threads or len(ids) or 1
    groups = [ids[i::n_threads] for i in range(n_threads)]
    threads = [threading.Thread(target=work, args=(g,)) for g in groups if g]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return records
