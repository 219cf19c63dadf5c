"""World-size-2 gloo run on CPU of the sharding / all-gather plumbing in retrieval.py.  The CUDA kernels cannot run
here, so the oracle stands in for them through retrieval's injectable callables; what is under test is the host
logic: shard ranges, index globalisation, gather layout, merge order, ragged image shards."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_fns():
    from oracle import knn_oracle

    def knn_fn(q, t):
        i, d = knn_oracle.knn2_hamming(q.numpy(), t.numpy())
        return torch.from_numpy(i), torch.from_numpy(d)

    def merge_fn(gi, gd):
        n, nq, _ = gi.shape
        i = gi.permute(1, 0, 2).reshape(nq, -1).numpy()
        d = gd.permute(1, 0, 2).reshape(nq, -1).numpy().astype(np.float64)
        d = np.where(i >= 0, d, np.inf)
        order = np.lexsort((np.where(i >= 0, i, 1 << 30), d), axis=1)[:, :2]
        rows = np.arange(nq)[:, None]
        oi, od = i[rows, order], np.where(i[rows, order] >= 0, d[rows, order], 0)
        return torch.from_numpy(oi.astype(np.int32)), torch.from_numpy(od.astype(np.int32))

    def ratio_fn(idx, d, ratio):
        keep = knn_oracle.ratio_keep(idx.numpy(), d.numpy(), ratio)
        return torch.from_numpy(keep.astype(np.uint8)), torch.tensor([int(keep.sum())], dtype=torch.int32)

    def score_fn(q, db, offs, ratio):
        s, _ = knn_oracle.segmented_scores(q.numpy(), db.numpy(), offs.numpy(), ratio)
        return torch.from_numpy(s)

    return knn_fn, merge_fn, ratio_fn, score_fn


def _worker(rank, nranks, port, out_dir):
    sys.path.insert(0, ROOT)
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
    rt = hp.retrieval
    knn_fn, merge_fn, ratio_fn, score_fn = _oracle_fns()
    # -- row-sharded kNN (C2 plumbing) with an uneven split and cross-shard ties
    q, t = hp.synth.orb_like(200, 1001, seed=4, planted=0.3)
    t[900] = t[10]  # equal rows on different ranks: the lower global index must win
    lo, hi = rt.shard_range(len(t), rank, nranks, align=128)
    idx, d, keep, count = rt.knn2_sharded(torch.from_numpy(q), torch.from_numpy(t[lo:hi]), lo, kind="hamming", knn_fn=knn_fn,
                                          merge_fn=merge_fn, ratio_fn=ratio_fn)
    # -- the same problem with the QUERIES sharded (uneven slices) and the database replicated
    nq2 = len(q) - 1   # odd: the last rank's slice is one row short and is padded for the gather
    qlo, qhi = rt.shard_range(nq2, rank, nranks)
    qi, qd, qkeep, qcount = rt.knn2_query_sharded(torch.from_numpy(q[qlo:qhi]), torch.from_numpy(t), nq2, kind="hamming",
                                                  knn_fn=knn_fn, ratio_fn=ratio_fn)
    # -- image-sharded retrieval (C3 plumbing) with ragged shards
    query, db, offs, planted = hp.synth.orb_database(7, per_image=64, seed=6, ragged=True)
    a, b = rt.shard_segments(offs, rank, nranks)
    local_offs = torch.from_numpy(offs[a:b + 1] - offs[a])
    scores = rt.image_scores_sharded(torch.from_numpy(query), torch.from_numpy(db[offs[a]:offs[b]]), local_offs, 0.8,
                                     score_fn=score_fn)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), idx=idx.numpy(), d=d.numpy(), keep=keep.numpy(),
             count=count.numpy(), scores=scores.numpy(), qidx=qi.numpy(), qd=qd.numpy(), qkeep=qkeep.numpy(),
             qcount=qcount.numpy())
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    import importlib
    rt = importlib.import_module("humanpose-and-urbanscene-matching_b200").retrieval
    for n in (0, 1, 127, 128, 1000, 1 << 20):
        for w in (1, 2, 3, 8):
            spans = [rt.shard_range(n, r, w, align=128) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert all(lo % 128 == 0 or lo == n for lo, _ in spans)


def test_streamed_matcher_chunk_bounds():
    """StreamedL2Matcher._bounds: the ranges tile [0, n) in order, and no range has fewer than 2 rows unless n does."""
    import importlib
    rt = importlib.import_module("humanpose-and-urbanscene-matching_b200").retrieval
    for n in (1, 2, 3, 700, 2049, 4097, 1 << 20):
        for first in (0, 1, 2, 2048, n, n + 5):
            for chunk in (1, 2, 1000, 4096, 1 << 20):
                b = rt.StreamedL2Matcher._bounds(n, first, chunk)
                assert b[0][0] == 0 and b[-1][1] == n
                assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
                assert all(hi - lo >= min(2, n) for lo, hi in b), (n, first, chunk, b)


def test_world2_gloo_sharded_equals_unsharded(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    import importlib
    hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
    from oracle import knn_oracle
    q, t = hp.synth.orb_like(200, 1001, seed=4, planted=0.3)
    t[900] = t[10]
    ri, rd = knn_oracle.knn2_hamming(q, t)
    query, db, offs, planted = hp.synth.orb_database(7, per_image=64, seed=6, ragged=True)
    rs, _ = knn_oracle.segmented_scores(query, db, offs, 0.8)
    for r in range(2):
        g = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        np.testing.assert_array_equal(g["idx"], ri)
        np.testing.assert_array_equal(g["d"], rd)
        np.testing.assert_array_equal(g["keep"].astype(bool), knn_oracle.ratio_keep(ri, rd, 0.8))
        np.testing.assert_array_equal(g["scores"], rs)
        np.testing.assert_array_equal(g["qidx"], ri[:-1])
        np.testing.assert_array_equal(g["qd"], rd[:-1])
        np.testing.assert_array_equal(g["qkeep"], g["keep"][:-1])
        assert int(g["qcount"][0]) == int(g["keep"][:-1].sum())
