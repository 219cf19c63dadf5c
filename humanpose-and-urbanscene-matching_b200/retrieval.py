"""Query-versus-database drivers, sharded over the ranks of one node (torch.distributed; NCCL on GPUs).

The database shards by rows (descriptors, C2), by images (C3) or by poses (C5); queries are replicated; every rank
runs the single-GPU kernels on its shard and the only exchange is one small all-gather of per-rank partial answers
(SURVEY.md 8e):
    kNN-2        [nq,2] (idx, dist) per rank  -> lexicographic (dist, global idx) merge kernel + ratio kernel
    retrieval    per-image survivor counts     -> concatenation in image order
    pose top-k   k (score, global idx) per rank -> top-k of the union
The compute callables are injectable so the sharding / gather plumbing is testable on CPU ranks (gloo) with the
oracle standing in for the kernels (tests/test_sharded_host.py); the defaults are the CUDA engine.
"""
import torch
import torch.distributed as dist

from . import engine


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, nranks, align=1):
    """Rows [lo, hi) of an n-row database owned by `rank`: equal `align`-aligned slabs, the tail on the last ranks."""
    per = -(-n // nranks)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_segments(seg_offsets, rank, nranks):
    """Images [lo, hi) owned by `rank` (each image's descriptor block stays whole on one rank)."""
    nseg = len(seg_offsets) - 1
    return shard_range(nseg, rank, nranks)


def globalize(idx, row_offset):
    """Local train indices -> database-global ones; empty slots (-1) stay empty."""
    return torch.where(idx >= 0, idx + row_offset, idx)


def allgather(t):
    rank, n = world()
    if n == 1:
        return t.unsqueeze(0)
    t = t.contiguous()
    out = torch.empty((n * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)  # concatenated form
    dist.all_gather_into_tensor(out, t)
    return out.view((n,) + tuple(t.shape))


def knn2_sharded(q, t_shard, row_offset, kind="l2", ratio=0.8, knn_fn=None, merge_fn=None, ratio_fn=None):
    """Global kNN-2 + Lowe ratio over a row-sharded database.  Every rank returns the same (idx, dist, keep, count)."""
    if knn_fn is None:
        knn_fn = engine.knn2_l2 if kind == "l2" else engine.knn2_hamming
    merge_fn = merge_fn or engine.knn2_merge
    ratio_fn = ratio_fn or engine.ratio_filter
    idx, d = knn_fn(q, t_shard)
    gi, gd = allgather(globalize(idx, row_offset)), allgather(d)
    if gi.shape[0] > 1:
        idx, d = merge_fn(gi, gd)
    else:
        idx, d = gi[0], gd[0]
    keep, count = ratio_fn(idx, d, ratio)
    return idx, d, keep, count


def image_scores_sharded(q, db_shard, seg_offsets_shard, ratio=0.8, score_fn=None):
    """Per-image ratio-survivor counts for a database sharded by images: [total images] int32 on every rank.
    Ranks may own different numbers of images; counts are padded to the largest shard for the gather."""
    score_fn = score_fn or engine.knn2_hamming_segmented
    local = score_fn(q, db_shard, seg_offsets_shard, ratio)
    rank, n = world()
    if n == 1:
        return local
    sizes = allgather(torch.tensor([local.numel()], dtype=torch.int64, device=local.device)).flatten().tolist()
    padded = torch.zeros((max(sizes),), dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    g = allgather(padded)
    return torch.cat([g[r, :sizes[r]] for r in range(n)])


def pose_topk_sharded(query, db_shard, row_offset, k=16, match_fn=None, topk_fn=None):
    """k best-scoring matching poses of a pose database sharded by rows: (scores [k], global indices [k], matches)."""
    match_fn = match_fn or engine.pose_match_batch
    topk_fn = topk_fn or engine.pose_topk
    match, score = match_fn(query, db_shard)
    ts, ti, cnt = topk_fn(match, score, k, row_offset)
    gs, gi, gc = allgather(ts), allgather(ti), allgather(cnt)
    s, i = gs.flatten(), gi.flatten()
    s = torch.where(i >= 0, s, torch.full_like(s, float("inf")))
    # lexicographic (score, index): stable sort by index first, then by score
    by_idx = torch.argsort(i, stable=True)
    order = by_idx[torch.argsort(s[by_idx], stable=True)][:k]
    return s[order], i[order], gc.sum()


def scene_retrieval(query_desc, query_kp, db_desc, db_kp, seg_offsets, ratio=0.8, min_match=16, thresh=5.0, nhyp=2000,
                    confidence=0.995, seed=0, query_pose=None, db_poses=None, query_size=None):
    """One query image against an image database, the whole of features.match_and_draw + validate_homography per
    database image (features.py:57-105,:177-211) without leaving the device: segmented Hamming kNN-2 + ratio ->
    per-image match lists -> batched RANSAC in both directions -> homography sanity flags.

    Returns a dict of device tensors, one entry per database image: n_matches, n_inliers, H (model->input, i.e.
    query->image), H2 (image->query), valid, plus the match lists (pair_offsets, src, dst, mask).  With `query_pose`
    [18,2], `db_poses` [nimg,18,2] (float64) and `query_size` = (width, height) of the query image the scene score of
    urban_scene.match_scene_multi (perspective correction + affine fit of pose and 5 background matches,
    urban_scene.py:80-145) is added as `score`: +inf where the homography is missing or fails validate_homography,
    to be compared with thresholds.AFFINE_TRANS_WHOLE_DISTANCE (matching.py:75)."""
    score, best = engine.knn2_hamming_segmented(query_desc, db_desc, seg_offsets, ratio, want_best_idx=True)
    offs, src, dst, sel = engine.segment_matches(best, score, query_kp, db_kp, min_match)
    H, mask, ninl, _ = engine.ransac_homography_batch(src, dst, offs, thresh, nhyp, confidence, seed)
    H2, _, _, _ = engine.ransac_homography_batch(dst, src, offs, thresh, nhyp, confidence, seed)
    valid = engine.validate_homography_batch(H)
    out = {"n_matches": score, "n_inliers": ninl, "H": H, "H2": H2, "valid": valid, "pair_offsets": offs, "src": src,
           "dst": dst, "mask": mask, "sel": sel}
    if query_pose is not None and db_poses is not None and query_size is not None:
        sc, picks = engine.scene_score_batch(src, dst, offs, mask, H2, query_pose, db_poses, query_size[0], query_size[1], seed)
        out["score"] = torch.where(valid.bool(), sc, torch.full_like(sc, float("inf")))
        out["picks"] = picks
    return out
