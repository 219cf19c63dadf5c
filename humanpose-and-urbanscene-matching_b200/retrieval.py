"""Query-versus-database drivers, sharded over the ranks of one node (torch.distributed; NCCL on GPUs).

The database shards by rows (descriptors, C2; knn2_query_sharded is the form with the queries sharded and the database
replicated), by images (C3) or by poses (C5); queries are replicated; every rank
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
    idx = globalize(idx, row_offset)
    rank, n = world()
    if n > 1:
        # ONE all-gather: (idx, dist) packed as 4 x 32 bit per query (the distance travels as its bit pattern)
        bits = torch.int32 if d.dtype in (torch.float32, torch.int32) else None
        if bits is None:
            raise TypeError("knn2_sharded: distances must be 32-bit, got %s" % d.dtype)
        g = allgather(torch.cat([idx.to(torch.int32), d.contiguous().view(torch.int32)], dim=1))  # [n, nq, 4]
        idx, d = merge_fn(g[:, :, :2].contiguous(), g[:, :, 2:].contiguous().view(d.dtype))
    keep, count = ratio_fn(idx, d, ratio)
    return idx, d, keep, count


def knn2_query_sharded(q_slice, t_full, nq_total, kind="l2", ratio=0.8, knn_fn=None, ratio_fn=None):
    """Global kNN-2 + Lowe ratio with the QUERIES sharded and the database replicated (the zero-exchange
    alternative): rank r matches its rows shard_range(nq_total, r, world) against the whole database; the only
    collective is the all-gather of the answers (16 B per query).  The form to use when the database fits every GPU and
    the queries are many: a rank's work is exactly 1/N of the single-GPU work, whereas row-sharding the database keeps the
    per-query costs (operand preparation, re-rank, the start-up of every running top-2 list) on every rank.
    Every rank returns the same (idx [nq_total,2], dist [nq_total,2], keep, count)."""
    if knn_fn is None:
        knn_fn = engine.knn2_l2 if kind == "l2" else engine.knn2_hamming
    ratio_fn = ratio_fn or engine.ratio_filter
    idx, d = knn_fn(q_slice, t_full)
    rank, n = world()
    if n > 1:
        per = -(-int(nq_total) // n)
        packed = torch.full((per, 4), -1, dtype=torch.int32, device=idx.device)
        packed[:idx.shape[0], :2] = idx
        packed[:idx.shape[0], 2:] = d.contiguous().view(torch.int32)
        g = allgather(packed).reshape(n * per, 4)[:int(nq_total)]
        idx, d = g[:, :2].contiguous(), g[:, 2:].contiguous().view(d.dtype)
    keep, count = ratio_fn(idx, d, ratio)
    return idx, d, keep, count


class StreamedL2Matcher:
    """kNN-2 (L2) of HOST descriptor arrays, the host->device copies pipelined under the compute.

    `knnMatch` is handed host arrays (features.py:59); copying 2 x 537 MB first and computing afterwards leaves the
    GPU idle for the whole transfer (19 ms; 1 M x 1 M on a B200: 132 ms copy-then-match, 124 ms streamed, 117 ms
    with the arrays already resident).  Here the queries are cut into row chunks and
    the database into a SMALL first chunk plus large ones; all copies are queued on a copy stream in the order
    q0, t0, q1 .. qN, t1, t2 .. and the compute stream sweeps every query chunk against t0 while t1 is still in
    flight, then against t1, ...  Only q0 and t0 are exposed.  The per-database-chunk top-2 lists are merged once at
    the end by the lexicographic merge kernel, which is exact because the global top-2 lies in the union of the
    per-chunk top-2s.  The database is cut in few pieces on purpose: every piece restarts the running top-2 lists of
    the sweep and adds a candidate pair per query to the exact re-rank (measured: a third piece that would keep the GPU
    busy while the rest of the database is still in flight costs 3.5 ms more than the idle time it removes).

    Callable as knn_fn(q_host, t_host) -> (idx [nq,2] int32, dist [nq,2] float32) on the device, so it drops into
    knn2_sharded.  Buffers, streams and events are created once; the returned tensors may alias them and are valid
    until the next call.
    """

    def __init__(self, nq, nt, d, device, q_chunk=None, t_first=1 << 17, t_chunk=1 << 20, mode=None, query_slices=False):
        self.nq, self.nt, self.d, self.device = int(nq), int(nt), int(d), torch.device(device)
        # query_slices: every rank is handed only ITS rows shard_range(nq, rank, world) of the (replicated) query array;
        # the slices are exchanged by one all-gather over NVLink, so a query row crosses PCIe once per node, not once per rank
        self.query_slices = bool(query_slices)
        rank, nranks = world()
        self.q_per = -(-self.nq // nranks) if self.query_slices else self.nq
        self.q_lo, self.q_hi = shard_range(self.nq, rank, nranks) if self.query_slices else (0, self.nq)
        self.mode = engine.L2_AUTO if mode is None else mode
        if q_chunk is None:
            # Query chunks are whole waves of the persistent kernel's 256-query units (chunks that are not a multiple of
            # SM count x 256 rows leave the last wave partly empty: 134 ms with 131072-row chunks, 124 ms with 7-wave ones),
            # and they GROW: 1, 1, 2, 3, 4, 6, 9 ... waves.  Only the first chunk's upload is exposed, and a chunk 1.5 x
            # the size of the one being matched against the first database chunk still lands before that match is done
            # (PCIe delivers ~107 k rows/ms, the sweep over 131072 database rows consumes ~71 k query rows/ms).
            wave = 256 * torch.cuda.get_device_properties(self.device).multi_processor_count
            cuts, k = [0], [1, 1, 2, 3, 4, 6, 9]
            while cuts[-1] < self.nq:
                w = k[len(cuts) - 1] if len(cuts) - 1 < len(k) else 9
                cuts.append(min(self.nq, cuts[-1] + w * wave))
            if len(cuts) > 2 and cuts[-1] - cuts[-2] < wave // 2:   # a sliver joins its neighbour
                del cuts[-2]
            self.q_bounds = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
        else:
            self.q_bounds = self._bounds(self.nq, 0, q_chunk)
        self.t_bounds = self._bounds(self.nt, t_first, t_chunk)
        dev = self.device
        nrows = self.q_per * nranks if self.query_slices else self.nq
        self.dq = torch.empty((nrows, self.d), dtype=torch.float32, device=dev)
        self.dq_slice = torch.empty((self.q_per, self.d), dtype=torch.float32, device=dev) if self.query_slices else None
        self.dt = torch.empty((self.nt, self.d), dtype=torch.float32, device=dev)
        qmax = max([b - a for a, b in self.q_bounds] or [0])
        tmax = max([b - a for a, b in self.t_bounds] or [0])
        nparts = max(len(self.t_bounds), 1)
        self.part_idx = torch.empty((nparts, self.nq, 2), dtype=torch.int32, device=dev)
        self.part_dist = torch.empty((nparts, self.nq, 2), dtype=torch.float32, device=dev)
        # database chunks after the first are matched against ALL queries in one call (they have all landed by then): one
        # operand preparation of the big chunk instead of one per query chunk
        self.ws = engine.knn2_l2_workspace(self.nq if len(self.t_bounds) > 1 or self.query_slices else qmax, tmax, self.d,
                                           self.mode, dev) if qmax and tmax else None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.q_ready = [torch.cuda.Event() for _ in self.q_bounds]
        self.t_ready = [torch.cuda.Event() for _ in self.t_bounds]
        self.h2d_bytes = ((self.q_hi - self.q_lo) + self.nt) * self.d * 4

    @staticmethod
    def _bounds(n, first, chunk):
        """Row ranges: `first` rows (if > 0), then ranges of `chunk` rows; a range of fewer than 2 rows joins its
        neighbour (a database chunk must be able to fill a top-2 list)."""
        chunk = max(int(chunk), 2)
        cuts = [0]
        if 0 < first < n:
            cuts.append(int(first))
        while cuts[-1] < n:
            cuts.append(min(n, cuts[-1] + chunk))
        if len(cuts) > 2 and cuts[1] - cuts[0] < 2:
            del cuts[1]
        if len(cuts) > 2 and cuts[-1] - cuts[-2] < 2:
            del cuts[-2]
        return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]

    def __call__(self, hq, ht):
        if tuple(hq.shape) != (self.q_hi - self.q_lo, self.d) or tuple(ht.shape) != (self.nt, self.d):
            raise ValueError("StreamedL2Matcher was built for [%d,%d] x [%d,%d]" % (self.q_hi - self.q_lo, self.d, self.nt, self.d))
        if hq.is_cuda or ht.is_cuda:
            raise ValueError("StreamedL2Matcher takes HOST tensors; use engine.knn2_l2 for device tensors")
        if self.nq == 0 or self.nt == 0:
            return engine.knn2_l2(hq.to(self.device), ht.to(self.device), mode=self.mode)
        cur = torch.cuda.current_stream(self.device)
        cs = self.copy_stream
        cs.wait_stream(cur)  # the previous call's kernels no longer read the device copies
        nqc, ntc = len(self.q_bounds), len(self.t_bounds)
        order = [("q", 0), ("t", 0)] + [("q", i) for i in range(1, nqc)] + [("t", j) for j in range(1, ntc)]
        if self.query_slices:
            order = [("qs", 0)] + [("t", j) for j in range(ntc)]
        with torch.cuda.stream(cs):
            for kind, k in order:
                if kind == "qs":
                    self.dq_slice[:self.q_hi - self.q_lo].copy_(hq, non_blocking=True)
                    self.q_ready[0].record(cs)
                elif kind == "q":
                    a, b = self.q_bounds[k]
                    self.dq[a:b].copy_(hq[a:b], non_blocking=True)
                    self.q_ready[k].record(cs)
                else:
                    a, b = self.t_bounds[k]
                    self.dt[a:b].copy_(ht[a:b], non_blocking=True)
                    self.t_ready[k].record(cs)
        if self.query_slices:
            cur.wait_event(self.q_ready[0])
            dist.all_gather_into_tensor(self.dq, self.dq_slice)  # rank r's rows land at [r * q_per, (r + 1) * q_per)
        for j, (ta, tb) in enumerate(self.t_bounds):
            cur.wait_event(self.t_ready[j])
            if j > 0 or self.query_slices:   # every query row is on the device: one call per database chunk
                engine.knn2_l2(self.dq[:self.nq], self.dt[ta:tb], mode=self.mode, out=(self.part_idx[j], self.part_dist[j]),
                               workspace=self.ws)
            else:
                for i, (qa, qb) in enumerate(self.q_bounds):
                    cur.wait_event(self.q_ready[i])
                    pi, pd = self.part_idx[j, qa:qb], self.part_dist[j, qa:qb]
                    engine.knn2_l2(self.dq[qa:qb], self.dt[ta:tb], mode=self.mode, out=(pi, pd), workspace=self.ws)
            if ta:
                self.part_idx[j] = globalize(self.part_idx[j], ta)
        if ntc == 1:
            return self.part_idx[0], self.part_dist[0]
        return engine.knn2_merge(self.part_idx, self.part_dist)


class StreamedImageScorer:
    """Per-image ratio-survivor counts (retrieval, C3) for a HOST descriptor database, the upload pipelined under the
    compute: the database is cut into runs of whole images, every run is queued on a copy stream and scored by
    hpusm_knn2_hamming_segmented as soon as it has landed.  Images are independent, so there is nothing to merge.

    Callable as score_fn(q_host, db_host, seg_offsets, ratio) -> scores [nimg] int32 on the device, so it drops into
    image_scores_sharded.  `seg_offsets` (host or device int64, [nimg + 1]) is fixed at construction.  The returned
    tensor is an internal buffer, valid until the next call.
    """

    def __init__(self, nq, nbytes, seg_offsets, device, images_per_chunk=512, hamming_engine="auto"):
        self.device = torch.device(device)
        self.hamming_engine = hamming_engine
        offs = torch.as_tensor(seg_offsets, dtype=torch.int64).cpu()
        self.nimg = int(offs.numel() - 1)
        self.nt = int(offs[-1]) if offs.numel() else 0
        self.nq, self.nb = int(nq), int(nbytes)
        step = max(int(images_per_chunk), 1)
        self.chunks = []  # (first image, last image + 1, first row, last row + 1, local offsets on the device)
        for i0 in range(0, self.nimg, step):
            i1 = min(self.nimg, i0 + step)
            r0, r1 = int(offs[i0]), int(offs[i1])
            self.chunks.append((i0, i1, r0, r1, (offs[i0:i1 + 1] - r0).to(self.device)))
        self.dq = torch.empty((self.nq, self.nb), dtype=torch.uint8, device=self.device)
        self.ddb = torch.empty((self.nt, self.nb), dtype=torch.uint8, device=self.device)
        self.scores = torch.empty((self.nimg,), dtype=torch.int32, device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event() for _ in self.chunks]
        self.q_ready = torch.cuda.Event()

    def __call__(self, hq, hdb, seg_offsets=None, ratio=0.8):
        if tuple(hq.shape) != (self.nq, self.nb) or tuple(hdb.shape) != (self.nt, self.nb):
            raise ValueError("StreamedImageScorer was built for [%d,%d] x [%d,%d]" % (self.nq, self.nb, self.nt, self.nb))
        if hq.is_cuda or hdb.is_cuda:
            raise ValueError("StreamedImageScorer takes HOST tensors; use engine.knn2_hamming_segmented for device tensors")
        cur = torch.cuda.current_stream(self.device)
        cs = self.copy_stream
        cs.wait_stream(cur)  # the previous call's kernels no longer read the device copies
        with torch.cuda.stream(cs):
            self.dq.copy_(hq, non_blocking=True)
            self.q_ready.record(cs)
            for c, (i0, i1, r0, r1, _) in enumerate(self.chunks):
                self.ddb[r0:r1].copy_(hdb[r0:r1], non_blocking=True)
                self.ready[c].record(cs)
        cur.wait_event(self.q_ready)
        for c, (i0, i1, r0, r1, local) in enumerate(self.chunks):
            cur.wait_event(self.ready[c])
            if r1 == r0:  # a run of empty images
                self.scores[i0:i1] = 0
                continue
            self.scores[i0:i1] = engine.knn2_hamming_segmented(self.dq, self.ddb[r0:r1], local, ratio,
                                                               engine=self.hamming_engine)
        return self.scores


class StreamedHomographyBatch:
    """cv2.findHomography(.., RANSAC, ..) for a batch of image pairs whose putative matches live in HOST memory
    (features.py:66,:69 called once per pair), the upload pipelined under the compute: the batch is cut into runs of
    whole pairs, every run is queued on a copy stream and solved as soon as it has landed.  `pair_index0` keeps each
    run on its rows of the (seed, pair, hypothesis) sample table, so the answer is the one-call answer bit for bit.

    __call__(src_host, dst_host) -> (H [npairs,3,3] f64, mask [total] u8, ninl [npairs] i32) on the device (internal
    buffers, valid until the next call).
    """

    def __init__(self, pair_offsets, device, thresh=5.0, nhyp=2000, confidence=0.995, seed=0, pairs_per_chunk=None,
                 sampler="counter"):
        self.device = torch.device(device)
        offs = torch.as_tensor(pair_offsets, dtype=torch.int64).cpu()
        self.npairs, self.total = int(offs.numel() - 1), int(offs[-1])
        self.args = (float(thresh), int(nhyp), float(confidence), int(seed))
        self.sampler = sampler
        # The scoring kernel runs one CTA per pair, three CTAs per SM: a run that is not a whole number of such waves ends
        # in a partly empty one (512-pair runs on 148 SMs: 1.15 waves, i.e. 58 % of two -- measured as 5.5e8 hyps/s end to
        # end against 7.9e8 resident).  Default: a first run of ONE wave (only its upload is exposed), then runs of two.
        wave = 3 * torch.cuda.get_device_properties(self.device).multi_processor_count
        if pairs_per_chunk is None:
            cuts = [0, min(wave, self.npairs)]
            while cuts[-1] < self.npairs:
                cuts.append(min(self.npairs, cuts[-1] + 2 * wave))
        else:
            step = max(int(pairs_per_chunk), 1)
            cuts = list(range(0, self.npairs, step)) + [self.npairs]
        self.chunks = []
        for p0, p1 in zip(cuts[:-1], cuts[1:]):
            if p1 <= p0:
                continue
            r0, r1 = int(offs[p0]), int(offs[p1])
            self.chunks.append((p0, p1, r0, r1, (offs[p0:p1 + 1] - r0).to(self.device)))
        dev = self.device
        self.src = torch.empty((self.total, 2), dtype=torch.float32, device=dev)
        self.dst = torch.empty((self.total, 2), dtype=torch.float32, device=dev)
        self.H = torch.empty((self.npairs, 3, 3), dtype=torch.float64, device=dev)
        self.mask = torch.empty((self.total,), dtype=torch.uint8, device=dev)
        self.ninl = torch.empty((self.npairs,), dtype=torch.int32, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in self.chunks]

    def __call__(self, hsrc, hdst):
        if tuple(hsrc.shape) != (self.total, 2) or tuple(hdst.shape) != (self.total, 2):
            raise ValueError("StreamedHomographyBatch was built for %d matches" % self.total)
        if hsrc.is_cuda or hdst.is_cuda:
            raise ValueError("StreamedHomographyBatch takes HOST tensors; use engine.ransac_homography_batch for device tensors")
        cur = torch.cuda.current_stream(self.device)
        cs = self.copy_stream
        cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            for c, (p0, p1, r0, r1, _) in enumerate(self.chunks):
                self.src[r0:r1].copy_(hsrc[r0:r1], non_blocking=True)
                self.dst[r0:r1].copy_(hdst[r0:r1], non_blocking=True)
                self.ready[c].record(cs)
        thresh, nhyp, conf, seed = self.args
        for c, (p0, p1, r0, r1, local) in enumerate(self.chunks):
            cur.wait_event(self.ready[c])
            H, mask, ninl, _ = engine.ransac_homography_batch(self.src[r0:r1], self.dst[r0:r1], local, thresh, nhyp, conf,
                                                              seed=seed, pair_index0=p0, sampler=self.sampler)
            self.H[p0:p1], self.mask[r0:r1], self.ninl[p0:p1] = H, mask, ninl
        return self.H, self.mask, self.ninl


def image_scores_sharded(q, db_shard, seg_offsets_shard, ratio=0.8, score_fn=None, hamming_engine="auto"):
    """Per-image ratio-survivor counts for a database sharded by images: [total images] int32 on every rank.
    Ranks may own different numbers of images; counts are padded to the largest shard for the gather."""
    if score_fn is None:
        local = engine.knn2_hamming_segmented(q, db_shard, seg_offsets_shard, ratio, engine=hamming_engine)
    else:
        local = score_fn(q, db_shard, seg_offsets_shard, ratio)
    rank, n = world()
    if n == 1:
        return local
    sizes = allgather(torch.tensor([local.numel()], dtype=torch.int64, device=local.device)).flatten().tolist()
    padded = torch.zeros((max(sizes),), dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    g = allgather(padded)
    return torch.cat([g[r, :sizes[r]] for r in range(n)])


def pose_topk_sharded(query, db_shard, row_offset, k=16, match_fn=None, topk_fn=None, merge_fn=None):
    """k best-scoring matching poses of a pose database sharded by rows: (scores [k], global indices [k], matches).

    One all-gather: every rank packs its k (score, index) pairs and its match count into one int64 vector.  The union
    is reduced by `merge_fn` (default on CUDA tensors: one launch of the top-k merge kernel; otherwise torch sorts)."""
    match_fn = match_fn or engine.pose_match_batch
    topk_fn = topk_fn or engine.pose_topk
    match, score = match_fn(query, db_shard)
    ts, ti, cnt = topk_fn(match, score, k, row_offset)
    rank, n = world()
    if n == 1:
        return ts, ti, cnt.sum()
    packed = torch.cat([ts.contiguous().view(torch.int32).to(torch.int64), ti.to(torch.int64), cnt.reshape(1).to(torch.int64)])
    g = allgather(packed)  # [n, 2k + 1]
    s = g[:, :k].to(torch.int32).contiguous().view(torch.float32).flatten()
    i = g[:, k:2 * k].contiguous().flatten()
    total = g[:, 2 * k].sum()
    if merge_fn is None and s.is_cuda and s.numel() <= 4096:
        merge_fn = engine.pose_topk_merge
    if merge_fn is not None:
        ms, mi = merge_fn(s, i, k)
        return ms, mi, total
    s = torch.where(i >= 0, s, torch.full_like(s, float("inf")))
    # lexicographic (score, index): stable sort by index first, then by score
    by_idx = torch.argsort(i, stable=True)
    order = by_idx[torch.argsort(s[by_idx], stable=True)][:k]
    return s[order], i[order], total



def scene_retrieval(query_desc, query_kp, db_desc, db_kp, seg_offsets, ratio=0.8, min_match=16, thresh=5.0, nhyp=2000,
                    confidence=0.995, seed=0, query_pose=None, db_poses=None, query_size=None, sampler="counter",
                    image_index0=0):
    """One query image against an image database, the whole of features.match_and_draw + validate_homography per
    database image (features.py:57-105,:177-211) without leaving the device: segmented Hamming kNN-2 + ratio ->
    per-image match lists -> batched RANSAC in both directions -> homography sanity flags.

    Returns a dict of device tensors, one entry per database image: n_matches, n_inliers, H (model->input, i.e.
    query->image), H2 (image->query), valid, plus the match lists (pair_offsets, src, dst, mask).  With `query_pose`
    [18,2], `db_poses` [nimg,18,2] (float64) and `query_size` = (width, height) of the query image the scene score of
    urban_scene.match_scene_multi (perspective correction + affine fit of pose and 5 background matches,
    urban_scene.py:80-145) is added as `score`: +inf where the homography is missing or fails validate_homography,
    to be compared with thresholds.AFFINE_TRANS_WHOLE_DISTANCE (matching.py:75).
    `image_index0`: index of this call's first image in the whole database (a database streamed in runs of images draws
    the RANSAC samples of the one-call run: the sample table is keyed on the image's global index)."""
    score, best = engine.knn2_hamming_segmented(query_desc, db_desc, seg_offsets, ratio, want_best_idx=True)
    offs, src, dst, sel = engine.segment_matches(best, score, query_kp, db_kp, min_match)
    H, mask, ninl, _ = engine.ransac_homography_batch(src, dst, offs, thresh, nhyp, confidence, seed, sampler=sampler,
                                                      pair_index0=image_index0)
    H2, _, _, _ = engine.ransac_homography_batch(dst, src, offs, thresh, nhyp, confidence, seed, sampler=sampler,
                                                 pair_index0=image_index0)
    valid = engine.validate_homography_batch(H)
    out = {"n_matches": score, "n_inliers": ninl, "H": H, "H2": H2, "valid": valid, "pair_offsets": offs, "src": src,
           "dst": dst, "mask": mask, "sel": sel}
    if query_pose is not None and db_poses is not None and query_size is not None:
        sc, picks = engine.scene_score_batch(src, dst, offs, mask, H2, query_pose, db_poses, query_size[0], query_size[1], seed)
        out["score"] = torch.where(valid.bool(), sc, torch.full_like(sc, float("inf")))
        out["picks"] = picks
    return out
