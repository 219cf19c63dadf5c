"""Parity of the CUDA path (through the C ABI of libhpusm.so) with the CPU oracle and the golden vectors
recorded from the reference.  Needs a B200: run with `-m gpu`.

Bars (BASELINE.json north_star): match indices bit-exact for Hamming and for L2 after the FP32 re-rank
(ties within 1e-5 relative distance tolerated for general floats); homographies / transforms within 1e-4;
inlier sets identical under the shared hypothesis seed."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, ROOT
from oracle import knn_oracle, pose_oracle, ransac_oracle

pytestmark = pytest.mark.gpu


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def assert_l2_parity(idx, dist, ref_idx, ref_dist, q, t, rel=1e-5):
    """Indices equal, except where the competing distances are within `rel` (stated tie tolerance)."""
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    np.testing.assert_allclose(dist, ref_dist, rtol=2e-6, atol=0)
    bad = np.nonzero((idx != ref_idx).any(1))[0]
    for r in bad:
        for c in range(2):
            if idx[r, c] != ref_idx[r, c]:
                da = np.sqrt(((q[r].astype(np.float64) - t[idx[r, c]].astype(np.float64)) ** 2).sum())
                db = np.sqrt(((q[r].astype(np.float64) - t[ref_idx[r, c]].astype(np.float64)) ** 2).sum())
                assert abs(da - db) <= rel * max(da, db), (r, c, da, db)


# ---- K2 Hamming ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["knn_hamming_256.npz", "knn_hamming_512.npz"])
def test_hamming_golden_cv2(hpusm, cuda_device, name):
    g = load_golden(name)
    idx, dist = hpusm.knn2_hamming(dev(g["q"], cuda_device), dev(g["t"], cuda_device))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["idx"])
    np.testing.assert_array_equal(dist.cpu().numpy(), g["dist"])


def test_hamming_scene_pair_orb(hpusm, cuda_device):
    """C1 with ORB: 14138 x 9130 descriptors of pisa1 / pisa2, cv2.BFMatcher result recorded from the reference."""
    g = load_golden("scene_pisa1_pisa2_orb.npz")
    idx, dist = hpusm.knn2_hamming(dev(g["desc_model"], cuda_device), dev(g["desc_input"], cuda_device))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["knn_idx"])
    np.testing.assert_array_equal(dist.cpu().numpy().astype(np.float32), g["knn_dist"])
    keep, count = hpusm.ratio_filter(idx, dist, 0.8)
    assert int(count.item()) == len(g["p_model"]) == 250
    p_q, p_t, sel, cnt = hpusm.gather_matches(dev(g["kp_model"], cuda_device), dev(g["kp_input"], cuda_device), idx, keep)
    n = int(cnt.item())
    np.testing.assert_array_equal(p_q[:n].cpu().numpy(), g["p_model"].reshape(-1, 2))
    np.testing.assert_array_equal(p_t[:n].cpu().numpy(), g["p_input"].reshape(-1, 2))


def test_hamming_second_scene_pair_orb(hpusm, cuda_device):
    """taj1 x taj2 through the reference's ORB detector (3626 x 5108 descriptors): knnMatch and filter_matches."""
    g = load_golden("scene_taj1_taj2_orb_knn.npz")
    idx, dist = hpusm.knn2_hamming(dev(g["desc_model"], cuda_device), dev(g["desc_input"], cuda_device))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["knn_idx"])
    np.testing.assert_array_equal(dist.cpu().numpy().astype(np.float32), g["knn_dist"])
    keep, count = hpusm.ratio_filter(idx, dist, 0.8)
    n = len(g["p_model"])
    assert int(count.item()) == n
    p_q, p_t, sel, cnt = hpusm.gather_matches(dev(g["kp_model"], cuda_device), dev(g["kp_input"], cuda_device), idx, keep)
    np.testing.assert_array_equal(p_q[:n].cpu().numpy(), g["p_model"].reshape(-1, 2))
    np.testing.assert_array_equal(p_t[:n].cpu().numpy(), g["p_input"].reshape(-1, 2))


@pytest.mark.parametrize("nq,nt,nb", [(1, 1, 32), (3, 2, 32), (513, 129, 32), (1000, 5000, 32), (37, 70000, 32),
                                      (4100, 300, 64), (5, 0, 32), (0, 5, 32)])
def test_hamming_edge_shapes(hpusm, cuda_device, nq, nt, nb):
    rng = np.random.default_rng(nq * 7 + nt)
    q = rng.integers(0, 256, (nq, nb), dtype=np.uint8)
    t = rng.integers(0, 4, (nt, nb), dtype=np.uint8) if nt < 100 else rng.integers(0, 256, (nt, nb), dtype=np.uint8)
    idx, dist = hpusm.knn2_hamming(dev(q, cuda_device), dev(t, cuda_device))
    ri, rd = knn_oracle.knn2_hamming(q, t)
    np.testing.assert_array_equal(idx.cpu().numpy(), ri)
    np.testing.assert_array_equal(dist.cpu().numpy(), rd)


def test_hamming_ties_lowest_index(hpusm, cuda_device):
    q = np.zeros((600, 32), np.uint8)
    t = np.zeros((3000, 32), np.uint8)  # every distance is 0: the answer must be train rows 0 and 1
    idx, dist = hpusm.knn2_hamming(dev(q, cuda_device), dev(t, cuda_device))
    assert (idx.cpu().numpy() == np.array([0, 1])).all() and (dist.cpu().numpy() == 0).all()


def test_hamming_segmented_retrieval(hpusm, cuda_device):
    """C3 shape at small scale: ragged database images incl. nearly empty ones."""
    query, db, offs, planted = hpusm.synth.orb_database(60, per_image=300, seed=5, ragged=True)
    score, best = hpusm.knn2_hamming_segmented(dev(query, cuda_device), dev(db, cuda_device), dev(offs, cuda_device), 0.8,
                                               want_best_idx=True)
    rs, rb = knn_oracle.segmented_scores(query, db, offs, 0.8)
    np.testing.assert_array_equal(score.cpu().numpy(), rs)
    np.testing.assert_array_equal(best.cpu().numpy(), rb)
    top = np.argsort(-rs, kind="stable")[:len(planted)]
    assert set(top.tolist()) & set(planted.tolist())


def _ragged_orb_case(seed, nimg, lo, hi, nq):
    """Images of lo..hi rows with the edge lengths of the tensor engine's 128-row tiles, empty and 1-row images, exact
    duplicates of query rows (distance 0 twice: the index decides) and near duplicates."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(lo, hi + 1, nimg).astype(np.int64)
    sizes[:12] = [0, 1, 2, 127, 128, 129, 255, 256, 257, 0, 64, 65]
    rng.shuffle(sizes)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    db = rng.integers(0, 256, (int(offs[-1]), 32), dtype=np.uint8)
    query = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    for im in rng.choice(nimg, nimg // 3, replace=False):
        n = int(sizes[im])
        if n < 2:
            continue
        rows = rng.choice(n, min(n, nq) // 2, replace=False)
        src = rng.choice(nq, len(rows), replace=False)
        d = query[src].copy()
        d[:, rng.integers(0, 32)] ^= np.uint8(1 << rng.integers(0, 8))      # one bit away
        db[offs[im] + rows] = d
        db[offs[im] + rows[: len(rows) // 4]] = query[src[: len(rows) // 4]]  # exact copies
        if n > 4:
            db[offs[im] + rows[0]] = db[offs[im] + rows[1]] = query[src[0]]   # two rows at distance 0
    return query, db, offs


@pytest.mark.parametrize("nq", [128, 700, 1024])
def test_hamming_tensor_engine_equals_oracle(hpusm, cuda_device, nq):
    """The tcgen05 engine of the segmented matcher (keys 128 * Hamming + column in the s32 accumulators) against the CPU
    restatement of BFMatcher(NORM_HAMMING).knnMatch + filter_matches per image (features.py:59, :45-55): ragged images,
    images shorter than / equal to / one past a tile, empty images, ties."""
    query, db, offs = _ragged_orb_case(3 + nq, 70, 200, 900, nq)
    score, best = hpusm.knn2_hamming_segmented(dev(query, cuda_device), dev(db, cuda_device), dev(offs, cuda_device), 0.8,
                                               want_best_idx=True, engine="tc")
    rs, rb = knn_oracle.segmented_scores(query, db, offs, 0.8)
    np.testing.assert_array_equal(score.cpu().numpy(), rs)
    np.testing.assert_array_equal(best.cpu().numpy(), rb)
    assert rs.sum() > 1000


def test_hamming_engines_agree(hpusm, cuda_device):
    """Both engines, many units per SM, a query count that is not a multiple of the 256-row unit, several ratios."""
    query, db, offs = _ragged_orb_case(91, 900, 256, 2300, 2048 + 77)
    q, t, o = dev(query, cuda_device), dev(db, cuda_device), dev(offs, cuda_device)
    for ratio in (0.8, 0.75, 1.0):
        s1, b1 = hpusm.knn2_hamming_segmented(q, t, o, ratio, want_best_idx=True, engine="popc")
        s2, b2 = hpusm.knn2_hamming_segmented(q, t, o, ratio, want_best_idx=True, engine="tc")
        s3 = hpusm.knn2_hamming_segmented(q, t, o, ratio)   # AUTO resolves to the tensor engine at this shape
        assert torch.equal(s1, s2) and torch.equal(b1, b2) and torch.equal(s1, s3)
        assert int(s1.sum()) > 0


def test_hamming_tensor_engine_shape_limits(hpusm, cuda_device):
    """Outside its shape limits the explicit tensor engine refuses; AUTO takes the __popc engine and answers."""
    query, db, offs, _ = hpusm.synth.orb_database(40, per_image=64, seed=2)
    q, t, o = dev(query, cuda_device), dev(db, cuda_device), dev(offs, cuda_device)
    with pytest.raises(RuntimeError, match="tensor-core engine"):
        hpusm.knn2_hamming_segmented(q, t, o, 0.8, engine="tc")
    rs, _ = knn_oracle.segmented_scores(query, db, offs, 0.8)
    np.testing.assert_array_equal(hpusm.knn2_hamming_segmented(q, t, o, 0.8).cpu().numpy(), rs)


def test_knn_merge_equals_unsplit(hpusm, cuda_device):
    q, t = hpusm.synth.orb_like(700, 9000, seed=3)
    full_i, full_d = hpusm.knn2_hamming(dev(q, cuda_device), dev(t, cuda_device))
    bounds = [0, 2500, 2501, 9000]
    pi, pd = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        i, d = hpusm.knn2_hamming(dev(q, cuda_device), dev(t[a:b], cuda_device))
        pi.append(torch.where(i >= 0, i + a, i))
        pd.append(d)
    mi, md = hpusm.engine.knn2_merge(torch.stack(pi), torch.stack(pd))
    assert torch.equal(mi, full_i) and torch.equal(md, full_d)


# ---- K1 L2 ----------------------------------------------------------------------------------------------
L2_MODES = ["auto", "simt"]


def _mode(hpusm, name):
    return {"auto": hpusm.engine.L2_AUTO, "simt": hpusm.engine.L2_SIMT, "tc_int": hpusm.engine.L2_TC_INT,
            "tc_split": hpusm.engine.L2_TC_SPLIT, "tc_i8": hpusm.engine.L2_TC_I8}[name]


@pytest.mark.parametrize("mode", L2_MODES)
@pytest.mark.parametrize("name", ["knn_l2_siftlike.npz", "knn_l2_float.npz", "knn_l2_single_train.npz"])
def test_l2_golden_cv2(hpusm, cuda_device, name, mode):
    g = load_golden(name)
    q, t = g["q"].astype(np.float32), g["t"].astype(np.float32)
    idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=_mode(hpusm, mode))
    if name == "knn_l2_siftlike.npz":  # integer-valued: bit exact, ties included
        np.testing.assert_array_equal(idx.cpu().numpy(), g["idx"])
        np.testing.assert_array_equal(dist.cpu().numpy(), g["dist"])
    else:
        assert_l2_parity(idx, dist, g["idx"], g["dist"], q, t)


@pytest.mark.parametrize("mode", L2_MODES)
def test_l2_scene_pair_sift(hpusm, cuda_device, mode):
    """C1: pisa1 x pisa2 SIFT, 4323 x 1085 x 128; knnMatch + filter_matches recorded from the reference."""
    g = load_golden("scene_pisa1_pisa2_sift.npz")
    q, t = g["desc_model"].astype(np.float32), g["desc_input"].astype(np.float32)
    idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=_mode(hpusm, mode))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["knn_idx"])
    np.testing.assert_array_equal(dist.cpu().numpy(), g["knn_dist"])
    keep, count = hpusm.ratio_filter(idx, dist, 0.8)
    assert int(count.item()) == 161
    p_q, p_t, sel, cnt = hpusm.gather_matches(dev(g["kp_model"], cuda_device), dev(g["kp_input"], cuda_device), idx, keep)
    np.testing.assert_array_equal(p_q[:161].cpu().numpy(), g["p_model"].reshape(-1, 2))
    np.testing.assert_array_equal(p_t[:161].cpu().numpy(), g["p_input"].reshape(-1, 2))


@pytest.mark.parametrize("mode", ["auto", "tc_i8", "tc_int", "tc_split", "simt"])
@pytest.mark.parametrize("tag", ["taj1_taj2", "eifel2_eifel3", "india1_india2"])
def test_l2_more_scene_pairs_sift(hpusm, cuda_device, tag, mode):
    """Three more real image pairs of the reference (785 x 1129, 2554 x 997, 1115 x 3080 SIFT descriptors): every
    engine returns cv2's knnMatch indices and distances bit for bit, and filter_matches' survivors."""
    g = load_golden("scene_%s_sift_knn.npz" % tag)
    q, t = g["desc_model"].astype(np.float32), g["desc_input"].astype(np.float32)
    idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=_mode(hpusm, mode))
    if mode == "auto":
        assert hpusm.engine.knn2_l2_last_mode() == hpusm.engine.L2_TC_I8
    np.testing.assert_array_equal(idx.cpu().numpy(), g["knn_idx"])
    np.testing.assert_array_equal(dist.cpu().numpy(), g["knn_dist"])
    keep, count = hpusm.ratio_filter(idx, dist, 0.8)
    n = len(g["p_model"].reshape(-1, 2))
    assert int(count.item()) == n
    p_q, p_t, sel, cnt = hpusm.gather_matches(dev(g["kp_model"], cuda_device), dev(g["kp_input"], cuda_device), idx, keep)
    np.testing.assert_array_equal(p_q[:n].cpu().numpy(), g["p_model"].reshape(-1, 2))
    np.testing.assert_array_equal(p_t[:n].cpu().numpy(), g["p_input"].reshape(-1, 2))


@pytest.mark.parametrize("engine", ["tc_int", "tc_i8"])
def test_l2_tensor_engine_units_and_splits(hpusm, cuda_device, engine):
    """Explicit tcgen05 engines (HPUSM_L2_TC_INT bf16, HPUSM_L2_TC_I8 u8): several 256-query units per CTA wave,
    database splits, ragged tails on both sides, duplicate database rows (ties -> lower index)."""
    mode = _mode(hpusm, engine)
    for nq, nt, seed in [(5000, 70001, 1), (257, 1085, 2), (40000, 4323, 3), (1, 129, 4), (200, 100000, 5)]:
        q, t = hpusm.synth.sift_like(nq, nt, seed=seed, planted=0.3)
        t[nt // 2] = t[3]
        t[nt - 1] = t[0]
        idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=mode)
        assert hpusm.engine.knn2_l2_last_mode() == mode
        ri, rd = knn_oracle.knn2_l2(q, t)
        np.testing.assert_array_equal(idx.cpu().numpy(), ri)
        np.testing.assert_array_equal(dist.cpu().numpy(), rd)


def test_l2_i8_engine_ties_and_parity(hpusm, cuda_device):
    """HPUSM_L2_TC_I8 orders by s = q.t - (||t||^2 >> 1) and settles the dropped parity bit only for columns that can
    matter.  Tiny value ranges make thousands of columns tie in s with differing parity, and make exact distance ties:
    the result must still be the oracle's (distance, then lower index), bit for bit."""
    for nq, nt, hi, seed in [(700, 9000, 2, 1), (300, 5000, 4, 2), (513, 2049, 3, 3), (64, 300, 256, 4)]:
        rng = np.random.default_rng(seed)
        q = rng.integers(0, hi, (nq, 128)).astype(np.float32)
        t = rng.integers(0, hi, (nt, 128)).astype(np.float32)
        if hi <= 4:
            q[:, 24:] = 0  # few active columns -> distances take a handful of values -> massive ties
            t[:, 24:] = 0
        idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=hpusm.engine.L2_TC_I8)
        ri, rd = knn_oracle.knn2_l2(q, t)
        np.testing.assert_array_equal(idx.cpu().numpy(), ri)
        np.testing.assert_array_equal(dist.cpu().numpy(), rd)


def test_l2_i8_fuzz_against_fp32_engine(hpusm, cuda_device):
    """Random shapes (ragged tails on both sides, fewer rows than one tile, several database splits), random value
    ranges (sparse rows, saturated rows, duplicates): the u8 engine and the exact FP32 CUDA-core engine agree bit for
    bit on indices and distances."""
    rng = np.random.default_rng(2024)
    E = hpusm.engine
    for case in range(24):
        nq = int(rng.integers(1, 3000))
        nt = int(rng.integers(2, 40000)) if case % 4 else int(rng.integers(2, 300))
        hi = int(rng.choice([2, 16, 100, 256]))
        q = rng.integers(0, hi, (nq, 128)).astype(np.float32)
        t = rng.integers(0, hi, (nt, 128)).astype(np.float32)
        if hi == 256 and case % 2:  # sparse rows next to dense ones (a dense uniform row has ||t||^2 ~ 2.8e6 < 4e6)
            t *= rng.integers(0, 3, (nt, 128)) == 0
            q *= rng.integers(0, 3, (nq, 128)) == 0
        dup = rng.integers(0, nt, max(1, nt // 50))
        t[dup] = t[rng.integers(0, nt, dup.size)]  # exact duplicates -> distance ties
        q[rng.integers(0, nq, max(1, nq // 20))] = t[rng.integers(0, nt, max(1, nq // 20))]  # zero-distance hits
        dq, dt = dev(q, cuda_device), dev(t, cuda_device)
        ia, da = hpusm.knn2_l2(dq, dt, mode=E.L2_TC_I8)
        ib, db_ = hpusm.knn2_l2(dq, dt, mode=E.L2_SIMT)
        assert torch.equal(ia, ib) and torch.equal(da, db_), (case, nq, nt, hi)


def test_l2_auto_engine_selection(hpusm, cuda_device):
    """HPUSM_L2_AUTO: u8 engine for integers in [0,255] with bounded norms, bf16 exact-int engine for other integers
    in [0,256], the fixed-point integer engine (Q16) for the rest (mixed signs included).  The explicit integer engines never synchronise: on
    data they cannot hold the device-side check leaves every idx at -1 and hpusm_async_status() reports it after the
    next synchronisation -- an error, not a wrong answer."""
    rng = np.random.default_rng(11)
    E = hpusm.engine
    cases = [
        (rng.integers(0, 256, (300, 128)), rng.integers(0, 100, (900, 128)), E.L2_TC_I8),
        (rng.integers(-9, 10, (300, 128)), rng.integers(0, 100, (900, 128)), E.L2_TC_Q16),      # negative query values
        (rng.integers(0, 257, (300, 128)), rng.integers(0, 100, (900, 128)), E.L2_TC_INT),      # a value of 256
        (rng.integers(0, 100, (300, 128)), rng.integers(200, 256, (900, 128)), E.L2_TC_INT),    # ||t||^2 > 4e6
        (rng.integers(0, 100, (300, 128)) + 0.5, rng.integers(0, 100, (900, 128)), E.L2_TC_Q16),
    ]
    for q, t, want in cases:
        q, t = q.astype(np.float32), t.astype(np.float32)
        idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=E.L2_AUTO)
        assert E.knn2_l2_last_mode() == want
        ri, rd = knn_oracle.knn2_l2(q, t)
        if want == E.L2_TC_Q16:
            assert_l2_parity(idx, dist, ri, rd, q, t)
        else:
            np.testing.assert_array_equal(idx.cpu().numpy(), ri)
            np.testing.assert_array_equal(dist.cpu().numpy(), rd)
        assert E.async_status() == 0
        for explicit in (E.L2_TC_I8, E.L2_TC_INT):
            ii, dd = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=explicit)
            torch.cuda.synchronize()
            holds = want == E.L2_TC_I8 or (want == E.L2_TC_INT and explicit == E.L2_TC_INT)
            if holds:
                assert E.async_status() == 0
                np.testing.assert_array_equal(ii.cpu().numpy(), ri)
            else:
                assert E.async_status() == -4 and E.async_status() == 0   # HPUSM_ERR_UNSUPPORTED, reported once
                assert bool((ii == -1).all())


@pytest.mark.parametrize("mode", L2_MODES)
@pytest.mark.parametrize("nq,nt,d,kind", [(1, 2, 128, "int"), (65, 63, 128, "int"), (300, 20000, 128, "int"),
                                          (2000, 3000, 128, "float"), (129, 4097, 64, "float"), (50, 700, 36, "float"),
                                          (7, 0, 128, "int"), (0, 7, 128, "int")])
def test_l2_edge_shapes(hpusm, cuda_device, nq, nt, d, kind, mode):
    rng = np.random.default_rng(nq + 3 * nt + d)
    if kind == "int":
        q, t = hpusm.synth.sift_like(nq, nt, seed=nq + nt, planted=0.2, d=d)
    else:
        q = rng.standard_normal((nq, d)).astype(np.float32)
        t = rng.standard_normal((nt, d)).astype(np.float32)
    idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=_mode(hpusm, mode))
    ri, rd = knn_oracle.knn2_l2(q, t)
    if kind == "int":
        np.testing.assert_array_equal(idx.cpu().numpy(), ri)
        np.testing.assert_array_equal(dist.cpu().numpy(), rd)
    else:
        assert_l2_parity(idx, dist, ri, rd, q, t)


def test_ransac_streamed_batch_equals_one_call(hpusm, cuda_device):
    """retrieval.StreamedHomographyBatch (host matches uploaded in runs of whole pairs, each run solved with its
    pair_index0) returns the one-call H, mask and inlier counts bit for bit -- the sample table is keyed on the pair's
    index in the whole batch, not in the run."""
    src, dst, offs = hpusm.synth.ransac_batch(23, 300, 0.4, seed=3)
    sizes = np.random.default_rng(1).integers(3, 301, 23)  # ragged pairs, one with fewer than 4 matches (no model)
    sizes[7] = 3
    keep = np.concatenate([np.arange(offs[p], offs[p] + sizes[p]) for p in range(23)])
    src, dst = np.ascontiguousarray(src[keep]), np.ascontiguousarray(dst[keep])
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    want = hpusm.engine.ransac_homography_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device), 5.0, 256,
                                                0.995, seed=7)
    sb = hpusm.retrieval.StreamedHomographyBatch(offs, cuda_device, 5.0, 256, 0.995, seed=7, pairs_per_chunk=4)
    for _ in range(2):
        H, mask, ninl = sb(torch.from_numpy(src).pin_memory(), torch.from_numpy(dst).pin_memory())
        assert torch.equal(torch.nan_to_num(H, nan=-7.0), torch.nan_to_num(want[0], nan=-7.0))
        assert torch.equal(mask, want[1]) and torch.equal(ninl, want[2])
    # default chunking: a first run of one wave of the one-CTA-per-pair kernel (3 CTAs per SM), then runs of two waves
    src2, dst2, offs2 = hpusm.synth.ransac_batch(1000, 64, 0.5, seed=4)
    want2 = hpusm.engine.ransac_homography_batch(dev(src2, cuda_device), dev(dst2, cuda_device), dev(offs2, cuda_device), 5.0, 64,
                                                 0.995, seed=8)
    sb2 = hpusm.retrieval.StreamedHomographyBatch(offs2, cuda_device, 5.0, 64, 0.995, seed=8)
    assert len(sb2.chunks) >= 2 and sb2.chunks[0][1] == 3 * torch.cuda.get_device_properties(cuda_device).multi_processor_count
    H2, mask2, ninl2 = sb2(torch.from_numpy(src2).pin_memory(), torch.from_numpy(dst2).pin_memory())
    assert torch.equal(torch.nan_to_num(H2, nan=-7.0), torch.nan_to_num(want2[0], nan=-7.0))
    assert torch.equal(mask2, want2[1]) and torch.equal(ninl2, want2[2])
    # and the slice entry point on its own: pairs 5.. of the batch, scored as pairs 5.. of the sample table
    p0 = 5
    r0 = int(offs[p0])
    local = torch.from_numpy(offs[p0:] - r0).to(cuda_device)
    c_all = hpusm.engine.ransac_score_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device), 5.0, 128, seed=9)
    c_tail = hpusm.engine.ransac_score_batch(dev(src[r0:], cuda_device), dev(dst[r0:], cuda_device), local, 5.0, 128, seed=9,
                                             pair_index0=p0)
    assert torch.equal(c_all[p0:], c_tail)


def test_hamming_streamed_image_scorer(hpusm, cuda_device):
    """retrieval.StreamedImageScorer (host database uploaded in runs of whole images under the compute) returns the
    per-image counts of one hpusm_knn2_hamming_segmented call: ragged images, a run size that does not divide them."""
    q, db, offs, _ = hpusm.synth.orb_database(37, per_image=300, seed=5, ragged=True)
    want = hpusm.engine.knn2_hamming_segmented(dev(q, cuda_device), dev(db, cuda_device), dev(offs, cuda_device), 0.8)
    sc = hpusm.retrieval.StreamedImageScorer(q.shape[0], q.shape[1], offs, cuda_device, images_per_chunk=5)
    for _ in range(2):
        got = sc(torch.from_numpy(q).pin_memory(), torch.from_numpy(db).pin_memory(), None, 0.8)
        assert torch.equal(got, want)


def test_l2_streamed_host_matcher(hpusm, cuda_device):
    """retrieval.StreamedL2Matcher (host arrays, chunked H2D under the compute, per-chunk top-2 merged) returns exactly
    what one knn2_l2 call on the whole arrays returns: ragged chunks, a 1-row tail, duplicates across chunk borders."""
    for nq, nt, qc, tf, tc in [(5000, 9001, 2048, 1000, 3000), (300, 4097, 512, 1, 1024), (1000, 700, 4096, 4096, 4096),
                               (700, 2049, 100, 2048, 1 << 20)]:
        q, t = hpusm.synth.sift_like(nq, nt, seed=nq + nt, planted=0.3)
        t[nt - 1] = t[0]
        t[nt // 2] = t[1]
        m = hpusm.retrieval.StreamedL2Matcher(nq, nt, 128, cuda_device, q_chunk=qc, t_first=tf, t_chunk=tc)
        for _ in range(2):  # buffers are reused from call to call
            idx, dist = m(torch.from_numpy(q).pin_memory(), torch.from_numpy(t).pin_memory())
            ri, rd = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device))
            assert torch.equal(idx, ri) and torch.equal(dist, rd)
    with pytest.raises(ValueError):
        m(dev(q, cuda_device), dev(t, cuda_device))
    # default chunking: growing whole-wave query chunks against the first database chunk, then ONE call per later chunk
    nq, nt = 3 * 256 * 148 + 77, 300001
    q, t = hpusm.synth.sift_like(nq, nt, seed=9, planted=0.2)
    m = hpusm.retrieval.StreamedL2Matcher(nq, nt, 128, cuda_device, t_chunk=100000)
    assert len(m.q_bounds) >= 3 and len(m.t_bounds) >= 3
    idx, dist = m(torch.from_numpy(q).pin_memory(), torch.from_numpy(t).pin_memory())
    ri, rd = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device))
    assert torch.equal(idx, ri) and torch.equal(dist, rd)


def test_l2_auto_equals_simt_at_scale(hpusm, cuda_device):
    """Size-independent property at a size no CPU oracle reaches quickly: the tensor-core engine and the exact
    CUDA-core brute force agree bit for bit on integer-valued descriptors (65536 x 262144)."""
    g = torch.Generator(device=cuda_device).manual_seed(1)
    t = torch.randint(0, 120, (262144, 128), generator=g, device=cuda_device).float()
    q = torch.randint(0, 120, (65536, 128), generator=g, device=cuda_device).float()
    q[:1000] = t[5000:6000] + torch.randint(0, 3, (1000, 128), generator=g, device=cuda_device).float()
    ia, da = hpusm.knn2_l2(q, t, mode=hpusm.engine.L2_AUTO)
    ib, db_ = hpusm.knn2_l2(q, t, mode=hpusm.engine.L2_SIMT)
    assert torch.equal(ia, ib) and torch.equal(da, db_)
    assert (ia[:1000, 0].cpu() == torch.arange(5000, 6000)).all()


# ---- ORB description (the step in front of K2) -----------------------------------------------------------------------------
def test_orb_compute_is_cv2_orb_compute(hpusm, cuda_device):
    """hpusm_orb_compute against the reference detector's own descriptors (cv2.ORB.compute, features.py:19-20) on three
    img/ fixtures: every bit of every descriptor on all 8 pyramid levels; the smoothed pyramid against the oracle; and
    the descriptors feed knn2_hamming without leaving the device."""
    from oracle import orb_oracle
    g = load_golden("orb_descriptors.npz")
    assert np.array_equal(np.load(os.path.join(ROOT, "tests", "golden", "orb_pattern.npy")).astype(np.int8),
                          np.loadtxt(os.path.join(ROOT, "humanpose-and-urbanscene-matching_b200", "csrc", "orb_pattern.inc"),
                                     delimiter=",", comments="//", usecols=range(4 * 4)).reshape(256, 4).astype(np.int8))
    descs = []
    for k, name in enumerate(g["names"]):
        img, pt, octave, angle, want = g["img_%d" % k], g["pt_%d" % k], g["octave_%d" % k], g["angle_%d" % k], g["desc_%d" % k]
        lxy, cs = hpusm.engine.orb_keypoint_params(pt, octave, angle)
        ocx, ocy, oa, ob = orb_oracle.keypoint_params(pt, octave, angle)
        assert np.array_equal(lxy[:, 1], ocx) and np.array_equal(lxy[:, 2], ocy) and np.array_equal(cs[:, 0], oa)
        desc = hpusm.engine.orb_compute(dev(img, cuda_device), dev(lxy, cuda_device), dev(cs, cuda_device))
        np.testing.assert_array_equal(desc.cpu().numpy(), want, err_msg=str(name))
        descs.append(desc)
    assert sum(len(d) for d in descs) > 10000
    idx, dist = hpusm.knn2_hamming(descs[2], descs[0])         # pisa2 vs pisa7, straight from the descriptor kernel
    ri, rd = knn_oracle.knn2_hamming(g["desc_2"], g["desc_0"])
    np.testing.assert_array_equal(idx.cpu().numpy(), ri)
    np.testing.assert_array_equal(dist.cpu().numpy(), rd)


def test_orb_detect_and_compute_is_cv2(hpusm, cuda_device):
    """hpusm_orb_detect_and_compute against cv2.ORB.detectAndCompute of the reference's detector (urban_scene.py:31-32) on
    three img/ fixtures: the same keypoints in the same order -- pt, octave, angle, response, size identical -- and their
    descriptors bit for bit; per-level counts; a capacity that is too small is grown."""
    g = load_golden("orb_descriptors.npz")
    for k, name in enumerate(g["names"]):
        img = g["img_%d" % k]
        out = hpusm.engine.orb_detect_and_compute(dev(img, cuda_device), capacity=1000 if k == 0 else None)
        assert not out["over_quota"]
        for key in ("pt", "octave", "angle", "response", "size", "desc"):
            np.testing.assert_array_equal(out[key].cpu().numpy(), g["%s_%d" % (key, k)], err_msg="%s %s" % (name, key))
        assert out["level_counts"] == [int((g["octave_%d" % k] == l).sum()) for l in range(8)]
        lxy, cs = hpusm.engine.orb_keypoint_params(g["pt_%d" % k], g["octave_%d" % k], g["angle_%d" % k])
        np.testing.assert_array_equal(out["kp_lxy"].cpu().numpy(), lxy)
        np.testing.assert_array_equal(out["kp_cs"].cpu().numpy(), cs)


def test_orb_detect_odd_shapes_against_oracle(hpusm, cuda_device):
    """Detection on synthetic images whose sizes are not multiples of any tile (incl. one too small for the upper pyramid
    levels to hold a keypoint) against the oracle's detect."""
    from oracle import orb_oracle
    rng = np.random.default_rng(12)
    for (h, w) in ((97, 131), (241, 333), (77, 400)):
        base = rng.integers(0, 256, (h // 4 + 2, w // 4 + 2), dtype=np.uint8)
        img = np.kron(base, np.ones((4, 4), np.uint8))[:h, :w].copy()       # blocky texture: plenty of corners
        img = np.clip(img.astype(np.int32) + rng.integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
        want = orb_oracle.detect(img)
        out = hpusm.engine.orb_detect_and_compute(dev(img, cuda_device))
        assert len(want["pt"]) > 0
        for key in ("pt", "octave", "angle", "response", "size"):
            np.testing.assert_array_equal(out[key].cpu().numpy(), want[key], err_msg="%dx%d %s" % (h, w, key))
        np.testing.assert_array_equal(out["desc"].cpu().numpy(), orb_oracle.compute(img, want["pt"], want["octave"], want["angle"]))


def test_orb_compute_odd_shapes(hpusm, cuda_device):
    """Images whose sizes are not multiples of the kernels' tiles, keypoints on every level incl. close to the border
    margin OpenCV keeps (31 px): the device descriptors equal the oracle's."""
    from oracle import orb_oracle
    rng = np.random.default_rng(8)
    for (h, w) in ((97, 131), (240, 333), (480, 641)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img = (0.5 * img + 0.5 * np.roll(img, 3, 1)).astype(np.uint8)
        pts, octs = [], []
        for level in range(8):
            lw, lh = orb_oracle.level_size(w, h, level)
            if lw <= 64 or lh <= 64:
                continue
            s = float(orb_oracle.level_scale(level))
            m = 40
            xs = rng.uniform(31, lw - 32, m) * s
            ys = rng.uniform(31, lh - 32, m) * s
            pts.append(np.stack([xs, ys], 1)); octs.append(np.full(m, level))
        pts, octs = np.concatenate(pts).astype(np.float32), np.concatenate(octs).astype(np.int32)
        angs = rng.uniform(0, 360, len(pts)).astype(np.float32)
        lxy, cs = hpusm.engine.orb_keypoint_params(pts, octs, angs)
        desc = hpusm.engine.orb_compute(dev(img, cuda_device), dev(lxy, cuda_device), dev(cs, cuda_device))
        np.testing.assert_array_equal(desc.cpu().numpy(), orb_oracle.compute(img, pts, octs, angs))


# ---- K3 RANSAC --------------------------------------------------------------------------------------------
SAMPLERS = [pytest.param(0, id="counter"), pytest.param(1, id="cv2")]


def assert_h_close(H, ref, tol=1e-6):
    """Homographies agree to `tol` of their largest entry (rows of NaN = no model on both sides)."""
    H, ref = np.asarray(H).reshape(-1, 9), np.asarray(ref).reshape(-1, 9)
    assert np.array_equal(np.isnan(H), np.isnan(ref))
    ok = ~np.isnan(ref).any(axis=1)
    scale = np.abs(ref[ok]).max(axis=1, keepdims=True)
    assert (np.abs(H[ok] - ref[ok]) / scale).max(initial=0.0) <= tol


@pytest.mark.parametrize("sampler", SAMPLERS)
def test_ransac_counts_identical_to_oracle(hpusm, cuda_device, sampler):
    src, dst, offs = hpusm.synth.ransac_batch(6, 700, inlier_ratio=0.4, seed=8)
    offs = np.array([0, 700, 1400, 1403, 1407, 3000, 4200], np.int64)  # ragged pairs incl. K < 4 and K == 4
    counts = hpusm.ransac_score_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device), 5.0, 512, seed=9,
                                      sampler=sampler)
    ref = ransac_oracle.score_batch(src, dst, offs, 5.0, 512, 9, sampler)
    np.testing.assert_array_equal(counts.cpu().numpy(), ref)
    assert (ref[0] >= 0).all() and (ref[2] == -1).all()  # every iteration of a real pair scores a model; K < 4: none


@pytest.mark.parametrize("sampler", SAMPLERS)
def test_ransac_large_pair_streams_in_chunks(hpusm, cuda_device, sampler):
    """A pair with more matches than fit in shared memory (staged chunk by chunk, samples read from global memory)."""
    src, dst, _ = hpusm.synth.ransac_pair(9001, inlier_ratio=0.4, seed=12)
    offs = np.array([0, 9001], np.int64)
    counts = hpusm.ransac_score_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device), 5.0, 300, seed=3,
                                      sampler=sampler)
    np.testing.assert_array_equal(counts.cpu().numpy(), ransac_oracle.score_batch(src, dst, offs, 5.0, 300, 3, sampler))
    H, mask, ninl, best = hpusm.ransac_homography_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device),
                                                        5.0, 300, 0.995, seed=3, sampler=sampler)
    rH, rmask, rninl, rbest = ransac_oracle.homography_batch(src, dst, offs, 5.0, 300, 0.995, 3, True, sampler)
    np.testing.assert_array_equal(best.cpu().numpy(), rbest)
    np.testing.assert_array_equal(mask.cpu().numpy(), rmask)
    assert_h_close(H.cpu().numpy(), rH)


@pytest.mark.parametrize("sampler", SAMPLERS)
@pytest.mark.parametrize("confidence", [0.995, 0.0])
@pytest.mark.parametrize("refine", [True, False])
def test_ransac_homography_identical_inlier_sets(hpusm, cuda_device, confidence, sampler, refine):
    g = load_golden("ransac_cv2.npz")
    src, dst, offs = g["src"], g["dst"], g["pair_offsets"]
    H, mask, ninl, best = hpusm.ransac_homography_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device),
                                                        5.0, 2000, confidence, seed=7, refine=refine, sampler=sampler)
    rH, rmask, rninl, rbest = ransac_oracle.homography_batch(src, dst, offs, 5.0, 2000, confidence, 7, refine, sampler)
    np.testing.assert_array_equal(best.cpu().numpy(), rbest)
    np.testing.assert_array_equal(mask.cpu().numpy(), rmask)
    np.testing.assert_array_equal(ninl.cpu().numpy(), rninl)
    assert_h_close(H.cpu().numpy(), rH, 1e-7 if not refine else 1e-6)
    if refine and confidence > 0:
        # against cv2 itself: the mask always; H everywhere under cv2's own sample sequence, and where the result does not
        # depend on the sequence (noise-free inliers) under the counter sampler
        np.testing.assert_array_equal(mask.cpu().numpy(), g["mask"])
        rows = range(len(rH)) if sampler == 1 else np.nonzero(g["noise"] == 0.0)[0]
        assert_h_close(H.cpu().numpy()[list(rows)], g["H"][list(rows)], 1e-4)


def _real_pairs():
    g = load_golden("ransac_real_pairs.npz")
    for k, name in enumerate(g["names"]):
        yield (str(name), g["p_model_%d" % k].reshape(-1, 2), g["p_input_%d" % k].reshape(-1, 2), g["mask_%d" % k].ravel(),
               g["mask2_%d" % k].ravel(), g["H_%d" % k], g["H2_%d" % k], g["valid_%d" % k])


def test_ransac_cv2_sampler_is_find_homography_on_real_pairs(hpusm, cuda_device):
    """The two cv2.findHomography calls of features.match_and_draw (features.py:66,:69) on 13 real image pairs, all 26
    problems in ONE batched call with OpenCV's sample sequence replayed: mask identical to cv2's and H within 1e-4
    (north_star's tolerance) wherever the reference accepts the homography (validate_homography); elsewhere (ill-posed
    fits whose 10 LM iterations do not converge) the inlier count of the RANSAC stage is pinned through the oracle."""
    srcs, dsts, offs, want = [], [], [0], []
    for name, pm, pi, mask, mask2, H, H2, valid in _real_pairs():
        for a, b, m, h, ok in ((pm, pi, mask, H, valid[0]), (pi, pm, mask2, H2, valid[1])):
            srcs.append(a); dsts.append(b); offs.append(offs[-1] + len(a)); want.append((name, m, h, bool(ok)))
    src, dst, offs = np.concatenate(srcs), np.concatenate(dsts), np.array(offs, np.int64)
    H, mask, ninl, best = hpusm.ransac_homography_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device),
                                                        5.0, 2000, 0.995, sampler="cv2")
    rH, rmask, rninl, rbest, rmin = ransac_oracle.homography_batch(src, dst, offs, 5.0, 2000, 0.995, 0, True, 1, 0, True)
    np.testing.assert_array_equal(best.cpu().numpy(), rbest)  # same winning iteration as the oracle on all 26
    H, mask = H.cpu().numpy(), mask.cpu().numpy()
    n_valid = 0
    for p, (name, m, h, ok) in enumerate(want):
        if ok:
            np.testing.assert_array_equal(mask[offs[p]:offs[p + 1]], m, err_msg=name)
            assert_h_close(H[p], h, 1e-4)
            assert np.abs(H[p] / h - 1).max() < 1e-5, name
            n_valid += 1
    assert n_valid >= 20


def test_ransac_cv2_sampler_is_find_homography_on_1000_noisy_pairs(hpusm, cuda_device):
    """cv2.findHomography(src, dst, cv2.RANSAC, 5.0) on 1000 seeded noisy synthetic pairs, one batched call: identical
    masks, H within 1e-4 on every pair."""
    from oracle.ransac_cases import ransac_synth_batch
    g = load_golden("ransac_synth_cv2.npz")
    src, dst, offs = ransac_synth_batch()
    want_mask = np.unpackbits(g["mask_bits"])[:offs[-1]]
    H, mask, ninl, best = hpusm.ransac_homography_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device),
                                                        5.0, 2000, 0.995, sampler="cv2")
    np.testing.assert_array_equal(mask.cpu().numpy(), want_mask)
    assert_h_close(H.cpu().numpy(), g["H"], 1e-4)
    np.testing.assert_array_equal(ninl.cpu().numpy(), np.add.reduceat(want_mask.astype(np.int64), offs[:-1]))


@pytest.mark.parametrize("sampler", SAMPLERS)
def test_ransac_scene_pair(hpusm, cuda_device, sampler):
    """C1's matches (161, ~11 % inliers): every iteration scores a model; same winner and mask as the oracle; under the
    cv2 sampler the mask is the one the reference's match_and_draw returned (18 inliers)."""
    g = load_golden("scene_pisa1_pisa2_sift.npz")
    src, dst = g["p_model"].reshape(-1, 2), g["p_input"].reshape(-1, 2)
    offs = np.array([0, len(src)], np.int64)
    counts = hpusm.ransac_score_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device), 5.0, 2000, seed=0,
                                      sampler=sampler)
    assert int((counts >= 4).sum()) == 2000
    H, mask, ninl, best = hpusm.ransac_homography_batch(dev(src, cuda_device), dev(dst, cuda_device), dev(offs, cuda_device),
                                                        5.0, 2000, 0.995, seed=0, sampler=sampler)
    rH, rmask, rninl, rbest = ransac_oracle.homography_batch(src, dst, offs, 5.0, 2000, 0.995, 0, True, sampler)
    np.testing.assert_array_equal(best.cpu().numpy(), rbest)
    if sampler == 1:
        assert int(ninl.item()) == int(g["mask"].sum()) == 18


@pytest.mark.parametrize("sampler", SAMPLERS)
def test_ransac_no_model(hpusm, cuda_device, sampler):
    src = np.zeros((10, 2), np.float32)  # all points coincide: no sample passes the subset check
    offs = np.array([0, 10], np.int64)
    H, mask, ninl, best = hpusm.ransac_homography_batch(dev(src, cuda_device), dev(src, cuda_device), dev(offs, cuda_device),
                                                        5.0, 64, 0.995, seed=1, sampler=sampler)
    assert torch.isnan(H).all() and int(ninl.item()) == 0 and int(best.item()) == -1 and int(mask.sum()) == 0


def test_validate_homography_batch_equals_reference(hpusm, cuda_device):
    """hpusm_validate_homography_batch against features.validate_homography (features.py:177-211) recorded on 1000
    matrices around every threshold (determinant sign incl. exactly singular blocks, row norms 0.1 / 4, perspective
    norm 0.002); NaN rows (no model) are invalid."""
    g = load_golden("validate_homography.npz")
    H = np.concatenate([g["H"], np.full((3, 3, 3), np.nan)])
    valid = hpusm.engine.validate_homography_batch(dev(H, cuda_device)).cpu().numpy().astype(bool)
    np.testing.assert_array_equal(valid[:-3], g["valid"])
    assert not valid[-3:].any()


def test_perspective_transform_cv2(hpusm, cuda_device):
    g = load_golden("perspective_transform.npz")
    out = hpusm.perspective_transform(dev(g["pts"], cuda_device), dev(g["H"], cuda_device))
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=1e-6, atol=1e-5)


# ---- K4 pose ------------------------------------------------------------------------------------------------
def test_affine_lsq_reference(hpusm, cuda_device):
    g = load_golden("pose_find_transformation.npz")
    for npts in (5, 7, 10):
        sel = np.nonzero(g["npts"] == npts)[0]
        out, A, has = hpusm.affine_lsq_batch(dev(g["model"][sel][:, :npts], cuda_device), dev(g["input"][sel][:, :npts], cuda_device))
        np.testing.assert_array_equal(has.cpu().numpy().astype(bool), g["has_A"][sel])
        np.testing.assert_allclose(out.cpu().numpy(), g["out"][sel][:, :npts], rtol=1e-4, atol=1e-6)
        hs = g["has_A"][sel]
        np.testing.assert_allclose(A.cpu().numpy()[hs], g["A"][sel][hs], rtol=1e-4, atol=1e-6)


def test_procrustes_reference(hpusm, cuda_device):
    g = load_golden("pose_procrustes.npz")
    n = len(g["d"])
    for o in range(6):
        sel = np.arange(o, n, 6)
        rf = int(g["reflection"][o])
        refl = "best" if rf == 0 else (rf == 1)
        d, Z, tf = hpusm.procrustes_batch(dev(g["X"][sel], cuda_device), dev(g["Y"][sel], cuda_device), bool(g["scaling"][o]), refl)
        np.testing.assert_allclose(d.cpu().numpy(), g["d"][sel], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(Z.cpu().numpy(), g["Z"][sel], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(tf.cpu().numpy(), g["tform"][sel], rtol=1e-4, atol=1e-6)


def _pose_groups(M, I):
    """Group the golden (model, input) pairs by model pose so each group is one `query vs db` launch."""
    keys = {}
    for k in range(len(M)):
        keys.setdefault(M[k].tobytes(), []).append(k)
    return list(keys.values())


def test_pose_match_single_reference(hpusm, cuda_device):
    """GPU vs the oracle on the float32-rounded poses (tight), and vs the reference's own float64 results (1e-4)."""
    g = load_golden("pose_match_single.npz")
    M32, I32 = g["model"].astype(np.float32), g["input"].astype(np.float32)
    n_tight = n_ref = n_flip = 0
    for rows in _pose_groups(M32, I32):
        q = M32[rows[0]]
        match, score, detail = hpusm.pose_match_batch(dev(q, cuda_device), dev(I32[rows], cuda_device), True, want_detail=True)
        match, score, detail = match.cpu().numpy(), score.cpu().numpy(), detail.cpu().numpy()
        for j, k in enumerate(rows):
            with np.errstate(all="ignore"):
                try:
                    r = pose_oracle.match_single(q.astype(np.float64), I32[k].astype(np.float64))
                except (ValueError, np.linalg.LinAlgError):
                    assert match[j] == 0
                    continue
            if not np.isfinite(r["error_score"]):
                continue
            assert bool(match[j]) == r["match"], k
            np.testing.assert_allclose(detail[j, 7], r["error_score"], rtol=1e-7, atol=1e-10)
            np.testing.assert_allclose(detail[j, 35:].reshape(22, 2), r["input_transformation"], rtol=1e-6, atol=1e-8)
            n_tight += 1
            if g["match"][k] >= 0 and np.isfinite(g["score"][k]):
                np.testing.assert_allclose(score[j], g["score"][k], rtol=1e-4, atol=1e-6)
                n_flip += int(bool(match[j]) != bool(g["match"][k]))  # float32 storage of the database can flip a borderline
                n_ref += 1
    assert n_tight > 600 and n_ref > 600 and n_flip <= 2


@pytest.mark.parametrize("qim", [True, False])
def test_pose_fast_path_equals_exact_kernel(hpusm, cuda_device, qim):
    """hpusm_pose_match_batch with a workspace (FP32 fast path + exact FP64 kernel for what it cannot certify) against the
    exact FP64 kernel on every pose: identical decisions on all 1 M poses (missing joints, noise levels 0-15 px), scores
    within 2e-6; and against the oracle on a slice.  Includes the reference's corner cases (no joints, zero extent,
    negative coordinates), which the fast path must hand over."""
    db = hpusm.synth.pose_database(hpusm.synth.CANONICAL_POSE, 1_000_000, seed=21)
    db[0] = 0
    db[1, :, 0] = 100.0
    db[2, 3] = [-5.0, 40.0]
    db[3, 2:] = 0
    q = hpusm.synth.CANONICAL_POSE.astype(np.float32)
    dq, ddb = dev(q, cuda_device), dev(db, cuda_device)
    m_fast, s_fast = hpusm.pose_match_batch(dq, ddb, qim)
    m_exact, s_exact = hpusm.pose_match_batch(dq, ddb, qim, exact=True)
    assert torch.equal(m_fast, m_exact)
    ok = torch.isfinite(s_exact)
    assert torch.equal(torch.isfinite(s_fast), ok)
    assert float((s_fast[ok] - s_exact[ok]).abs().max()) < 2e-6
    k = 3000
    rm, rs = pose_oracle.match_single_batch(q.astype(np.float64), db[:k].astype(np.float64), qim)
    np.testing.assert_array_equal(m_fast[:k].cpu().numpy().astype(bool), rm)
    np.testing.assert_allclose(s_fast[:k].cpu().numpy(), rs, rtol=1e-5, atol=2e-6, equal_nan=True)
    # a query with a negative coordinate or an undetected joint: same contract
    for qq in (np.where(np.arange(18)[:, None] == 6, 0.0, q).astype(np.float32), (q - np.float32([300, 0])).astype(np.float32)):
        a = hpusm.pose_match_batch(dev(qq, cuda_device), ddb[:200000], qim)
        b = hpusm.pose_match_batch(dev(qq, cuda_device), ddb[:200000], qim, exact=True)
        assert torch.equal(a[0], b[0])
        fin = torch.isfinite(b[1])
        assert float((a[1][fin] - b[1][fin]).abs().max()) < 2e-6


def test_pose_topk(hpusm, cuda_device):
    rng = np.random.default_rng(4)
    n = 200000
    match = (rng.random(n) < 0.3).astype(np.uint8)
    score = rng.random(n).astype(np.float32)
    score[rng.integers(0, n, 50)] = np.nan
    score[1234] = score[77] = 0.0  # tie at the top: lower index first
    match[1234] = match[77] = 1
    ts, ti, cnt = hpusm.pose_topk(dev(match, cuda_device), dev(score, cuda_device), k=16, index_offset=1000)
    ok = (match == 1) & ~np.isnan(score)
    order = np.lexsort((np.arange(n), np.where(ok, score, np.inf)))[:16]
    np.testing.assert_array_equal(ti.cpu().numpy(), order + 1000)
    np.testing.assert_array_equal(ts.cpu().numpy(), score[order])
    assert int(cnt.item()) == int(match.sum())


@pytest.mark.parametrize("case", ["all_equal", "few", "none", "unaligned", "one_tile_hot", "tiny"])
def test_pose_topk_threshold_selection_edges(hpusm, cuda_device, case):
    """The two-pass selection (tile minima -> k-th smallest as threshold -> candidate list): every score equal (the list
    holds whole tiles), fewer matches than k, no match at all, views that are not 16-byte aligned, all of the best entries
    inside ONE tile, and a database smaller than one tile."""
    rng = np.random.default_rng(17)
    n, k, off = 300_000, 16, 0
    match = (rng.random(n) < 0.4).astype(np.uint8)
    score = rng.random(n).astype(np.float32)
    if case == "all_equal":
        score[:] = 0.25
    elif case == "few":
        match[:] = 0
        match[[5, 99_999, 250_000]] = 1
    elif case == "none":
        match[:] = 0
    elif case == "one_tile_hot":
        match[8192:8192 + 4096] = 1
        score[8192:8192 + 4096] = -1.0 - rng.random(4096).astype(np.float32)
    elif case == "tiny":
        n = 37
        match, score = match[:n].copy(), score[:n].copy()
        match[:] = 1
    m, s = dev(match, cuda_device), dev(score, cuda_device)
    if case == "unaligned":
        m, s, match, score, off = m[3:], s[3:], match[3:], score[3:], 3
    ts, ti, cnt = hpusm.pose_topk(m, s, k=k, index_offset=off)
    nn = len(match)
    ok = match == 1
    order = np.lexsort((np.arange(nn), np.where(ok, score, np.inf)))[:k]
    order = order[ok[order]]
    want_i = np.full(k, -1, np.int64)
    want_s = np.full(k, np.inf, np.float32)
    want_i[:len(order)] = order + off
    want_s[:len(order)] = score[order]
    np.testing.assert_array_equal(ti.cpu().numpy(), want_i)
    np.testing.assert_array_equal(ts.cpu().numpy(), want_s)
    assert int(cnt.item()) == int(match.sum())


def test_pose_topk_large_and_merge(hpusm, cuda_device):
    """3 M entries with heavy ties (many candidates at the threshold) and the candidate-list merge used
    after the all-gather: both equal the lexicographic (score, index) reference, padding and ties included."""
    rng = np.random.default_rng(9)
    n = 3_000_000
    match = (rng.random(n) < 0.5).astype(np.uint8)
    score = np.round(rng.random(n) * 1000).astype(np.float32)  # heavy ties
    ts, ti, cnt = hpusm.pose_topk(dev(match, cuda_device), dev(score, cuda_device), k=32, index_offset=0)
    order = np.lexsort((np.arange(n), np.where(match == 1, score, np.inf)))[:32]
    np.testing.assert_array_equal(ti.cpu().numpy(), order)
    np.testing.assert_array_equal(ts.cpu().numpy(), score[order])
    assert int(cnt.item()) == int(match.sum())
    # merge of 8 per-rank lists of 16, some padded
    cs = np.round(rng.random((8, 16)) * 20).astype(np.float32)
    ci = rng.permutation(10_000_000)[:128].reshape(8, 16).astype(np.int64)
    cs[3, 9:], ci[3, 9:] = np.inf, -1
    cs[6, :], ci[6, :] = np.inf, -1
    ms, mi = hpusm.engine.pose_topk_merge(dev(cs, cuda_device), dev(ci, cuda_device), k=16)
    flat_s, flat_i = cs.ravel(), ci.ravel()
    valid = flat_i >= 0
    ref = np.lexsort((flat_i[valid], flat_s[valid]))[:16]
    np.testing.assert_array_equal(mi.cpu().numpy(), flat_i[valid][ref])
    np.testing.assert_array_equal(ms.cpu().numpy(), flat_s[valid][ref])


# ---- fused post-match stage ---------------------------------------------------------------------
def _scene_database(hpusm, nimg=14, per=400, nq=500, seed=21):
    """A query image (descriptors + keypoints) and a database in which some images show the query scene under a
    homography (descriptors perturbed by a few bit flips, keypoints moved by H + noise)."""
    rng = np.random.default_rng(seed)
    q_desc = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    q_kp = np.stack([rng.uniform(0, 640, nq), rng.uniform(0, 480, nq)], 1).astype(np.float32)
    sizes = rng.integers(per // 2, per + 1, nimg)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    db_desc = rng.integers(0, 256, (int(offs[-1]), 32), dtype=np.uint8)
    db_kp = np.stack([rng.uniform(0, 640, offs[-1]), rng.uniform(0, 480, offs[-1])], 1).astype(np.float32)
    planted = [2, 5, 11]
    for n_shared, im in zip((120, 40, 10), planted):  # the last one stays below MIN_MATCH_COUNT
        H = hpusm.synth.random_homography(rng)
        rows = rng.choice(int(sizes[im]), n_shared, replace=False) + offs[im]
        src = rng.choice(nq, n_shared, replace=False)
        bits = np.unpackbits(q_desc[src], axis=1)
        flip = rng.random(bits.shape) < 0.03
        db_desc[rows] = np.packbits(bits ^ flip.astype(np.uint8), axis=1)
        p = np.c_[q_kp[src], np.ones(n_shared)] @ H.T
        db_kp[rows] = (p[:, :2] / p[:, 2:] + rng.normal(0, 0.5, (n_shared, 2))).astype(np.float32)
    return q_desc, q_kp, db_desc, db_kp, offs, planted


def test_scene_retrieval_fused_stage(hpusm, cuda_device):
    q_desc, q_kp, db_desc, db_kp, offs, planted = _scene_database(hpusm)
    out = hpusm.retrieval.scene_retrieval(dev(q_desc, cuda_device), dev(q_kp, cuda_device), dev(db_desc, cuda_device),
                                          dev(db_kp, cuda_device), dev(offs, cuda_device), nhyp=500, seed=5)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    # oracle: the reference's per-image loop (match_and_draw + validate_homography) with the shared sample seed
    nimg = len(offs) - 1
    srcs, dsts, poffs = [], [], [0]
    for s in range(nimg):
        lo, hi = offs[s], offs[s + 1]
        i2, d2 = knn_oracle.knn2_hamming(q_desc, db_desc[lo:hi])
        p1, p2, sel = knn_oracle.filter_matches_points(q_kp, db_kp[lo:hi], i2, d2, 0.8)
        assert got["n_matches"][s] == len(sel)
        if len(sel) >= 16:
            srcs.append(p1.reshape(-1, 2)); dsts.append(p2.reshape(-1, 2))
        poffs.append(poffs[-1] + (len(sel) if len(sel) >= 16 else 0))
    src, dst = np.concatenate(srcs), np.concatenate(dsts)
    np.testing.assert_array_equal(got["pair_offsets"], np.array(poffs))
    np.testing.assert_array_equal(got["src"], src)
    np.testing.assert_array_equal(got["dst"], dst)
    rH, rmask, rninl, _ = ransac_oracle.homography_batch(src, dst, poffs, 5.0, 500, 0.995, 5)
    rH2, _, _, _ = ransac_oracle.homography_batch(dst, src, poffs, 5.0, 500, 0.995, 5)
    np.testing.assert_array_equal(got["mask"], rmask)
    np.testing.assert_array_equal(got["n_inliers"], rninl)
    assert_h_close(got["H"], rH)
    assert_h_close(got["H2"], rH2)
    # the two well-supported planted images come out as valid homographies, the others (incl. the one below
    # MIN_MATCH_COUNT) do not
    assert got["valid"][planted[0]] == 1 and got["valid"][planted[1]] == 1
    assert got["n_inliers"][planted[0]] >= 100 and got["n_inliers"][planted[2]] == 0
    assert np.isnan(got["H"][planted[2]]).all()


@pytest.mark.parametrize("sampler", ["counter", "cv2"])
def test_imagedb_streamed_retrieval_equals_resident(hpusm, cuda_device, tmp_path, sampler):
    """Packed image database (descriptors + keypoints) streamed from disk in runs of whole images through
    scene_retrieval == one resident scene_retrieval call over the whole database, bit for bit."""
    q_desc, q_kp, db_desc, db_kp, offs, planted = _scene_database(hpusm)
    nimg = len(offs) - 1
    path = str(tmp_path / "images.hpusm")
    descs = [db_desc[offs[i]:offs[i + 1]] for i in range(nimg)]
    kps = [db_kp[offs[i]:offs[i + 1]] for i in range(nimg)]
    assert hpusm.imagedb.pack_arrays(descs, kps, path) == (nimg, int(offs[-1]))
    idb = hpusm.imagedb.ImageDB(path)
    got = idb.retrieve(q_desc, q_kp, device=cuda_device, images_per_chunk=3, nhyp=500, seed=5, sampler=sampler)
    want = hpusm.retrieval.scene_retrieval(dev(q_desc, cuda_device), dev(q_kp, cuda_device), dev(db_desc, cuda_device),
                                           dev(db_kp, cuda_device), dev(offs, cuda_device), nhyp=500, seed=5, sampler=sampler)
    for key in ("n_matches", "n_inliers", "valid"):
        np.testing.assert_array_equal(got[key], want[key].cpu().numpy(), err_msg=key)
    for key in ("H", "H2"):
        np.testing.assert_array_equal(np.nan_to_num(got[key], nan=-7.0), np.nan_to_num(want[key].cpu().numpy(), nan=-7.0), err_msg=key)
    assert got["valid"][planted[0]] == 1 and got["valid"][planted[1]] == 1


def test_posedb_streaming_search(hpusm, cuda_device, tmp_path):
    """Packed pose database streamed from disk in chunks (copy stream + two pinned buffers) == one resident pass."""
    db = hpusm.synth.pose_database(hpusm.synth.CANONICAL_POSE, 50000, seed=9)
    path = str(tmp_path / "poses.hpusm")
    assert hpusm.posedb.pack_array(db, path) == len(db)
    pdb = hpusm.posedb.PoseDB(path)
    q = hpusm.synth.CANONICAL_POSE.astype(np.float32)
    s, i, total = pdb.search(q, k=16, device=cuda_device, chunk=7001)  # 8 ragged chunks
    match, score = hpusm.pose_match_batch(dev(q, cuda_device), dev(db, cuda_device))
    ts, ti, cnt = hpusm.pose_topk(match, score, k=16)
    np.testing.assert_array_equal(i, ti.cpu().numpy())
    np.testing.assert_array_equal(s, ts.cpu().numpy())
    assert total == int(cnt.item()) == int(match.sum().item())


# ---- K1 split-operand tensor engine (general float descriptors) ---------------------------------------------------
def test_l2_split_engine_general_floats(hpusm, cuda_device):
    """HPUSM_L2_TC_SPLIT (and AUTO, which takes the fixed-point engine) on non-integer descriptors: indices equal the float64 brute force (ties within the
    stated 1e-5 relative tolerance aside), distances to 2e-6; units, database splits and ragged tails included."""
    rng = np.random.default_rng(77)
    for nq, nt in [(300, 900), (1, 5), (129, 3), (4000, 70001), (20000, 1085)]:
        q = rng.standard_normal((nq, 128)).astype(np.float32)
        t = rng.standard_normal((nt, 128)).astype(np.float32)
        k = min(nq, nt) // 3
        if k:
            t[:k] = q[:k] + 0.05 * rng.standard_normal((k, 128)).astype(np.float32)  # planted near neighbours
        for mode, resolved in ((hpusm.engine.L2_TC_SPLIT, hpusm.engine.L2_TC_SPLIT), (hpusm.engine.L2_AUTO, hpusm.engine.L2_TC_Q16)):
            idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=mode)
            assert hpusm.engine.knn2_l2_last_mode() == resolved
            ri, rd = knn_oracle.knn2_l2(q, t)
            assert_l2_parity(idx, dist, ri, rd, q, t)


def test_l2_split_engine_sift_like_unnormalised(hpusm, cuda_device):
    """Real-valued SIFT-like rows (norm ~ 512, non-integer): large norms stress the error bound of the proof."""
    q, t = hpusm.synth.sift_like(3000, 20000, seed=5, planted=0.3)
    q = (q * np.float32(1.0137)).astype(np.float32)
    t = (t * np.float32(1.0137)).astype(np.float32)
    idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=hpusm.engine.L2_TC_SPLIT)
    assert hpusm.engine.knn2_l2_last_mode() == hpusm.engine.L2_TC_SPLIT
    ri, rd = knn_oracle.knn2_l2(q, t)
    assert_l2_parity(idx, dist, ri, rd, q, t)


def test_l2_split_engine_near_ties_take_the_exact_fallback(hpusm, cuda_device):
    """Database rows packed so tightly around the queries that the top-4 lists cannot be proven complete: those
    rows must be redone by the exact FP32 engine, and the answer must still be the exact one."""
    rng = np.random.default_rng(78)
    nq, nt = 257, 6000
    q = rng.standard_normal((nq, 128)).astype(np.float32)
    t = rng.standard_normal((nt, 128)).astype(np.float32)
    for r in range(40):  # 20 database rows within 1e-4 of query r
        t[100 * r:100 * r + 20] = q[r] + 1e-4 * rng.standard_normal((20, 128)).astype(np.float32)
    idx, dist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=hpusm.engine.L2_TC_SPLIT)
    assert hpusm.engine.knn2_l2_last_fallback_rows() >= 30
    sidx, sdist = hpusm.knn2_l2(dev(q, cuda_device), dev(t, cuda_device), mode=hpusm.engine.L2_SIMT)
    assert torch.equal(idx, sidx) and torch.equal(dist, sdist)  # both end in the same exact FP32 re-rank


# ---- K1q fixed-point integer tensor engine (general float descriptors) ---------------------------------------------
def _q16_equals_exact(hpusm, cuda_device, q, t):
    """Both engines end in the same exact FP32 re-rank, and the proof is strict: the answers must be IDENTICAL."""
    dq, dt = dev(q, cuda_device), dev(t, cuda_device)
    idx, dist = hpusm.knn2_l2(dq, dt, mode=hpusm.engine.L2_TC_Q16)
    assert hpusm.engine.knn2_l2_last_mode() == hpusm.engine.L2_TC_Q16
    fb = hpusm.engine.knn2_l2_last_fallback_rows()
    sidx, sdist = hpusm.knn2_l2(dq, dt, mode=hpusm.engine.L2_SIMT)
    assert torch.equal(idx, sidx) and torch.equal(dist, sdist)
    return idx, dist, fb


def test_l2_q16_engine_general_floats(hpusm, cuda_device):
    """HPUSM_L2_TC_Q16 on non-integer descriptors: identical to the exact FP32 engine and equal to the float64 brute
    force (ties within the stated 1e-5 relative tolerance aside); units, database splits, ragged tails on both sides,
    fewer rows than one tile."""
    rng = np.random.default_rng(177)
    for nq, nt in [(300, 900), (1, 5), (129, 3), (257, 64), (256, 65), (4000, 70001), (20000, 1085)]:
        q = rng.standard_normal((nq, 128)).astype(np.float32)
        t = rng.standard_normal((nt, 128)).astype(np.float32)
        k = min(nq, nt) // 3
        if k:
            t[:k] = q[:k] + 0.05 * rng.standard_normal((k, 128)).astype(np.float32)  # planted near neighbours
        idx, dist, fb = _q16_equals_exact(hpusm, cuda_device, q, t)
        assert fb <= max(2, nq // 50), (nq, nt, fb)   # the proof covers (nearly) every row: the tensor path did the work
        ri, rd = knn_oracle.knn2_l2(q, t)
        assert_l2_parity(idx, dist, ri, rd, q, t)


def test_l2_q16_engine_sift_like_and_value_ranges(hpusm, cuda_device):
    """Real-valued SIFT-like rows (norm ~ 512: the quantisation bound is at its loosest relative to the neighbour gaps),
    tiny and huge magnitudes (the scale is measured on the data), one dominant component, flat rows, all-zero input."""
    rng = np.random.default_rng(178)
    q, t = hpusm.synth.sift_like(3000, 20000, seed=5, planted=0.3)
    q = (q * np.float32(1.0137)).astype(np.float32)
    t = (t * np.float32(1.0137)).astype(np.float32)
    idx, dist, fb = _q16_equals_exact(hpusm, cuda_device, q, t)
    assert fb < 300
    ri, rd = knn_oracle.knn2_l2(q, t)
    assert_l2_parity(idx, dist, ri, rd, q, t)
    for scale in (1e-12, 1e12):
        _q16_equals_exact(hpusm, cuda_device, (rng.standard_normal((500, 128)) * scale).astype(np.float32),
                          (rng.standard_normal((3000, 128)) * scale).astype(np.float32))
    t = rng.standard_normal((5000, 128)).astype(np.float32)
    t[:, 0] += 1000.0
    _q16_equals_exact(hpusm, cuda_device, rng.standard_normal((300, 128)).astype(np.float32), t)
    _q16_equals_exact(hpusm, cuda_device, np.zeros((130, 128), np.float32), np.zeros((700, 128), np.float32))
    flat = np.full((700, 128), 3.25, np.float32) + rng.integers(0, 2, (700, 128)).astype(np.float32)
    _q16_equals_exact(hpusm, cuda_device, np.full((130, 128), 3.25, np.float32), flat)


def test_l2_q16_near_ties_and_bad_values_take_the_exact_fallback(hpusm, cuda_device):
    """Rows whose lists cannot be proven complete (20 database rows within 1e-4 of a query) are redone by the exact engine;
    a non-finite value sends EVERY row there -- the answer is the exact engine's either way."""
    rng = np.random.default_rng(179)
    nq, nt = 257, 6000
    q = rng.standard_normal((nq, 128)).astype(np.float32)
    t = rng.standard_normal((nt, 128)).astype(np.float32)
    for r in range(40):
        t[100 * r:100 * r + 20] = q[r] + 1e-4 * rng.standard_normal((20, 128)).astype(np.float32)
    _, _, fb = _q16_equals_exact(hpusm, cuda_device, q, t)
    assert fb >= 30
    tiny = (q * np.float32(1e-24)).astype(np.float32), (t * np.float32(1e-24)).astype(np.float32)
    _, _, fb = _q16_equals_exact(hpusm, cuda_device, *tiny)   # squares underflow FP32: no bound can be built, all rows exact
    assert fb == nq
    t[17, 5] = np.inf
    dq, dt = dev(q, cuda_device), dev(t, cuda_device)
    hpusm.knn2_l2(dq, dt, mode=hpusm.engine.L2_TC_Q16)
    assert hpusm.engine.knn2_l2_last_fallback_rows() == nq


def test_l2_q16_cascade_through_the_split_engine(hpusm, cuda_device):
    """Clustered data: every query has 7 database rows at squared distances 0.05 (1 + 0.25 k): the 6th is within the
    fixed-point bound (~0.1) of the 2nd, so a 6-deep list cannot be proven complete, while the 4th is farther from the 2nd
    than the bf16 split engine's bound (~0.01).  With >= 1024 unproven rows the fixed-point engine hands them to the split
    engine (only ITS unproven rows reach the CUDA-core engine): the answer is still the exact engine's, and few rows needed it."""
    rng = np.random.default_rng(181)
    nq, nt = 3000, 30000
    q = rng.standard_normal((nq, 128)).astype(np.float32)
    t = rng.standard_normal((nt, 128)).astype(np.float32)
    for r in range(nq):
        u = rng.standard_normal((7, 128))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        t[8 * r:8 * r + 7] = (q[r] + np.sqrt(0.05 * (1 + 0.25 * np.arange(7)))[:, None] * u).astype(np.float32)
    os.environ["HPUSM_Q16_CASCADE"] = "0"
    try:
        _, _, direct = _q16_equals_exact(hpusm, cuda_device, q, t)
    finally:
        del os.environ["HPUSM_Q16_CASCADE"]
    assert direct >= 1024, direct     # the fixed-point proof fails on (nearly) every row of this data ...
    _, _, fb = _q16_equals_exact(hpusm, cuda_device, q, t)
    assert fb < direct // 4, (fb, direct)   # ... and the split engine in between settles most of them


def test_l2_auto_pads_narrow_descriptors_onto_the_tensor_path(hpusm, cuda_device):
    """SURF rows are 64 floats (reference features.py:16).  On a problem of >= 2^24 pairs HPUSM_L2_AUTO zero-pads rows
    narrower than 128 inside its workspace and runs the tensor path: the answer is the d-wide exact engine's bit for
    bit (zeros change no distance and add fma(0, 0, s) = s to the re-rank); smaller problems stay on the FP32 engine."""
    rng = np.random.default_rng(180)
    E = hpusm.engine
    for d, nq, nt in [(64, 5000, 4000), (36, 4100, 4100), (124, 3000, 6000)]:
        q = rng.standard_normal((nq, d)).astype(np.float32)
        t = rng.standard_normal((nt, d)).astype(np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)     # SURF descriptors are unit vectors
        t /= np.linalg.norm(t, axis=1, keepdims=True)
        t[:1000] = q[:1000] + 0.02 * rng.standard_normal((1000, d)).astype(np.float32)
        dq, dt = dev(q, cuda_device), dev(t, cuda_device)
        idx, dist = hpusm.knn2_l2(dq, dt, mode=E.L2_AUTO)
        assert E.knn2_l2_last_mode() == E.L2_TC_Q16
        sidx, sdist = hpusm.knn2_l2(dq, dt, mode=E.L2_SIMT)
        assert torch.equal(idx, sidx) and torch.equal(dist, sdist)
        ri, rd = knn_oracle.knn2_l2(q, t)
        assert_l2_parity(idx, dist, ri, rd, q, t)
    small = dev(rng.standard_normal((300, 64)).astype(np.float32), cuda_device)
    hpusm.knn2_l2(small, small, mode=E.L2_AUTO)
    assert E.knn2_l2_last_mode() == E.L2_SIMT
    ints = dev(rng.integers(0, 200, (5000, 64)).astype(np.float32), cuda_device)   # narrow AND integer: the u8 engine
    idx, dist = hpusm.knn2_l2(ints, ints[:4000].contiguous(), mode=E.L2_AUTO)
    assert E.knn2_l2_last_mode() == E.L2_TC_I8
    sidx, sdist = hpusm.knn2_l2(ints, ints[:4000].contiguous(), mode=E.L2_SIMT)
    assert torch.equal(idx, sidx) and torch.equal(dist, sdist)


def test_scene_score_batch_reference(hpusm, cuda_device):
    """Batched scene score vs the reference's perspective_correction + affine_multi (recorded, picks pinned), and the
    seeded picks vs the oracle."""
    g = load_golden("scene_score.npz")
    offs = g["pair_offsets"]
    n = len(g["score"])
    mask = np.ones(offs[-1], np.uint8)
    args = (dev(g["src"], cuda_device), dev(g["dst"], cuda_device), dev(offs, cuda_device), dev(mask, cuda_device),
            dev(g["H2"], cuda_device), dev(g["model_pose"], cuda_device), dev(g["input_pose"], cuda_device), 640.0, 480.0)
    score, used = hpusm.engine.scene_score_batch(*args, picks=dev(g["picks"], cuda_device))
    np.testing.assert_allclose(score.cpu().numpy(), g["score"], rtol=1e-6)
    np.testing.assert_array_equal(used.cpu().numpy(), g["picks"])
    # seeded picks: distinct ordinals; the score equals the oracle evaluated on those picks; masks are honoured
    mask2 = mask.copy()
    mask2[offs[0]:offs[0] + 40] = 0
    args2 = args[:3] + (dev(mask2, cuda_device),) + args[4:]
    score, used = hpusm.engine.scene_score_batch(*args2, seed=11)
    used = used.cpu().numpy()
    for p in range(n):
        lo, hi = offs[p], offs[p + 1]
        keep = mask2[lo:hi].astype(bool)
        assert len(set(used[p].tolist())) == 5 and used[p].min() >= 0 and used[p].max() < keep.sum()
        ref = pose_oracle.scene_score(g["src"][lo:hi][keep], g["dst"][lo:hi][keep], g["H2"][p], g["model_pose"][p],
                                      g["input_pose"][p], 640.0, 480.0, used[p])
        np.testing.assert_allclose(score[p].item(), ref, rtol=1e-6)
    # sentinels: no homography -> inf, fewer than 5 inliers -> 1
    H2 = g["H2"].copy(); H2[1] = np.nan
    mask3 = mask.copy(); mask3[offs[2]:offs[3]] = 0; mask3[offs[2]:offs[2] + 4] = 1
    score, _ = hpusm.engine.scene_score_batch(args[0], args[1], args[2], dev(mask3, cuda_device), dev(H2, cuda_device), args[5],
                                              args[6], 640.0, 480.0)
    assert np.isinf(score[1].item()) and score[2].item() == 1.0


def test_scene_retrieval_with_scores(hpusm, cuda_device):
    """The whole scene stage for one query against an image database, scores included: planted views score low."""
    q_desc, q_kp, db_desc, db_kp, offs, planted = _scene_database(hpusm)
    nimg = len(offs) - 1
    rng = np.random.default_rng(3)
    qpose = hpusm.synth.CANONICAL_POSE * 0.8 + 30.0
    db_poses = rng.uniform(50, 400, (nimg, 18, 2))
    out = hpusm.retrieval.scene_retrieval(dev(q_desc, cuda_device), dev(q_kp, cuda_device), dev(db_desc, cuda_device),
                                          dev(db_kp, cuda_device), dev(offs, cuda_device), nhyp=500, seed=5,
                                          query_pose=dev(qpose, cuda_device), db_poses=dev(db_poses, cuda_device),
                                          query_size=(640.0, 480.0))
    score, valid = out["score"].cpu().numpy(), out["valid"].cpu().numpy().astype(bool)
    assert np.isinf(score[~valid]).all() and np.isfinite(score[valid]).all() and valid[planted[0]] and valid[planted[1]]
    # oracle on the same picks
    H2, picks = out["H2"].cpu().numpy(), out["picks"].cpu().numpy()
    po, src, dst, mask = (out[k].cpu().numpy() for k in ("pair_offsets", "src", "dst", "mask"))
    for p in (planted[0], planted[1]):
        keep = mask[po[p]:po[p + 1]].astype(bool)
        ref = pose_oracle.scene_score(src[po[p]:po[p + 1]][keep], dst[po[p]:po[p + 1]][keep], H2[p], qpose, db_poses[p],
                                      640.0, 480.0, picks[p])
        np.testing.assert_allclose(score[p], ref, rtol=1e-6)
