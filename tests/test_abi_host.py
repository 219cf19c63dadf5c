"""CPU-side checks: the C-ABI library loads and exports every symbol include/hpusm.h declares, argument
validation and error reporting work without a GPU, and the host-side helpers behave."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hpusm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hpusm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(hpusm):
    lib = hpusm._lib.lib
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libhpusm.so does not export %s" % n
    assert set(names) == set(hpusm._lib.SIGNATURES), "ctypes table and header disagree"


def test_version_and_error_string(hpusm):
    lib = hpusm._lib.lib
    assert lib.hpusm_version() == 202
    rc = lib.hpusm_knn2_hamming(None, 4, None, 4, 48, None, None, None, 0, None)  # unsupported width, no GPU needed
    assert rc == hpusm._lib.ERR_UNSUPPORTED
    assert "48" in hpusm._lib.last_error()
    rc = lib.hpusm_knn2_l2(None, -1, None, 4, 128, 0, None, None, None, 0, None)
    assert rc == hpusm._lib.ERR_INVALID
    with pytest.raises(hpusm._lib.HpusmError):
        hpusm._lib.check(rc)


def test_l2_mode_constants_match_the_header():
    """The Python layer names the engines by the header's numbers (a new engine must be added in both places)."""
    import importlib
    import re
    lib = importlib.import_module("humanpose-and-urbanscene-matching_b200._lib")
    text = open(os.path.join(ROOT, "include", "hpusm.h")).read()
    header = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define HPUSM_(L2_[A-Z0-9_]+)\s+(\d+)", text)}
    assert set(header) == {"L2_AUTO", "L2_SIMT", "L2_TC_INT", "L2_TC_SPLIT", "L2_TC_I8", "L2_TC_Q16"}
    for name, value in header.items():
        assert getattr(lib, name) == value, name


def test_workspace_queries_are_pure(hpusm):
    lib = hpusm._lib.lib
    assert lib.hpusm_knn2_hamming_workspace_bytes(0, 0, 32) == 256
    assert lib.hpusm_gather_matches_workspace_bytes(5000) % 256 == 0
    assert lib.hpusm_pose_topk_workspace_bytes(10 ** 7, 16) > 0
    assert lib.hpusm_knn2_l2_workspace_bytes(1 << 20, 1 << 20, 128, 0) < (8 << 30)


def test_no_cpu_fallback(hpusm):
    """Compute entry points refuse CPU tensors instead of silently computing on the host."""
    import torch
    with pytest.raises(ValueError):
        hpusm.knn2_hamming(torch.zeros(4, 32, dtype=torch.uint8), torch.zeros(4, 32, dtype=torch.uint8))
    with pytest.raises(ValueError):
        hpusm.knn2_l2(torch.zeros(4, 128), torch.zeros(4, 128))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "humanpose-and-urbanscene-matching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_synth_generators_are_seeded(hpusm):
    a, b = hpusm.synth.sift_like(50, 80, seed=3)
    c, d = hpusm.synth.sift_like(50, 80, seed=3)
    assert np.array_equal(a, c) and np.array_equal(b, d)
    assert a.dtype == np.float32 and np.array_equal(a, np.rint(a)) and a.max() <= 255 and a.min() >= 0
    norms = np.linalg.norm(b, axis=1)
    assert 400 < norms.mean() < 620
    q, db, offs, planted = hpusm.synth.orb_database(20, per_image=64, seed=1, ragged=True)
    assert offs[0] == 0 and offs[-1] == len(db) and len(offs) == 21
    P = hpusm.synth.pose_database(hpusm.synth.CANONICAL_POSE, 100, seed=2)
    assert P.shape == (100, 18, 2) and P.dtype == np.float32


def test_openpose_json_parser_and_packed_db(hpusm, tmp_path):
    """posedb.parse_openpose_json == common.parse_JSON_multi_person (common.py:215-239) on keypoint triples recorded
    from the reference's fixtures; pack() / PoseDB round trip."""
    import json
    g = np.load(os.path.join(ROOT, "tests", "golden", "pose_json_parser.npz"))
    paths = []
    for fi in range(len(g["n"])):
        n = int(g["n"][fi])
        doc = {"version": 1.0, "people": [{"pose_keypoints": g["raw"][fi, k].tolist(), "face_keypoints": []} for k in range(n)]}
        p = tmp_path / ("img%d_keypoints.json" % fi)
        p.write_text(json.dumps(doc))
        paths.append(str(p))
        got = hpusm.posedb.parse_openpose_json(str(p))
        assert len(got) == n
        for k in range(n):
            np.testing.assert_array_equal(got[k], g["poses"][fi, k])
    out = str(tmp_path / "db.hpusm")
    total = hpusm.posedb.pack(paths, out)
    db = hpusm.posedb.PoseDB(out)
    assert len(db) == total == int(g["n"].sum()) and db.n_files == len(paths) and db.names[0] == "img0_keypoints.json"
    np.testing.assert_array_equal(db.poses[0], g["poses"][0, 0].astype(np.float32))
    assert db.file_index[-1] == len(paths) - 1 and db.person_index[0] == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
def test_dropin_install_over_reference_checkout():
    """dropin.install(reference_root): the reference's own pipeline modules import, see OUR hot-path functions, and
    still reach the reference's non-hot-path helpers through the fall-through.  (No compute here: no GPU.)"""
    import subprocess
    import sys
    code = r'''
import sys, math
import numpy as np
np.math = math
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle/refshim")
import hpusm
hpusm.dropin.install(reference_root="/root/reference")
import matching, urbanscene.urban_scene as us, posematching.multi_person as mp, common, urbanscene.features as ft
assert ft.match_and_draw.__module__.endswith("dropin.urbanscene.features")
assert us.features is ft and mp.find_transformation.__module__.endswith("dropin.common")
assert mp.match.__module__.endswith("dropin.posematching.multi_person")               # batched permutation stage
assert mp.plot_match.__module__.startswith("_hpusm_reference")                        # fall-through to the plot helper
assert common.parse_JSON_multi_person.__module__.startswith("_hpusm_reference")      # fall-through
poses = common.parse_JSON_multi_person("/root/reference/json_data/duo1.json")
assert poses[0].shape == (18, 2)
assert hasattr(ft, "explore_match")                                                   # GUI helper of the reference
det, m = ft.init_feature("orb-flann")        # the name every driver script of the reference uses (urbanscene/main.py:15)
assert type(m).__name__ == "BFMatcher" and m.normType == ft.NORM_HAMMING and det is not None
det, m = ft.init_feature("sift-flann")
assert type(m).__name__ == "BFMatcher" and m.normType == ft.NORM_L2
import os
os.environ["HPUSM_FLANN"] = "cv2"
det, m = ft.init_feature("orb-flann")
assert "Flann" in type(m).__name__
print("OK")
''' % (ROOT, ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]


def test_packed_image_database_round_trip(hpusm, tmp_path):
    """imagedb.pack_arrays / ImageDB: ragged images, an image without keypoints, 61-byte rows padded to 64."""
    rng = np.random.default_rng(2)
    counts = [5, 0, 17, 1]
    descs = [rng.integers(0, 256, (c, 61), dtype=np.uint8) if c else None for c in counts]
    kps = [rng.uniform(0, 500, (c, 2)).astype(np.float32) for c in counts]
    path = str(tmp_path / "img.hpusm")
    assert hpusm.imagedb.pack_arrays(descs, kps, path, names=["a", "b", "c", "d"], sizes=[(10, 20)] * 4) == (4, 23)
    db = hpusm.imagedb.ImageDB(path)
    assert len(db) == 4 and db.n_rows == 23 and db.width == 64 and db.names == ["a", "b", "c", "d"]
    np.testing.assert_array_equal(db.seg_offsets, [0, 5, 5, 22, 23])
    d2, k2 = db.image(2)
    np.testing.assert_array_equal(d2[:, :61], descs[2])
    assert not d2[:, 61:].any()
    np.testing.assert_array_equal(k2, kps[2])
    assert db.image(1)[0].shape == (0, 64) and tuple(db.sizes[3]) == (10, 20)
    assert [c[:2] for c in db.chunks(3)] == [(0, 3), (3, 4)]


def test_binary_descriptor_padding():
    """AKAZE's 61-byte descriptors are zero-padded to the kernel's 64-byte rows (Hamming distances unchanged)."""
    import importlib
    ft = importlib.import_module("humanpose-and-urbanscene-matching_b200.dropin.urbanscene.features")
    d = np.random.default_rng(0).integers(0, 256, (7, 61), dtype=np.uint8)
    p = ft._pad_binary(d)
    assert p.shape == (7, 64) and np.array_equal(p[:, :61], d) and not p[:, 61:].any()
    assert ft._pad_binary(d[:, :32]) .shape == (7, 32) and ft._pad_binary(d[:, :20]).shape == (7, 32)
    with pytest.raises(ValueError):
        ft._pad_binary(np.zeros((2, 65), np.uint8))


def test_public_python_surface_is_complete():
    """Every name the docs and the bench rely on exists (an edit that drops a function must fail on CPU, not on the GPU box)."""
    import importlib
    pkg = importlib.import_module("humanpose-and-urbanscene-matching_b200")
    for name in ["world", "shard_range", "shard_segments", "globalize", "allgather", "knn2_sharded", "image_scores_sharded",
                 "pose_topk_sharded", "scene_retrieval", "StreamedL2Matcher", "StreamedImageScorer", "StreamedHomographyBatch"]:
        assert hasattr(pkg.retrieval, name), "retrieval.%s is missing" % name
    for name in ["knn2_l2", "knn2_hamming", "knn2_hamming_segmented", "knn2_merge", "ratio_filter", "gather_matches",
                 "segment_matches", "ransac_homography_batch", "ransac_score_batch", "perspective_transform", "affine_lsq_batch",
                 "procrustes_batch", "pose_match_batch", "pose_match_pairs", "scene_score_batch", "pose_topk", "pose_topk_merge",
                 "L2_AUTO", "L2_SIMT", "L2_TC_INT", "L2_TC_SPLIT", "L2_TC_I8", "L2_TC_Q16"]:
        assert hasattr(pkg.engine, name), "engine.%s is missing" % name
    for name in pkg.__all__:
        assert hasattr(pkg, name), "__all__ names %s, which does not exist" % name


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--nq", "256", "--nt", "4096"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"]:
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("reference", "port")


def test_header_is_plain_c99(tmp_path):
    """include/hpusm.h is the boundary a cgo / JNI / ctypes binding sees: it must compile as C (no C++, no CUDA types),
    and a C program that only uses it must link against libhpusm.so."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "hpusm.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %s\\n", hpusm_version(), hpusm_last_error()); '
                   'return hpusm_version() == HPUSM_ABI_VERSION ? 0 : 1; }\n')
    lib_dir = os.path.join(root, "humanpose-and-urbanscene-matching_b200")
    exe = tmp_path / "t"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                        "-o", str(exe), "-L", lib_dir, "-l:libhpusm.so", "-Wl,-rpath," + lib_dir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)  # no GPU needed: version + error string only
    assert r.returncode == 0, (r.stdout, r.stderr)
