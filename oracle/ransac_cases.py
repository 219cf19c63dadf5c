"""Seeded synthetic RANSAC inputs shared by oracle/record_golden.py (which records what cv2 returns for them) and the tests
(which regenerate them instead of storing 1000 point sets).  Test infrastructure only."""
import importlib

import numpy as np

N_SYNTH = 1000


def ransac_synth_case(k):
    """Inputs of synthetic case k of tests/golden/ransac_synth_cv2.npz: (src [n,2] f32, dst [n,2] f32, H_true)."""
    synth = importlib.import_module("humanpose-and-urbanscene-matching_b200.synth")
    rng = np.random.default_rng(7000 + k)
    n = int(rng.integers(40, 600))
    ratio = float(rng.uniform(0.25, 0.9))
    noise = float(rng.uniform(0.2, 1.5))
    return synth.ransac_pair(n, inlier_ratio=ratio, seed=9000 + k, noise=noise)


def ransac_synth_batch(first=0, count=N_SYNTH):
    """Cases first..first+count-1 back to back: (src, dst, pair_offsets int64)."""
    srcs, dsts, offs = [], [], [0]
    for k in range(first, first + count):
        s, d, _ = ransac_synth_case(k)
        srcs.append(s); dsts.append(d); offs.append(offs[-1] + len(s))
    return np.concatenate(srcs), np.concatenate(dsts), np.array(offs, np.int64)
