"""Drop-in mirror of the reference's hot-path modules: same module names, same call signatures, same return
types and sentinels -- bodies call the sm_100a kernels through the C ABI (no CPU fallback).

    reference module                         mirrored here
    urbanscene/features.py                   dropin.urbanscene.features
    urbanscene/transformation.py             dropin.urbanscene.transformation   (perspective_correction, affine_multi)
    posematching/procrustes.py               dropin.posematching.procrustes
    posematching/pose_comparison.py          dropin.posematching.pose_comparison
    posematching/single_person.py            dropin.posematching.single_person  (match_single)
    common.py (hot subset)                   dropin.common
    thresholds.py                            dropin.thresholds

`install()` registers the modules under the reference's import names (``common``, ``thresholds``,
``urbanscene.features`` ...) so the reference's callers (urban_scene.py, multi_person.py, matching.py) pick them
up unchanged; see INTEGRATION.md.
"""
import sys


def install():
    from . import common, thresholds
    from .urbanscene import features, transformation
    from .posematching import pose_comparison, procrustes, single_person
    from . import urbanscene as us_pkg, posematching as pm_pkg
    sys.modules["thresholds"] = thresholds
    sys.modules["common"] = common
    sys.modules["urbanscene"] = us_pkg
    sys.modules["urbanscene.features"] = features
    sys.modules["urbanscene.transformation"] = transformation
    sys.modules["posematching"] = pm_pkg
    sys.modules["posematching.pose_comparison"] = pose_comparison
    sys.modules["posematching.procrustes"] = procrustes
    sys.modules["posematching.single_person"] = single_person
