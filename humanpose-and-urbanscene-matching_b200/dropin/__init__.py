"""Drop-in mirror of the reference's hot-path modules: same module names, same call signatures, same return
types and sentinels -- bodies call the sm_100a kernels through the C ABI (no CPU fallback).

    reference module                         mirrored here
    urbanscene/features.py                   dropin.urbanscene.features
    urbanscene/transformation.py             dropin.urbanscene.transformation   (perspective_correction, affine_multi)
    posematching/procrustes.py               dropin.posematching.procrustes
    posematching/pose_comparison.py          dropin.posematching.pose_comparison
    posematching/single_person.py            dropin.posematching.single_person  (match_single)
    posematching/multi_person.py             dropin.posematching.multi_person   (match, find_ordered_matches, order_poses)
    common.py (hot subset)                   dropin.common
    thresholds.py                            dropin.thresholds

`install()` registers the modules under the reference's import names (``common``, ``thresholds``,
``urbanscene.features`` ...) so the reference's callers (urban_scene.py, multi_person.py, matching.py) pick them
up unchanged; see INTEGRATION.md.
"""
import sys


_OVERRIDES = {
    "thresholds": "thresholds.py",
    "common": "common.py",
    "urbanscene.features": "urbanscene/features.py",
    "urbanscene.transformation": "urbanscene/transformation.py",
    "posematching.pose_comparison": "posematching/pose_comparison.py",
    "posematching.procrustes": "posematching/procrustes.py",
    "posematching.single_person": "posematching/single_person.py",
    "posematching.multi_person": "posematching/multi_person.py",
}


def _fallthrough(ours, ref):
    """Names our module does not define resolve to the reference's own module (parsers, plotting, GUI helpers ...)."""
    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(ref, name)
    ours.__getattr__ = __getattr__
    ours._reference = ref


def install(reference_root=None):
    """Registers the drop-in modules under the reference's import names.  With `reference_root` (a checkout of the
    reference) every replaced module keeps a fall-through to its original, so callers that also use the parts we do
    not replace (common.parse_JSON_multi_person, features.explore_match, the plot helpers ...) keep working; packages
    `urbanscene` / `posematching` then stay the reference's own, only the hot-path submodules are swapped."""
    import importlib
    import importlib.util
    import os
    ours = {name: importlib.import_module(__name__ + "." + name) for name in _OVERRIDES}
    if reference_root is None:
        from . import urbanscene as us_pkg, posematching as pm_pkg
        sys.modules["urbanscene"] = us_pkg
        sys.modules["posematching"] = pm_pkg
        for name, mod in ours.items():
            sys.modules[name] = mod
        return
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    # thresholds and common first: the other reference modules import them at load time
    for name in _OVERRIDES:
        path = os.path.join(reference_root, _OVERRIDES[name])
        spec = importlib.util.spec_from_file_location("_hpusm_reference." + name, path)
        ref = importlib.util.module_from_spec(spec)
        sys.modules[name] = ours[name]          # visible to the reference modules loaded next
        try:
            spec.loader.exec_module(ref)
            _fallthrough(ours[name], ref)
        except Exception:                        # e.g. matplotlib missing: the override still works on its own
            pass
        if "." in name:                          # make `import urbanscene.features` find ours inside the reference package
            pkg = importlib.import_module(name.split(".")[0])
            setattr(pkg, name.split(".")[1], ours[name])
