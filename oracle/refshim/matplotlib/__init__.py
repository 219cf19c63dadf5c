"""Inert stand-in for matplotlib (test infrastructure only).

The reference imports matplotlib at module top (features.py:4, transformation.py:4-7,17,
procrustes.py:1, single_person.py:2-3, multi_person.py:8-9, matching.py:6) but only touches it on
``plot=True`` branches.  matplotlib is not installed in this image, so the golden-vector recorder puts
this package on ``sys.path`` to make the UNMODIFIED reference importable.  Never used by the product.
"""


class _Inert:
    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Inert()
