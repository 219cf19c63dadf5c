"""CPU oracle of the matching hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (humanpose-and-urbanscene-matching_b200/) never does: it fails loudly when
libhpusm.so or a GPU is missing.

Parity pinning: the reference ships no tests and no golden vectors, so the pins are
outputs of the UNMODIFIED reference modules + the installed OpenCV/NumPy binaries, recorded in this
container by oracle/record_golden.py into tests/golden/*.npz.  tests/test_oracle_golden.py checks every
oracle function against them.  The one part that cannot be pinned on cv2 bit-for-bit is the RANSAC
sample sequence (cv2's sampler is internal and unseedable); there the oracle (ransac_oracle.c) shares a
counter-based sample table with the GPU, and cv2 pins the final H / inlier mask on inputs with a unique
consensus set.
"""
