#!/usr/bin/env python
"""Prints the in-kernel timestamp trace of l2_i8_knn2_kernel (CTA 0, database tiles 6000..6063 of its first unit).

Build the trace variant and run one full-size pass on the GPU box:

    cd humanpose-and-urbanscene-matching_b200/csrc
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
         -cudart static -DI8_TRACE -c l2_tc_i8.cu -o build/l2_tc_i8.o && bash build.sh
    HPUSM_I8_TRACE=gpurun_out/i8_trace.bin python bench.py --steps 1 --warmup 0 --no-cpu-baseline
    python tools/i8_trace.py gpurun_out/i8_trace.bin

Per piece (database tile, query tile): when the MMA warp started / stopped waiting for the accumulator and finished
issuing, when the eight epilogue warps saw the accumulator full, released it and finished scanning it (cycles).
"""
import numpy as np, sys
raw = np.fromfile(sys.argv[1], dtype=np.int64)
a = raw[:1024].reshape(-1, 8)
b = raw[1024:1024 + 128 * 32].reshape(128, 32)
t0 = a[:, :3][a[:, :3] > 0].min()
prev = None
for i in range(20, 44):
    r = a[i]
    arr = b[i, 0:8]; wstart = b[i, 8:16]; done = b[i, 16:24]; full = b[i, 24:32]
    print("%3d,%d mma: wait %6d -> %6d (+%4d) issued +%4d | epi: full seen %6d..+%3d  arrive last +%4d  proc done +%4d | wait-start min %6d" % (
        i // 2, i % 2, r[0] - t0, r[1] - t0, r[1] - r[0], r[2] - r[1], full.min() - t0, full.max() - full.min(), arr.max() - full.min(),
        done.max() - full.min(), wstart.min() - t0))
print("cycles per piece", (a[126, 1] - a[2, 1]) / 124.0)
