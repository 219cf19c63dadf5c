#!/usr/bin/env python
"""Issue rates of the instructions the hot kernels are made of (hpusm_microbench_*): LOP3 / POPC for the Hamming
kernel's two-pipe roofline, 2- and 3-input integer / float max for the max-scan epilogues of the L2 engines, and the
tensor pipe itself (back-to-back tcgen05.mma with bf16 and with u8 operands).
Output of round 1 on a B200: profiles/r1_microbench.txt."""
import torch, importlib, sys
sys.path.insert(0, ".")
import hpusm
dev = torch.device("cuda:0")
for op in ["lop3", "popc", "imax3", "fmax3", "imax", "fmax", "ffma", "ffma2"]:
    for thr in (128, 512):
        iters = 1 << 16
        blocks = 148 * (1024 // thr)
        hpusm.engine.microbench_popc(blocks, thr, iters, dev, op=op)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); hpusm.engine.microbench_popc(blocks, thr, iters, dev, op=op); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ops = blocks * thr * iters
        print(op, thr, "%.1f Gop/s" % (ops / ms / 1e6), "-> %.1f lanes/clk/SM at 1.9 GHz" % (ops / ms / 1e6 / 148 / 1.9))
for kind in ("bf16", "i8", "i8n128", "i8n128t", "i8n64", "i8n64t"):
    for rounds in (2000, 20000):
        hpusm.engine.microbench_mma(kind, 200, dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops = hpusm.engine.microbench_mma(kind, rounds, dev); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("mma %s rounds %d: %.1f ms -> %.0f T op/s" % (kind, rounds, ms, ops / ms / 1e9))
