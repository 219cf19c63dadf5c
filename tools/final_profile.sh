#!/usr/bin/env bash
# One GPU call that gathers the evidence committed under profiles/ (run through gpurun; outputs land in gpurun_out/):
#   the default bench line, the ncu launch list of the same command, one `ncu --set full` capture of each dominant kernel,
#   the per-call / C1 latency tools and the instruction-rate microbenchmarks.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --min-region-ms 1"
$SHORT > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_launches.csv $SHORT > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pose_match_fast -s 1 -c 1 -o $O/r2_pose_fast_final $SHORT --workload pose > $O/ncu_pose.log 2>&1; echo "ncu pose rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ransac_score -s 1 -c 1 -o $O/r2_ransac_score_final $SHORT --workload ransac --pairs 20000 > $O/ncu_rs.log 2>&1; echo "ncu ransac rc=$?"
ncu --set full --clock-control none --import-source on -k regex:l2_i8_knn2 -s 1 -c 1 -o $O/r2_l2_i8_final $SHORT --no-others > $O/ncu_i8.log 2>&1; echo "ncu i8 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ham_tc_segmented -s 1 -c 1 -o $O/r2_ham_tc $SHORT --workload orb > $O/ncu_ham.log 2>&1; echo "ncu hamming tc rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ham_tc -c 10 --csv --log-file $O/r2_orb_tc_launches.csv $SHORT --workload orb > $O/ncu_ham_l.log 2>&1; echo "orb launch list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pose_ -c 24 --csv --log-file $O/r2_pose_launches.csv $SHORT --workload pose > $O/ncu_pose_l.log 2>&1; echo "pose launch list rc=$?"
python bench.py --workload orb --steps 10 --warmup 3 > $O/r2_bench_orb_tc.json 2> $O/r2_bench_orb_tc.err; echo "bench orb rc=$?"
ncu --set full --clock-control none --import-source on -k regex:orb_ -c 6 -o $O/r2_orb_compute python -m pytest tests/test_gpu_parity.py -q -k test_orb_compute_is_cv2 > $O/ncu_orb.log 2>&1; echo "ncu orb rc=$?"
python tests/dropin_latency.py > $O/r2_dropin_latency.json 2>/dev/null; echo "dropin latency rc=$?"
python tools/c1_latency.py > $O/r2_c1_latency.json 2>/dev/null; echo "c1 latency rc=$?"
python tools/microbench.py > $O/r2_microbench.txt 2>/dev/null; echo "microbench rc=$?"
# general-float variant of C2 (fixed-point integer engine) + the decompositions of both kNN kernels + MMA issue rates at their shapes
python bench.py --descriptors float --steps 6 --warmup 3 --no-others > $O/r2_bench_float_q16.json 2> $O/r2_bench_float_q16.err; echo "bench float rc=$?"
ncu --set full --clock-control none --import-source on -k regex:l2_q16_knn -s 1 -c 1 -o $O/r2_l2_q16 $SHORT --descriptors float --no-others > $O/ncu_q16.log 2>&1; echo "ncu q16 rc=$?"
# (the two decompositions need a library built with HPUSM_EXTRA_FLAGS=-DHPUSM_DEV_SWITCHES)
python tools/q16_time.py 262144 10 2 1 0 > $O/r2_q16_decomposition.txt 2>&1; echo "q16 decomposition rc=$?"
python tools/q16_time.py --engine i8 524288 10 2 1 0 > $O/r2_i8_decomposition.txt 2>&1; echo "i8 decomposition rc=$?"
python tools/mma_shapes.py > $O/r2_mma_shapes.txt 2>&1; echo "mma shapes rc=$?"
ls -la $O/*.ncu-rep
