set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "hamming" 2>&1 | tail -3
python bench.py --workload orb --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/orb_tc.err > gpurun_out/r2_bench_orb_tc.json; cut -c1-600 gpurun_out/r2_bench_orb_tc.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_orb_tc_launches.csv python bench.py --workload orb --steps 2 --warmup 3 --no-cpu-baseline --min-region-ms 0 > gpurun_out/ncu_orb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ham_tc_segmented -s 2 -c 1 -o gpurun_out/r2_ham_tc python bench.py --workload orb --steps 2 --warmup 3 --no-cpu-baseline --min-region-ms 0 > gpurun_out/ncu_orb2.log 2>&1
ls -la gpurun_out | tail -5
