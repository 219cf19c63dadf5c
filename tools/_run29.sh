python bench.py --nt 131072 --steps 10 --warmup 3 --no-others --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nt131072', d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms'), d['gpu_launches'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"l2_|knn2|ratio|merge" -c 40 --csv --log-file gpurun_out/strong_launches.csv python bench.py --nt 131072 --steps 2 --warmup 1 --no-others --no-cpu-baseline > gpurun_out/ncu_strong.log 2>&1
