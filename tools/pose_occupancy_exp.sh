set -e
cd $GRAFT_REPO_ROOT
P='import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
for v in 4 5 6; do
  touch humanpose-and-urbanscene-matching_b200/csrc/pose.cu
  HPUSM_EXTRA_FLAGS="-DPF_MIN_BLOCKS=$v" bash humanpose-and-urbanscene-matching_b200/csrc/build.sh > /dev/null 2>&1
  echo "PF_MIN_BLOCKS=$v"
  python bench.py --workload pose --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"
done
