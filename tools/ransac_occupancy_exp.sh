set -e
cd ${GRAFT_REPO_ROOT:-.}
P='import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["find_homography_resident"]["value"])'
for v in 2 3; do
  touch humanpose-and-urbanscene-matching_b200/csrc/ransac.cu
  HPUSM_EXTRA_FLAGS="-DRS_MIN_BLOCKS=$v" bash humanpose-and-urbanscene-matching_b200/csrc/build.sh > /dev/null 2>&1
  echo "RS_MIN_BLOCKS=$v"
  python bench.py --workload ransac --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"
done
python -m pytest tests -m gpu -x -q -k "ransac" 2>&1 | tail -3
