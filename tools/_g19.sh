set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for w in "E --ticks 200" "C3 --pop 8192" C3 B D C1 B64k; do
  timeout 300 python tools/profile_run.py --workload $w --iters 3 2>&1 | grep "iter 2"
done
