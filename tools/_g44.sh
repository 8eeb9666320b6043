for w in "C3 --pop 8192" "C3 --pop 16384" "C3" "D" "E --ticks 200" "B"; do
  python tools/profile_run.py --workload $w --iters 4 2>&1 | grep "iter 3\|rror"
done
