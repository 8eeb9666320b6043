echo "== own8 variant"
for w in "E --ticks 200" "C3 --pop 8192" "C3" "B" "D"; do
  timeout 300 python tools/abtest/run_with_lib.py tools/abtest/lib_own8.so --workload $w --iters 3 2>&1 | grep "iter 2\|rror"
done
echo "== main (previous commit)"
for w in "E --ticks 200" "C3 --pop 8192" "C3" "B" "D"; do
  python tools/profile_run.py --workload $w --iters 3 2>&1 | grep "iter 2\|rror"
done
