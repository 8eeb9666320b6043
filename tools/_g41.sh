echo "== merge variant"
for w in "B" "C3 --pop 2048" "C1 --pop 1024" "C3"; do
  timeout 300 python tools/abtest/run_with_lib.py tools/abtest/lib_merge.so --workload $w --iters 4 2>&1 | grep "iter 3\|rror"
done
echo "== main (HEAD)"
for w in "B" "C3 --pop 2048" "C1 --pop 1024" "C3"; do
  python tools/profile_run.py --workload $w --iters 4 2>&1 | grep "iter 3\|rror"
done
