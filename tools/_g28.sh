for v in mb640 mb768; do
  echo "== $v"
  for w in E; do
    timeout 300 python tools/abtest/run_with_lib.py tools/abtest/lib_$v.so --workload $w --ticks 200 --iters 3 2>&1 | grep "iter 2\|rror"
  done
done
echo "== main"
python tools/profile_run.py --workload E --ticks 200 --iters 3 2>&1 | grep "iter 2\|rror"
python tools/profile_run.py --workload C3 --iters 3 2>&1 | grep "iter 2\|rror"
python tools/profile_run.py --workload D --iters 3 2>&1 | grep "iter 2\|rror"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_population_golden.py -x -q -m gpu 2>&1 | tail -3
