python tools/profile_run.py --workload E --ticks 200 --iters 3 2>&1 | grep "iter 2\|rror"
python tools/profile_run.py --workload C3 --iters 3 2>&1 | grep "iter 2\|rror"
python tools/profile_run.py --workload C3 --pop 8192 --iters 3 2>&1 | grep "iter 2\|rror"
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
