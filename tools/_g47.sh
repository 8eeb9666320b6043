echo "== E x200: tail weights in smem (default) / off / off+no tanh smem"
python tools/profile_run.py --workload E --ticks 200 --iters 4 2>&1 | grep "iter 3\|rror"
R2P2_NO_WTAIL=1 python tools/profile_run.py --workload E --ticks 200 --iters 4 2>&1 | grep "iter 3\|rror"
R2P2_NO_WTAIL=1 R2P2_NO_TANH_SMEM=1 python tools/profile_run.py --workload E --ticks 200 --iters 4 2>&1 | grep "iter 3\|rror"
echo "== E 32768 x 1000"
python tools/profile_run.py --workload E --pop 32768 --iters 3 2>&1 | grep "iter 2\|rror"
R2P2_NO_WTAIL=1 python tools/profile_run.py --workload E --pop 32768 --iters 3 2>&1 | grep "iter 2\|rror"
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -2
