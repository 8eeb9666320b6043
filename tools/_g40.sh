timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -5
echo "== E 32768 robots x 1000 (the 8-GPU shard): sliced vs not"
timeout 300 python tools/profile_run.py --workload E --pop 32768 --iters 3 2>&1 | grep "iter 2\|rror"
R2P2_NO_SLICES=1 timeout 300 python tools/profile_run.py --workload E --pop 32768 --iters 3 2>&1 | grep "iter 2\|rror"
echo "== E 65536 x 1000 (4-GPU shard)"
timeout 300 python tools/profile_run.py --workload E --pop 65536 --iters 3 2>&1 | grep "iter 2\|rror"
R2P2_NO_SLICES=1 timeout 300 python tools/profile_run.py --workload E --pop 65536 --iters 3 2>&1 | grep "iter 2\|rror"
echo "== E full x 200"
timeout 300 python tools/profile_run.py --workload E --ticks 200 --iters 3 2>&1 | grep "iter 2\|rror"
