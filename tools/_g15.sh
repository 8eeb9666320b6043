set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -4
for w in B "C3 --pop 8192" C3 B64k D "E --ticks 200"; do
  timeout 300 python tools/profile_run.py --workload $w --iters 3 2>&1 | tail -2
  R2P2_NO_TANH_SMEM=1 timeout 300 python tools/profile_run.py --workload $w --iters 3 2>&1 | tail -2
done
