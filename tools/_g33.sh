for b in 256 384 448 512; do
  echo "== E block $b"; R2P2_BLOCK=$b python tools/profile_run.py --workload E --ticks 200 --iters 3 2>&1 | grep "iter 2\|rror"
done
for b in 320 384 448 512; do
  echo "== C3 block $b"; R2P2_BLOCK=$b python tools/profile_run.py --workload C3 --iters 3 2>&1 | grep "iter 2\|rror"
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -3
