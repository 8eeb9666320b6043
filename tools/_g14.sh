set -x
for lib in r2p2_b200/csrc/libr2p2_b200.so tools/abtest/lib_minb5.so; do
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload E --ticks 200 --iters 2 2>&1 | tail -2
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload C3 --pop 8192 --iters 3 2>&1 | tail -2
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload C1 --pop 65536 --iters 2 2>&1 | tail -2
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:episode_kernel -s 1 -c 1 -o gpurun_out/r02_C3_65536_lpr1 python tools/profile_run.py --workload C3 --ticks 100 --iters 2 > gpurun_out/ncu_c3.log 2>&1
