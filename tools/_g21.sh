set -x
python tools/profile_run.py --workload C3 --ticks 100 --iters 2 > gpurun_out/plain_C3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:episode_kernel -s 1 -c 1 -o gpurun_out/r02b_C3_lpr1 python tools/profile_run.py --workload C3 --ticks 100 --iters 2 > gpurun_out/ncu_C3.log 2>&1
tail -n 3 gpurun_out/plain_C3.log
