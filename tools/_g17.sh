set -x
python tools/profile_run.py --workload E --ticks 20 --iters 2 > gpurun_out/plain_E.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:episode_kernel -s 1 -c 1 -o gpurun_out/r02b_E_lpr8 python tools/profile_run.py --workload E --ticks 20 --iters 2 > gpurun_out/ncu_E.log 2>&1
tail -n 3 gpurun_out/plain_E.log
