set -x
python tools/profile_run.py --workload E --ticks 20 --iters 2 > gpurun_out/plain_E.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:episode_kernel -s 1 -c 1 -o gpurun_out/r02c_E_lpr8 python tools/profile_run.py --workload E --ticks 20 --iters 2 > gpurun_out/ncu_E.log 2>&1
tail -n 3 gpurun_out/plain_E.log
python tools/profile_run.py --workload C3 --ticks 100 --iters 2 > gpurun_out/plain_C3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:episode_kernel -s 1 -c 1 -o gpurun_out/r02c_C3_lpr1 python tools/profile_run.py --workload C3 --ticks 100 --iters 2 > gpurun_out/ncu_C3.log 2>&1
tail -n 3 gpurun_out/plain_C3.log
python tools/profile_run.py --workload B --iters 2 > gpurun_out/plain_B.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:episode_kernel -s 1 -c 1 -o gpurun_out/r02c_B_lpr32 python tools/profile_run.py --workload B --iters 2 > gpurun_out/ncu_B.log 2>&1
tail -n 3 gpurun_out/plain_B.log
( time python bench.py ) > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -c 400 gpurun_out/bench_r02_n1.err
