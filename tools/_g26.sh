set -x
prof() {  # name workload ticks key
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:episode_kernel -s 1 -c 1 -o /tmp/$1 python tools/profile_run.py --workload $2 --ticks $3 --iters 2 > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/$1_summary.txt 2>&1
  python tools/ncu_by_line.py /tmp/$1.ncu-rep "$4" 160 > gpurun_out/$1_by_line.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page source --csv -k regex:episode_kernel 2>/dev/null | gzip -9 > gpurun_out/$1_source.csv.gz
  ls -la /tmp/$1.ncu-rep gpurun_out/
}
prof r02c_E_lpr8 E 20 episode_kernel_ws
prof r02c_C3_lpr1 C3 100 episode_kernel
prof r02c_B_lpr32 B 1000 episode_kernel
( time python bench.py ) > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -c 300 gpurun_out/bench_r02_n1.err
