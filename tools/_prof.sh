prof() {  # name workload extra-args
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:episode_kernel -s 1 -c 1 -o /tmp/$1 python tools/profile_run.py --workload $2 $3 --iters 2 > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/$1_summary.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page source --csv -k regex:episode_kernel 2>/dev/null | gzip -9 > gpurun_out/$1_source.csv.gz
}
prof r02e_E_lpr8 E "--ticks 20"
prof r02e_C3_lpr1 C3 "--ticks 100"
prof r02e_C3_8192_lpr8 C3 "--pop 8192 --ticks 300"
prof r02e_B_lpr32 B ""
ls gpurun_out | grep r02e
