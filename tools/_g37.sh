python tools/sweep_lpr.py --out gpurun_out/r02_lpr_sweep.json 2>&1 | grep -v Warning
