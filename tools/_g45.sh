timeout 900 ncu --set full --clock-control none -k regex:episode_kernel -s 1 -c 1 -o /tmp/c3raw python tools/profile_run.py --workload C3 --ticks 100 --iters 2 > /dev/null 2>&1
ncu -i /tmp/c3raw.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr,units,d=rows[0],rows[1],rows[2]
for h,u,v in zip(hdr,units,d):
    if any(k in h for k in ('l1tex__t_sector','l1tex__t_requests','lts__t_sectors_srcunit_tex','lts__t_sector_hit','l1tex__data_bank','lts__throughput','l1tex__throughput','lts__t_bytes','sm__throughput','l1tex__lsu_writeback','smsp__inst_executed_op_global','smsp__inst_executed_op_shared','l1tex__t_bytes','lts__t_sectors.sum','lts__t_sectors_op','dram__throughput','l1tex__data_pipe')):
        print('%-95s %-14s %s' % (h,u,v))
" > gpurun_out/c3_raw_mem.txt
wc -l gpurun_out/c3_raw_mem.txt
