python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --workload B --only > gpurun_out/n2_B.out 2> gpurun_out/n2_B.err
echo rc=$?
tail -n 30 gpurun_out/n2_B.err; tail -c 600 gpurun_out/n2_B.out
