N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err
echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_n$N.json').read().strip().splitlines()[-1]); print('N=$N E ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'NS ms', d['north_star']['ms_per_step'], d['north_star']['sonar_samples_per_s'], 'NS e2e', d['north_star']['e2e']['ms_per_step'])"
