set -x
timeout 900 python -m pytest tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -8
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --only --no-cpu --steps 4 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('E value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['fitness_equal_to_resident_run'], 'frac', d['roofline']['frac'])"
python bench.py --only --no-cpu --steps 4 --workload C3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3 value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['fitness_equal_to_resident_run'], 'frac', d['roofline']['frac'])"
