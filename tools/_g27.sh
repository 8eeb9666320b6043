set -x
for w in C3 D E; do
python bench.py --only --no-cpu --steps 4 --workload $w 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['fitness_equal_to_resident_run'], 'frac', d['roofline']['frac'])"
done
R2P2_NO_BALANCE=1 python bench.py --only --no-cpu --steps 4 --workload C3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3 nobal value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['fitness_equal_to_resident_run'], 'frac', d['roofline']['frac'])"
python bench.py --only --no-cpu --steps 4 --workload C3 --pop 8192 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3/8 value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['fitness_equal_to_resident_run'], 'frac', d['roofline']['frac'])"
R2P2_NO_BALANCE=1 python bench.py --only --no-cpu --steps 4 --workload C3 --pop 8192 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3/8 nobal value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['fitness_equal_to_resident_run'], 'frac', d['roofline']['frac'])"
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
