set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference ) > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
tail -c 200 gpurun_out/bench_r02_ref.err
( time python bench.py ) > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -c 200 gpurun_out/bench_r02_n1.err
