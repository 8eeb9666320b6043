set -x
timeout 2400 python -m pytest tests/test_host_controllers.py tests/test_gpu_dropin.py tests/test_logs.py tests/test_goal.py -x -q -m gpu 2>&1 | tail -30
