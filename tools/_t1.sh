timeout 600 python -m pytest tests/test_golden_running_stats.py tests/test_golden_gae.py -m gpu -x -q 2>&1 | tail -12
