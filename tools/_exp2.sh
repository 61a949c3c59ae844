VIP_B200_TIMING=1 python tools/run_configs.py c3 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
