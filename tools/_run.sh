set -u
timeout 300 python -m pytest tests/test_gpu_env.py tests/test_gpu_wrappers.py -x -q -k "host_session or fused_observation_pipelined or wrapped_pipelined or single_pass" 2>&1 | tail -15
