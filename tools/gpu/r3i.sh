timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sparams or Af_rhs" 2>&1 | tail -2
