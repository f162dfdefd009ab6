#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x -k "cholesky or likelihood or positive or gp_fit or appendix or cfg1 or bo_loop or multi_kernel or layer" 2>&1 | tail -5
python profiles/factor_bench.py 2>&1 | tail -80
