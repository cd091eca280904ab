#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_dp.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
