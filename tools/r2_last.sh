#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_model.py -x -q -m gpu -p no:cacheprovider -k "precise_decode or caller_position_ids or paged" 2>&1 | tail -4
