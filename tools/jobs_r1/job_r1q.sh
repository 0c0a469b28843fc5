#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_sbr.py -m gpu -x -q 2>&1 | tail -5
