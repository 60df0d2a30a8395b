#!/bin/bash
timeout 600 python -m pytest tests/test_fuzz_gpu.py -q -x -m gpu 2>&1 | tail -8
