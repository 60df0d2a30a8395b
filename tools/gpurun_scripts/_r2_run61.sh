#!/bin/bash
# flakiness check: the GPU suite three times in a row on one box
for i in 1 2 3; do timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -1; done
