python -m pytest tests/test_reference_traces.py -x -q 2>&1 | tail -15
