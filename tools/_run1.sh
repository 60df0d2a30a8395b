python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/s2_tests.txt
cat gpurun_out/s2_tests.txt
python tools/quick_bench.py > gpurun_out/s2_quick.txt 2>&1; cat gpurun_out/s2_quick.txt
python tools/parity_err.py > gpurun_out/s2_parity.txt 2>&1; tail -15 gpurun_out/s2_parity.txt
