#!/bin/bash
mkdir -p gpurun_out
{
echo "== PDL default"; timeout 200 python tools/e2e_pipe_probe.py 2>&1 | grep -E "rep [0-7]|H2D" | tail -5
echo "== PDL=0"; NGRTD_PDL=0 timeout 200 python tools/e2e_pipe_probe.py 2>&1 | grep -E "rep [0-7]|H2D" | tail -5
} | tee gpurun_out/r2_e2e7.txt
