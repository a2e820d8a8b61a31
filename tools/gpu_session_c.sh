#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; cut -c1-420 gpurun_out/probe.log
