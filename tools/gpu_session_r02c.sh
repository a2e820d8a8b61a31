cd /root/repo
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -x -q -m gpu -k "fast_mode or submodule or attention or tc_conv or golden" 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --steps 40 --warmup 8 2>&1 | tail -1 | tee gpurun_out/r02c_bench.json | cut -c1-300
echo "== ncu layer1"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 1 -o gpurun_out/r02_prof_conv_layer1 python tools/tc_probe.py --one layer1_perf 2>&1 | tail -2 | cut -c1-400
