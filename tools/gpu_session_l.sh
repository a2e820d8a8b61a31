#!/bin/bash
# TD2-FANet with the fused tcgen05 LeakyReLU stem and stacked q/k projections; stem regression tests; launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== fanet + stem tests"; timeout 500 python -m pytest tests/test_fanet_gpu.py tests/test_ops_gpu.py -q -m gpu --tb=short -k "fanet or fa_ or add_upsampled or stem" 2>&1 | tail -40
echo "== timing"; timeout 200 python tools/fanet_time.py 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/fanet_time.json
timeout 200 python tools/fanet_time.py --labels 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/fanet_time.json
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tdn|tc_|conv_simt|fa_|stem|psp_|ln_|copy_nhwc|bilinear|maxpool|image_to|upsample|add_upsampled|softmax' -c 600 --csv --log-file gpurun_out/fanet_launches.csv python tools/fanet_time.py --steps 2 --warmup 2 > gpurun_out/fanet_ncu.log 2>&1; tail -1 gpurun_out/fanet_ncu.log | cut -c1-200
