#!/bin/bash
# Device-side cv2-exact frame resize + quarter-size label preview: tests, end-to-end timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== ingest tests"; timeout 400 python -m pytest tests/test_ingest_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -k "ingest or preview or resize or forward_u8 or forward_labels" 2>&1 | tail -25
echo "== e2e preview"; timeout 200 python tools/e2e_preview_time.py 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/e2e_preview.json
timeout 200 python tools/e2e_preview_time.py --src 1080 1920 2>&1 | tail -1 | cut -c1-300 | tee -a gpurun_out/e2e_preview.json
