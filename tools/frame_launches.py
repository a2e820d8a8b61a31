"""Every kernel (this library's AND torch's) of a few steady-state td4-psp18 frames at 1024x2048, for an ncu launch list
WITHOUT a kernel-name filter -- the check that nothing but the frame plan runs inside a frame:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/frame_all.csv \\
        python tools/frame_launches.py
    python tools/launch_summary.py gpurun_out/frame_all.csv

CUDA graphs are switched off so that every kernel is an individual launch; a cudaProfilerStart/Stop bracket limits the
capture to the last frames (run ncu with --profile-from-start off)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TDNET_B200_CUDA_GRAPH"] = "0"

import torch  # noqa: E402

import bench  # noqa: E402
from tdnet_b200.model import td4_psp18  # noqa: E402
from tdnet_b200.model.arch import feature_hw  # noqa: E402
from tdnet_b200.synth import synth_clip  # noqa: E402


def main(frames=4):
    H, W = 1024, 2048
    h8, w8 = feature_hw(H, W)
    net = td4_psp18.td4_psp18(nclass=19, path_num=4, backbone="resnet18", ln_shape=(h8, w8)).eval()
    net.load_state_dict(bench._weights(h8, w8), strict=True)
    net.to("cuda:0")
    clip = [f.cuda() for f in synth_clip(4, H, W)]
    for i in range(8):
        net(clip[i % 4], pos_id=i % 4)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(8, 8 + frames):
        net(clip[i % 4], pos_id=i % 4)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
