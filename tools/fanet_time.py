"""Device time of one TD2-FANet call (frame pair -> logits of the current frame) at a given size; prints one JSON line.
    python tools/fanet_time.py [--backbone resnet18] [--size 1024 2048] [--steps 30] [--warmup 6] [--labels]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="resnet18")
    ap.add_argument("--size", type=int, nargs=2, default=(1024, 2048))
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--labels", action="store_true")
    a = ap.parse_args()
    g.build()
    from tdnet_b200.model import td2_fa
    from tdnet_b200.model.arch import feature_hw
    from tdnet_b200.synth import synth_clip, synth_state_dict
    H, W = a.size
    net = td2_fa.td2_fa(nclass=19, backbone=a.backbone, path_num=2, ln_shape=feature_hw(H, W)).eval()
    net.load_state_dict(synth_state_dict(net.state_dict(), seed=0), strict=True)
    net.to("cuda:0")
    frames = [f.cuda() for f in synth_clip(5, H, W, clip_id=0)]
    call = net.forward_labels if a.labels else net.forward
    for i in range(a.warmup):
        call([frames[i % 4], frames[i % 4 + 1]], pos_id=i % 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        call([frames[i % 4], frames[i % 4 + 1]], pos_id=i % 2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    net.check_numeric_range()
    eng, plan = net._last
    print(json.dumps({"model": f"td2_fa-{a.backbone}", "size": [H, W], "ms_per_call": ms, "calls_per_s": 1000.0 / ms,
                      "kernel_launches_per_call": plan.kernel_launches, "labels": a.labels}))


if __name__ == "__main__":
    main()
