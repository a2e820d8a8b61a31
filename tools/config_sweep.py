"""Throughput of the BASELINE.json configurations that are parity-test cases rather than the bench line (configs[0],
[3], [4]) plus the comparison models, on one GPU: frames/s with frames resident in HBM, CUDA events, steady state.
One JSON line per configuration.   python tools/config_sweep.py [--steps 24]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

CONFIGS = [  # label, arch, backbone, H, W, batch
    ("configs[0] td2-psp50 512x1024", "td2_psp50", "resnet50", 512, 1024, 1),
    ("configs[3] td2-'bise34' = td2_psp50(resnet34) 720x960", "td2_psp50", "resnet34", 720, 960, 1),
    ("configs[4] td4-psp50 = td4_psp18(resnet50) 1024x2048 n=1", "td4_psp18", "resnet50", 1024, 2048, 1),
    ("configs[4] td4-psp50 1024x2048 n=2", "td4_psp18", "resnet50", 1024, 2048, 2),
    ("configs[4] td4-psp50 1024x2048 n=4", "td4_psp18", "resnet50", 1024, 2048, 4),
    ("pspnet-101 1024x2048 (comparison model, TEST_README.md:31)", "pspnet", "resnet101", 1024, 2048, 1),
    # the size of the reference's published table (Testing/TEST_README.md:31-33, Titan Xp: 360 / 180 / 85 ms per frame)
    ("published-table size: td4-psp18 769x1537", "td4_psp18", "resnet18", 769, 1537, 1),
    ("published-table size: td2-psp50 769x1537", "td2_psp50", "resnet50", 769, 1537, 1),
    ("published-table size: pspnet-101 769x1537", "pspnet", "resnet101", 769, 1537, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--only", type=int, nargs="*")
    a = ap.parse_args()
    g.build()
    from tdnet_b200.model import pspnet, td2_psp50, td4_psp18
    from tdnet_b200.model.arch import feature_hw
    from tdnet_b200.synth import synth_state_dict
    for ci, (label, arch, backbone, H, W, n) in enumerate(CONFIGS):
        if a.only and ci not in a.only:
            continue
        h8, w8 = feature_hw(H, W)
        if arch == "pspnet":
            net = pspnet.pspnet(nclass=19, backbone=backbone)
        elif arch == "td4_psp18":
            net = td4_psp18.td4_psp18(nclass=19, path_num=4, backbone=backbone, ln_shape=(h8, w8))
        else:
            net = td2_psp50.td2_psp50(nclass=19, path_num=2, backbone=backbone, ln_shape=(h8, w8))
        net.load_state_dict(synth_state_dict(net.state_dict(), seed=0), strict=True)
        net.eval().to("cuda:0")
        gen = torch.Generator(device="cuda").manual_seed(ci)
        frames = [torch.randn(n, 3, H, W, generator=gen, device="cuda") for _ in range(4)]
        paths = net.PATHS
        for i in range(3 * paths + 2):                  # fill the FIFO, build every plan, capture every graph
            net(frames[i % 4], pos_id=i % paths)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            net(frames[i % 4], pos_id=i % paths)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        net.check_numeric_range()
        print(json.dumps({"config": label, "batch": n, "ms_per_step": round(ms, 4),
                          "frames_per_s": round(1000.0 * n / ms, 2), "steps": a.steps,
                          "mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}), flush=True)
        del net, frames
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
