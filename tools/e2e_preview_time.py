"""End-to-end frames/s of td4-psp18 at 1024x2048 with both edges of the path on the device (SURVEY.md 8f ranks 1 + 2):
pinned uint8 HWC camera frame -> H2D (6.3 MB) -> [cv2-exact resize when --src differs] -> forward_u8 (normalisation in
the stem) + forward_preview (quarter-size arg-max labels, Testing/test.py:61-64) -> D2H (0.13 MB).  One JSON line.
    python tools/e2e_preview_time.py [--src 1024 2048] [--steps 60]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", type=int, nargs=2, default=(1024, 2048))
    ap.add_argument("--steps", type=int, default=60)
    a = ap.parse_args()
    g.build()
    from tdnet_b200.ingest import FrameResizer
    from tdnet_b200.model import td4_psp18
    from tdnet_b200.synth import synth_state_dict
    H, W = 1024, 2048
    net = td4_psp18.td4_psp18(nclass=19, path_num=4, backbone="resnet18", ln_shape=(128, 256)).eval()
    net.load_state_dict(synth_state_dict(net.state_dict(), seed=0), strict=True)
    net.to("cuda:0")
    src = tuple(a.src)
    resize = FrameResizer(src, (H, W), "cuda:0") if src != (H, W) else None
    gen = torch.Generator().manual_seed(0)
    host = [torch.randint(0, 256, (1, src[0], src[1], 3), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(8)]
    out_host = [torch.empty((1, H // 4, W // 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    copy_s = torch.cuda.Stream()
    main_s = torch.cuda.current_stream()
    dev_in = [torch.empty_like(host[0], device="cuda") for _ in range(2)]
    ev_in, ev_used = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]

    def prefetch(i, slot, first=False):
        with torch.cuda.stream(copy_s):
            if not first:
                copy_s.wait_event(ev_used[slot])
            dev_in[slot].copy_(host[i % 8], non_blocking=True)
            ev_in[slot].record(copy_s)

    def loop(steps, step0):
        prefetch(step0, 0, first=True)
        for i in range(steps):
            slot = i % 2
            main_s.wait_event(ev_in[slot])
            frame = resize(dev_in[slot]) if resize else dev_in[slot]
            labels = net.forward_preview(frame, pos_id=(step0 + i) % 4, u8=True)
            ev_used[slot].record(main_s)
            if i + 1 < steps:
                prefetch(step0 + i + 1, slot ^ 1, first=(i == 0))
            out_host[slot].copy_(labels, non_blocking=True)
        return step0 + steps

    step = loop(16, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step = loop(a.steps, step)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    net.check_numeric_range()
    print(json.dumps({"pipeline": "uint8 frame H2D -> " + ("resize -> " if resize else "") + "forward_u8 + preview -> D2H",
                      "src": list(src), "ms_per_frame": round(ms, 4), "frames_per_s": round(1000.0 / ms, 1),
                      "h2d_bytes": src[0] * src[1] * 3, "d2h_bytes": (H // 4) * (W // 4)}))


if __name__ == "__main__":
    main()
