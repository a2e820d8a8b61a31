"""Wave-quantisation probe of the pair conv kernel (layer-4 shape: 3x3 512->512, dilation 2, map h x 256): the launch time as a
function of the number of M256xN256 pair tiles (74 CTA pairs per round).  If the time follows the ROUNDS (ceil(tiles / 74)) a
ragged last round costs a full tile time and splitting it pays; if it follows the WORK (tiles) it is free (power-limited clock).

    timeout 200 python tools/quant_probe.py [h ...]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tdnet_b200 import _cabi as cabi  # noqa: E402


def main():
    lib = cabi.load()
    hs = [int(a) for a in sys.argv[1:]] or [72, 80, 104, 112, 128, 144, 152]
    cin = cout = 512
    w = 256
    g = torch.Generator().manual_seed(1)
    wt = (torch.randn(cout, 9 * cin, generator=g) / (9 * cin) ** 0.5).cuda()
    wh = wt.half().contiguous()
    wl = (wt - wh.float()).half().contiguous()
    for h in hs:
        x = torch.randn(1, h, w, cin, generator=g).cuda()
        xh = x.half().contiguous()
        xl = (x - xh.float()).half().contiguous()
        oh, ol = torch.empty(1, h, w, cout, dtype=torch.half, device="cuda"), torch.empty(1, h, w, cout, dtype=torch.half, device="cuda")
        d = cabi.TcConvDesc()
        d.in_ = cabi.Tensor(xh.data_ptr(), xl.data_ptr(), 1, 1, h, w, cin, h * w * cin, w * cin, cin)
        d.out = cabi.Tensor(oh.data_ptr(), ol.data_ptr(), 1, 1, h, w, cout, h * w * cout, w * cout, cout)
        d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), 9 * cin
        d.cout, d.kh, d.kw, d.dilation, d.stride = cout, 3, 3, 2, 0
        ms = []
        for reps in (1, 20):
            for _ in range(3):
                cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv")
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(5 if reps == 1 else 1):
                e0.record()
                for _ in range(reps):
                    lib.tdn_conv2d_tc(C.byref(d), None)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / reps)
            ms.append(round(best, 4))
        import hashlib
        digest = hashlib.sha256(oh.cpu().numpy().tobytes() + ol.cpu().numpy().tobytes()).hexdigest()[:12]
        tiles = (h * w // 128 + 1) // 2 * 2
        print(json.dumps({"h": h, "pair_tiles": tiles, "rounds": round(tiles / 74, 2), "sha": digest, "ms_single_best_of_5": ms[0], "ms_x20": ms[1],
                          "ms_per_tile_round": round(ms[1] / -(-tiles // 74), 4), "us_per_tile_work": round(1e3 * ms[1] / tiles * 74, 2)}), flush=True)


if __name__ == "__main__":
    main()
