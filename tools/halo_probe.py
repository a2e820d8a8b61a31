"""3x3 convolution variants on the layer-1 / layer-2 shapes of td4-psp18 at 1024x2048: per-tap loads (tc_conv.cu), the no-swizzle
halo region (tc_conv_halo.cu) and the 128-byte-swizzled halo region (tc_conv_halo_sw.cu): bit-identity with the per-tap kernel
and time per launch (30 launches back to back).

    timeout 200 python tools/halo_probe.py"""
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
    shapes = [(1, 256, 512, 64, 64, 1), (1, 128, 256, 128, 128, 1), (1, 128, 256, 128, 128, 2), (2, 45, 77, 64, 128, 2),
              (1, 128, 256, 256, 256, 2), (1, 128, 256, 512, 128, 1), (1, 128, 256, 512, 512, 4), (1, 128, 256, 256, 512, 2),
              (1, 128, 256, 512, 512, 2), (2, 45, 77, 64, 256, 3)]
    if len(sys.argv) > 1:
        shapes = [s for s in shapes if s[4] % 256 == 0]
    g = torch.Generator().manual_seed(3)
    for n, h, w, cin, cout, dil in shapes:
        wt = (torch.randn(cout, 9 * cin, generator=g) / (9 * cin) ** 0.5).cuda()
        wh = wt.half().contiguous()
        wl = (wt - wh.float()).half().contiguous()
        x = torch.randn(n, h, w, cin, generator=g).cuda()
        xh = x.half().contiguous()
        xl = (x - xh.float()).half().contiguous()
        res = {"shape": [n, h, w, cin, cout, dil]}
        ref = None
        for name, variant in (("base", cabi.TC_BASE), ("base_ts", cabi.TC_BASE_TS), ("halo", cabi.TC_HALO), ("halo_sw", cabi.TC_HALO_SW), ("pair", cabi.TC_PAIR), ("band", cabi.TC_PAIR_BAND)):
            oh = torch.full((n, h, w, cout), float("nan"), dtype=torch.half, device="cuda")
            ol = torch.full((n, h, w, cout), float("nan"), dtype=torch.half, device="cuda")
            d = cabi.TcConvDesc()
            d.in_ = cabi.Tensor(xh.data_ptr(), xl.data_ptr(), 1, n, h, w, cin, h * w * cin, w * cin, cin)
            d.out = cabi.Tensor(oh.data_ptr(), ol.data_ptr(), 1, n, h, w, cout, h * w * cout, w * cout, cout)
            d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), 9 * cin
            d.cout, d.kh, d.kw, d.dilation, d.stride, d.variant = cout, 3, 3, dil, 0, variant
            rc = lib.tdn_conv2d_tc(C.byref(d), None)
            if rc:
                res[name] = "rc %d %s" % (rc, lib.tdn_last_error().decode()[:80])
                continue
            torch.cuda.synchronize()
            out = oh.float() + ol.float()
            if ref is None:
                ref = out
            if name == "halo":
                ref_halo = out
            if name == "band" and dil <= 2:
                res["band_mismatch_vs_halo"] = int((out != ref_halo).sum())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                lib.tdn_conv2d_tc(C.byref(d), None)
            e0.record()
            for _ in range(30):
                lib.tdn_conv2d_tc(C.byref(d), None)
            e1.record()
            torch.cuda.synchronize()
            res[name] = {"us": round(1e3 * e0.elapsed_time(e1) / 30, 2), "nan": int(torch.isnan(out).sum()),
                         "mismatch_vs_base": int((out != ref).sum()), "max_diff": float((out - ref).abs().nan_to_num(9e9).max())}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
