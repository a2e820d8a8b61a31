"""Probe for the EXPERIMENTAL 2-CTA-cluster attention kernel (tools/experimental/tc_attn_cluster.cu) -- run on a B200:

    bash tools/experimental/build.sh && timeout 120 python tools/experimental/attn_cluster_probe.py
    SUFFIX=_bulk bash tools/experimental/build.sh -DATC_BULK_HANDOFF=1 && timeout 120 python tools/experimental/attn_cluster_probe.py _bulk

For each shape: output of tdnx_attention_tc_cluster against the product tdn_attention_tc (expected bit-identical: same
probabilities, same accumulation order per output element) and against an fp64 reference, then both timed (CUDA events,
20 launches after 3 warm-up).  Wrap in `timeout`: a protocol bug in an unvalidated kernel can hang until the 2 s
mbarrier watchdog of every CTA fires."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tdnet_b200 import _cabi  # noqa: E402

SHAPES = [(1, 300, 100, 512), (2, 1000, 690, 512), (1, 2048, 64, 512), (1, 4096, 2048, 1024), (1, 32768, 2048, 512)]


def split(t):
    hi = t.half()
    return hi.contiguous(), (t - hi.float()).half().contiguous()


def main():
    suffix = sys.argv[1] if len(sys.argv) > 1 else ""        # e.g. "_bulk" for the library built with SUFFIX=_bulk
    lib = C.CDLL(os.path.join(ROOT, "tools", "experimental", "build", f"libtdnet_b200_x{suffix}.so"))
    print(f"library: libtdnet_b200_x{suffix}.so", flush=True)
    variants = [("product", lib.tdn_attention_tc)]
    for name in ("tdnx_attention_tc_cluster", "tdnx_attention_tc_cluster3"):
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = C.c_int, [C.POINTER(_cabi.AttentionDesc), C.c_void_p]
            variants.append((name.replace("tdnx_attention_tc_", ""), fn))
    lib.tdn_attention_tc.restype = C.c_int
    lib.tdn_attention_tc.argtypes = [C.POINTER(_cabi.AttentionDesc), C.c_void_p]
    lib.tdn_last_error.restype = C.c_char_p
    for n, pq, pk, dv in SHAPES:
        g = torch.Generator().manual_seed(pq + pk)
        q, k = torch.randn(n, pq, 64, generator=g) * 1.3, torch.randn(n, pk, 64, generator=g) * 1.4
        v, r = torch.randn(n, pk, dv, generator=g) * 3, torch.randn(n, pq, dv, generator=g)
        pkp = (pk + 63) // 64 * 64
        vt = torch.zeros(n, dv, pkp)
        vt[:, :, :pk] = v.transpose(1, 2)
        pl = {name: split(t.cuda()) for name, t in (("q", q), ("k", k), ("vt", vt), ("r", r))}
        outs = {}
        for which, fn in variants:
            out = torch.full((n, pq, dv), float("nan"), device="cuda")
            d = _cabi.AttentionDesc()
            d.q_hi, d.q_lo, d.q_ld, d.q_batch_stride = pl["q"][0].data_ptr(), pl["q"][1].data_ptr(), 64, pq * 64
            d.k_hi, d.k_lo, d.k_ld, d.k_batch_stride = pl["k"][0].data_ptr(), pl["k"][1].data_ptr(), 64, pk * 64
            d.vt_hi, d.vt_lo, d.vt_ld, d.vt_batch_stride = pl["vt"][0].data_ptr(), pl["vt"][1].data_ptr(), pkp, dv * pkp
            d.out = _cabi.Tensor(out.data_ptr(), None, 0, n, 1, pq, dv, pq * dv, pq * dv, dv)
            d.residual = _cabi.Tensor(pl["r"][0].data_ptr(), pl["r"][1].data_ptr(), 1, n, 1, pq, dv, pq * dv, pq * dv, dv)
            d.n, d.pq, d.pk, d.d_k, d.d_v = n, pq, pk, 64, dv
            rc = fn(C.byref(d), None)
            if rc:
                print(which, "rc", rc, lib.tdn_last_error().decode())
                continue
            torch.cuda.synchronize()
            for _ in range(3):
                fn(C.byref(d), None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn(C.byref(d), None)
            e1.record()
            torch.cuda.synchronize()
            outs[which] = (out, e0.elapsed_time(e1) / 20)
        res = {"shape": [n, pq, pk, dv]}
        a = torch.softmax(torch.bmm(q.double(), k.double().transpose(1, 2)) / 8.0, dim=2) if pq * pk <= 8e7 else None
        ref = torch.bmm(a, v.double()) + r.double() if a is not None else None
        for which, (out, ms) in outs.items():
            res[f"ms_{which}"] = round(ms, 4)
            if which != "product" and "product" in outs:
                res[f"max_diff_{which}"] = float((outs["product"][0] - out).abs().max())
            if ref is not None:
                res[f"max_abs_vs_fp64_{which}"] = float((out.cpu().double() - ref).abs().max())
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
