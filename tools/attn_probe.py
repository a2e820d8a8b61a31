"""Correctness + timing probe of tdn_attention_tc on a B200: the tensor-memory-operand kernels (tc_attn_ts.cu,
TDNET_ATTN_TS=1, default) against the shared-memory-operand kernels (tc_attn.cu, TDNET_ATTN_TS=0) and fp64.

    timeout 300 python tools/attn_probe.py [out.jsonl]

Each shape: max |diff| between the two kernel families, max |err| of each against an fp64 softmax(q k^T / 8) v + r
(transformer.py:126-139) where the [pq x pk] matrix fits, and CUDA-event timings (20 launches after 3 warm-up; the
big hop additionally 200 back-to-back launches)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tdnet_b200 import _cabi  # noqa: E402

SHAPES = [(1, 300, 100, 512), (2, 1000, 690, 512), (1, 2048, 64, 512), (1, 130, 1225, 128), (3, 257, 130, 256),
          (1, 4096, 2048, 1024), (1, 32768, 1225, 512), (1, 32768, 2048, 512)]


def split(t):
    hi = t.half()
    return hi.contiguous(), (t - hi.float()).half().contiguous()


def formats(lib):
    """Every out / residual format pair on two ragged shapes: TS against SS (bit-identical expected) and fp64."""
    for n, pq, pk, dv in ((2, 1000, 690, 512), (1, 300, 100, 256), (1, 32768, 200, 512)):
        g = torch.Generator().manual_seed(pq + pk)
        q, k = torch.randn(n, pq, 64, generator=g) * 1.3, torch.randn(n, pk, 64, generator=g) * 1.4
        v, r = torch.randn(n, pk, dv, generator=g) * 3, torch.randn(n, pq, dv, generator=g)
        pkp = (pk + 63) // 64 * 64
        vt = torch.zeros(n, dv, pkp)
        vt[:, :, :pk] = v.transpose(1, 2)
        pl = {name: split(t.cuda()) for name, t in (("q", q), ("k", k), ("vt", vt), ("r", r))}
        rf = r.cuda().contiguous()
        a = torch.softmax(torch.bmm(q.double(), k.double().transpose(1, 2)) / 8.0, dim=2)
        base = torch.bmm(a, v.double())
        for out_fmt in ("f32", "split"):
            for res_fmt in ("split", "f32", "none"):
                got = {}
                fams = tuple(os.environ.get("ATTN_PROBE_FAMILIES", "ss,ts").split(","))
                for which in fams:
                    os.environ["TDNET_ATTN_TS"] = {"ss": "0", "ts2": "2", "tq": "3", "s128": "4"}.get(which, "1")
                    of = torch.full((n, pq, dv), float("nan"), device="cuda")
                    oh = torch.full((n, pq, dv), float("nan"), device="cuda", dtype=torch.half)
                    ol = torch.full((n, pq, dv), float("nan"), device="cuda", dtype=torch.half)
                    d = _cabi.AttentionDesc()
                    d.q_hi, d.q_lo, d.q_ld, d.q_batch_stride = pl["q"][0].data_ptr(), pl["q"][1].data_ptr(), 64, pq * 64
                    d.k_hi, d.k_lo, d.k_ld, d.k_batch_stride = pl["k"][0].data_ptr(), pl["k"][1].data_ptr(), 64, pk * 64
                    d.vt_hi, d.vt_lo, d.vt_ld, d.vt_batch_stride = pl["vt"][0].data_ptr(), pl["vt"][1].data_ptr(), pkp, dv * pkp
                    if out_fmt == "f32":
                        d.out = _cabi.Tensor(of.data_ptr(), None, 0, n, 1, pq, dv, pq * dv, pq * dv, dv)
                    else:
                        d.out = _cabi.Tensor(oh.data_ptr(), ol.data_ptr(), 1, n, 1, pq, dv, pq * dv, pq * dv, dv)
                    if res_fmt == "split":
                        d.residual = _cabi.Tensor(pl["r"][0].data_ptr(), pl["r"][1].data_ptr(), 1, n, 1, pq, dv, pq * dv, pq * dv, dv)
                    elif res_fmt == "f32":
                        d.residual = _cabi.Tensor(rf.data_ptr(), None, 0, n, 1, pq, dv, pq * dv, pq * dv, dv)
                    d.n, d.pq, d.pk, d.d_k, d.d_v = n, pq, pk, 64, dv
                    rc = lib.tdn_attention_tc(C.byref(d), None)
                    if rc:
                        print(which, "rc", rc, lib.tdn_last_error().decode(), flush=True)
                        continue
                    torch.cuda.synchronize()
                    got[which] = of if out_fmt == "f32" else oh.float() + ol.float()
                ref = base + (r.double() if res_fmt != "none" else 0)
                last = fams[-1]
                print(json.dumps({"shape": [n, pq, pk, dv], "out": out_fmt, "res": res_fmt,
                                  "max_diff_ts_vs_ss": float((got["ss"] - got[last]).abs().max()),
                                  "nan_ts": int(torch.isnan(got[last]).sum()),
                                  "max_abs_vs_fp64_ts": float((got[last].cpu().double() - ref).abs().max())}), flush=True)


def main():
    # --profile: one launch per kernel family of the big hop only (for ncu)
    profile = "--profile" in sys.argv
    sustain = "--sustain" in sys.argv      # big hop only: ~1.5 s of back-to-back launches with NVML clock / power samples
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out_path = args[0] if args else None
    lib = _cabi.load()
    if "--formats" in sys.argv:
        return formats(lib)
    lines = []
    for n, pq, pk, dv in (SHAPES[-1:] if profile or sustain or "--bighop" in sys.argv else SHAPES):
        g = torch.Generator().manual_seed(pq + pk)
        q, k = torch.randn(n, pq, 64, generator=g) * 1.3, torch.randn(n, pk, 64, generator=g) * 1.4
        v, r = torch.randn(n, pk, dv, generator=g) * 3, torch.randn(n, pq, dv, generator=g)
        pkp = (pk + 63) // 64 * 64
        vt = torch.zeros(n, dv, pkp)
        vt[:, :, :pk] = v.transpose(1, 2)
        pl = {name: split(t.cuda()) for name, t in (("q", q), ("k", k), ("vt", vt), ("r", r))}
        outs = {}
        variants = tuple(os.environ.get("ATTN_PROBE_FAMILIES", "ss,ts").split(","))
        if profile and os.environ.get("ATTN_PROBE_VARIANTS"):
            variants = tuple(os.environ["ATTN_PROBE_VARIANTS"].split(","))
        if sustain and "--debug" in sys.argv:
            variants = tuple(os.environ.get("ATTN_PROBE_VARIANTS", "ts,ts_dbg1,ts_dbg2,ts_dbg4,ts_dbg8,ts_dbg15").split(","))
        for which in variants:
            os.environ["TDNET_ATTN_TS"] = {"ss": "0", "ts2": "2", "tq": "3", "s128": "4"}.get(which.split("_")[0], "1")
            os.environ["TDNET_ATTN_DEBUG"] = which.split("dbg")[1] if "dbg" in which else "0"
            out = torch.full((n, pq, dv), float("nan"), device="cuda")
            d = _cabi.AttentionDesc()
            d.q_hi, d.q_lo, d.q_ld, d.q_batch_stride = pl["q"][0].data_ptr(), pl["q"][1].data_ptr(), 64, pq * 64
            d.k_hi, d.k_lo, d.k_ld, d.k_batch_stride = pl["k"][0].data_ptr(), pl["k"][1].data_ptr(), 64, pk * 64
            d.vt_hi, d.vt_lo, d.vt_ld, d.vt_batch_stride = pl["vt"][0].data_ptr(), pl["vt"][1].data_ptr(), pkp, dv * pkp
            d.out = _cabi.Tensor(out.data_ptr(), None, 0, n, 1, pq, dv, pq * dv, pq * dv, dv)
            d.residual = _cabi.Tensor(pl["r"][0].data_ptr(), pl["r"][1].data_ptr(), 1, n, 1, pq, dv, pq * dv, pq * dv, dv)
            d.n, d.pq, d.pk, d.d_k, d.d_v = n, pq, pk, 64, dv
            rc = lib.tdn_attention_tc(C.byref(d), None)
            if rc:
                print(which, "rc", rc, lib.tdn_last_error().decode(), flush=True)
                continue
            torch.cuda.synchronize()
            if profile:
                outs[which] = (out, [])
                continue
            if sustain:
                import threading
                import time
                import pynvml
                pynvml.nvmlInit()
                h = pynvml.nvmlDeviceGetHandleByIndex(0)
                samples, stop = [], threading.Event()

                def sampler():
                    while not stop.is_set():
                        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
                        time.sleep(0.005)
                nl = 6000
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                th = threading.Thread(target=sampler)
                th.start()
                e0.record()
                for _ in range(nl):
                    lib.tdn_attention_tc(C.byref(d), None)
                e1.record()
                torch.cuda.synchronize()
                stop.set()
                th.join()
                body = samples[len(samples) // 4:]
                clk = sorted(c for c, _ in body)
                outs[which] = (out, [round(e0.elapsed_time(e1) / nl, 4), {"sm_mhz_median": clk[len(clk) // 2],
                               "sm_mhz_min": clk[0], "power_w_mean": round(sum(w for _, w in body) / len(body), 1),
                               "power_w_max": max(w for _, w in body), "samples": len(body)}])
                continue
            reps = [20] + ([200] if pq * pk >= 32768 * 2048 else [])
            ms = []
            for rep in reps:
                for _ in range(3):
                    lib.tdn_attention_tc(C.byref(d), None)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(rep):
                    lib.tdn_attention_tc(C.byref(d), None)
                e1.record()
                torch.cuda.synchronize()
                ms.append(round(e0.elapsed_time(e1) / rep, 4))
            outs[which] = (out, ms)
        res = {"shape": [n, pq, pk, dv]}
        a = torch.softmax(torch.bmm(q.double(), k.double().transpose(1, 2)) / 8.0, dim=2) if n * pq * pk <= 1.4e8 and not profile and not sustain else None
        ref = torch.bmm(a, v.double()) + r.double() if a is not None else None
        for which, (out, ms) in outs.items():
            res[f"ms_{which}"] = ms
            res[f"nan_{which}"] = int(torch.isnan(out).sum())
            if ref is not None:
                res[f"max_abs_vs_fp64_{which}"] = float((out.cpu().double() - ref).abs().max())
        if "ss" in outs and "ts" in outs:
            res["max_diff_ts_vs_ss"] = float((outs["ss"][0] - outs["ts"][0]).abs().max())
        if "ss" in outs and "tq" in outs:
            res["max_diff_tq_vs_ss"] = float((outs["ss"][0] - outs["tq"][0]).abs().max())
        print(json.dumps(res), flush=True)
        lines.append(json.dumps(res))
    if out_path:
        with open(out_path, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
