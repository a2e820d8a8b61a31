"""Bring-up probe for the tcgen05 conv/GEMM kernel (run on the B200 box via gpurun).

Each experiment runs in its own subprocess with a timeout so that a trap or a hang in one kernel
variant cannot take the rest of the report (or the box) with it.  Writes gpurun_out/tc_probe.txt.
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def split(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi.contiguous(), lo.contiguous()


def tensor(cabi, hi, lo, n, h, w, c, dtype=1):
    return cabi.Tensor(hi.data_ptr(), lo.data_ptr() if lo is not None else None, dtype, n, h, w, c, h * w * c, w * c, c)


def run_tc(x, wt, dil=1, scale=None, bias=None, residual=None, act=0, out_split=False, bias_along_m=False,
           batched_w=None, reps=0):
    """x: [n,h,w,cin] fp32 cuda; wt: [cout,kh,kw,cin] fp32 cuda (or batched [n,cout,cin]). Returns fp32 out."""
    import torch
    from tdnet_b200 import _cabi as cabi
    lib = cabi.load()
    n, h, w, cin = x.shape
    xh, xl = split(x)
    if batched_w is not None:
        wt = batched_w
        cout, K = wt.shape[1], wt.shape[2]
        kh = kw = 1
    else:
        cout, kh, kw, _ = wt.shape
        K = kh * kw * cin
    wh, wl = split(wt.reshape(-1, K) if batched_w is None else wt)
    d = cabi.TcConvDesc()
    d.in_ = tensor(cabi, xh, xl, n, h, w, cin)
    keep = [xh, xl, wh, wl]
    if out_split:
        oh_ = torch.empty(n, h, w, cout, dtype=torch.half, device="cuda")
        ol_ = torch.empty_like(oh_)
        d.out = tensor(cabi, oh_, ol_, n, h, w, cout)
    else:
        of = torch.empty(n, h, w, cout, dtype=torch.float32, device="cuda")
        d.out = tensor(cabi, of, None, n, h, w, cout, dtype=0)
    if residual is not None:
        rh, rl = split(residual)
        keep += [rh, rl]
        d.residual = tensor(cabi, rh, rl, n, h, w, cout)
    d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), K
    if batched_w is not None:
        d.weight_batched, d.weight_batch_stride = 1, cout * K
    if scale is not None:
        d.scale = scale.data_ptr()
    if bias is not None:
        d.bias = bias.data_ptr()
    d.bias_along_m = int(bias_along_m)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    d.range_flag = flag.data_ptr()
    d.cout, d.kh, d.kw, d.dilation, d.act, d.leaky_slope = cout, kh, kw, dil, act, 0.01
    cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
    torch.cuda.synchronize()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            lib.tdn_conv2d_tc(C.byref(d), None)
        e0.record()
        for _ in range(reps):
            lib.tdn_conv2d_tc(C.byref(d), None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    out = (oh_.float() + ol_.float()) if out_split else of
    return out, ms, int(flag.item())


def ref_conv(x, wt, dil=1):
    import torch
    import torch.nn.functional as F
    xd = x.permute(0, 3, 1, 2).double()
    wd = wt.permute(0, 3, 1, 2).double()
    k = wt.shape[1]
    return F.conv2d(xd, wd, None, 1, dil * (k - 1) // 2, dil).permute(0, 2, 3, 1)


def stats(out, ref):
    d = (out.double() - ref)
    return dict(max_abs=float(d.abs().max()), rel_l2=float(d.norm() / ref.norm()),
                mean_signed_rel=float((d / ref.abs().clamp_min(1e-3)).mean()), ref_absmax=float(ref.abs().max()))


def experiment(name):
    import torch
    torch.manual_seed(0)
    dev = "cuda"
    if name == "layout_debug":
        # one-hot A rows and index-coded B: out[m, n] must equal B[n, m % 64] = n * 64 + m % 64 exactly;
        # any swizzle / descriptor / lane-mapping mistake shows up as a permutation in the dump.
        import numpy as np
        m_idx = torch.arange(256, device=dev)
        x = torch.zeros(1, 1, 256, 64, device=dev)
        x[0, 0, m_idx, m_idx % 64] = 1.0
        w = (torch.arange(128, device=dev)[:, None] * 64 + torch.arange(64, device=dev)[None, :]).float().view(128, 1, 1, 64)
        out, _, _ = run_tc(x, w)
        ref = ref_conv(x, w)
        np.save(os.path.join(ROOT, "gpurun_out", "tc_layout_debug.npy"), out.cpu().numpy())
        bad = (out.double() != ref)
        st = stats(out, ref)
        st["mismatches"] = int(bad.sum())
        st["first_rows"] = out[0, 0, :3, :6].tolist()
        return st
    if name == "gemm_k64":
        x = torch.randn(1, 1, 256, 64, device=dev)
        w = torch.randn(128, 1, 1, 64, device=dev)
        out, _, _ = run_tc(x, w)
        return stats(out, ref_conv(x, w))
    if name == "gemm_k512_ragged":
        x = torch.randn(1, 1, 1000, 512, device=dev)
        w = torch.randn(256, 1, 1, 512, device=dev) / 22
        out, _, _ = run_tc(x, w)
        return stats(out, ref_conv(x, w))
    if name == "gemm_n64":
        x = torch.randn(1, 1, 384, 128, device=dev)
        w = torch.randn(64, 1, 1, 128, device=dev) / 11
        out, _, _ = run_tc(x, w)
        return stats(out, ref_conv(x, w))
    if name == "conv3x3_d2_ragged":
        x = torch.randn(2, 24, 40, 64, device=dev)
        w = torch.randn(128, 3, 3, 64, device=dev) / 24
        out, _, _ = run_tc(x, w, dil=2)
        return stats(out, ref_conv(x, w, 2))
    if name == "conv3x3_d8_97x193":
        x = torch.randn(1, 97, 193, 128, device=dev)
        w = torch.randn(96, 3, 3, 128, device=dev) / 34
        out, _, _ = run_tc(x, w, dil=8)
        return stats(out, ref_conv(x, w, 8))
    if name == "epilogue":
        x = torch.randn(1, 16, 32, 64, device=dev)
        w = torch.randn(128, 3, 3, 64, device=dev) / 24
        s, b = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev)
        r = torch.randn(1, 16, 32, 128, device=dev)
        out, _, flag = run_tc(x, w, scale=s, bias=b, residual=r, act=1, out_split=True)
        ref = torch.relu(ref_conv(x, w) * s.double() + b.double() + r.double())
        st = stats(out, ref)
        st["range_flag"] = flag
        out2, _, _ = run_tc(x, w, bias=torch.randn(512, device=dev), bias_along_m=True)
        return st
    if name == "batched_qk":
        q = torch.randn(2, 1, 300, 64, device=dev)
        k = torch.randn(2, 100, 64, device=dev)
        out, _, _ = run_tc(q, None, batched_w=k)
        ref = torch.bmm(q.view(2, 300, 64).double(), k.double().transpose(1, 2)).view(2, 1, 300, 100)
        return stats(out, ref)
    if name == "layer4_perf":
        x = torch.randn(1, 128, 256, 512, device=dev).relu()
        w = torch.randn(512, 3, 3, 512, device=dev) / 68
        out, ms, _ = run_tc(x, w, dil=4, reps=10)
        st = stats(out, ref_conv(x, w, 4))
        flop = 2 * 128 * 256 * 512 * 512 * 9
        st.update(ms=ms, algorithmic_tflops=flop / ms / 1e9, executed_tflops=3 * flop / ms / 1e9)
        return st
    if name == "wvs_perf":
        # Encoding.w_vs / Attention.fc shape: 1x1 512 -> 512 on the 128x256 map (8 K blocks per tile)
        x = torch.randn(1, 128, 256, 512, device=dev)
        w = torch.randn(512, 1, 1, 512, device=dev) / 22
        out, ms, _ = run_tc(x, w, reps=10)
        st = stats(out, ref_conv(x, w))
        st.update(ms=ms, algorithmic_tflops=2 * 128 * 256 * 512 * 512 / ms / 1e9)
        return st
    if name == "head_perf":
        # FCNHead conv5.0: 3x3 512 -> 128 on the 128x256 map
        x = torch.randn(1, 128, 256, 512, device=dev)
        w = torch.randn(128, 3, 3, 512, device=dev) / 68
        out, ms, _ = run_tc(x, w, reps=10)
        st = stats(out, ref_conv(x, w))
        st.update(ms=ms, algorithmic_tflops=2 * 128 * 256 * 512 * 128 * 9 / ms / 1e9)
        return st
    if name == "layer4_sustained":
        # same conv, 300 launches back to back over 4 rotating input/output sets (268 MB each > L2 share):
        # separates the power/clock and L2-residency effects from the single-launch number
        from tdnet_b200 import _cabi as cabi
        lib = cabi.load()
        w = torch.randn(512, 3, 3, 512, device=dev) / 68
        wh, wl = split(w.reshape(512, -1))
        descs, keep = [], []
        for i in range(4):
            x = torch.randn(1, 128, 256, 512, device=dev).relu()
            xh, xl = split(x)
            oh_, ol_ = torch.empty_like(xh), torch.empty_like(xh)
            d = cabi.TcConvDesc()
            d.in_ = tensor(cabi, xh, xl, 1, 128, 256, 512)
            d.out = tensor(cabi, oh_, ol_, 1, 128, 256, 512)
            d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), 4608
            d.cout, d.kh, d.kw, d.dilation = 512, 3, 3, 4
            descs.append(d); keep += [xh, xl, oh_, ol_]
        for d in descs:
            cabi.check(lib.tdn_conv2d_tc(C.byref(d), None))
        torch.cuda.synchronize()
        res = {}
        for reps in (8, 300):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(reps):
                lib.tdn_conv2d_tc(C.byref(descs[r % 4]), None)
            e1.record()
            torch.cuda.synchronize()
            res[f"ms_per_launch_{reps}"] = e0.elapsed_time(e1) / reps
        return res
    if name in ("stem_tc_perf", "stem_tc_small"):
        from tdnet_b200.engine import pack_stem_tc, View
        from tdnet_b200 import _cabi as cabi
        import torch.nn.functional as F
        lib = cabi.load()
        h, w = (1024, 2048) if name == "stem_tc_perf" else (70, 300)
        img = torch.randn(1, 3, h, w, device=dev)
        wt = torch.randn(64, 3, 7, 7, device=dev) / 12
        sc, bi = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.3
        ref = F.max_pool2d(F.relu(F.conv2d(img.double(), wt.double(), None, 2, 3) * sc.double().view(1, -1, 1, 1)
                                  + bi.double().view(1, -1, 1, 1)), 3, 2, 1).permute(0, 2, 3, 1)
        hp, wp = ref.shape[1:3]
        wk, inv = pack_stem_tc(wt)
        scd = (sc * inv).contiguous()
        out = View.alloc(1, hp, wp, 64, dev, split=True)
        t = out.ct()
        def run():
            cabi.check(lib.tdn_stem_conv_pool_tc(img.data_ptr(), None, None, 1, h, w, wk.data_ptr(), scd.data_ptr(),
                                                 bi.data_ptr(), C.byref(t), None, None), "stem_tc")
        run(); torch.cuda.synchronize()
        st = stats(out.torch(), ref)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        st["ms"] = e0.elapsed_time(e1) / 10
        # the fp32 CUDA-core stem on the same input
        out2 = View.alloc(1, hp, wp, 64, dev, split=True)
        t2 = out2.ct()
        wk2 = wt.permute(1, 2, 3, 0).reshape(147, 64).contiguous()
        def run2():
            cabi.check(lib.tdn_stem_conv_pool(img.data_ptr(), 1, h, w, wk2.data_ptr(), sc.data_ptr(), bi.data_ptr(),
                                              C.byref(t2), None), "stem")
        run2(); torch.cuda.synchronize()
        e0.record()
        for _ in range(10): run2()
        e1.record(); torch.cuda.synchronize()
        st["simt_ms"] = e0.elapsed_time(e1) / 10
        st["simt_max_abs"] = float((out2.torch().double() - ref).abs().max())
        return st
    if name == "halo_ragged":
        x = torch.randn(2, 23, 30, 64, device=dev)
        w = torch.randn(40, 3, 3, 64, device=dev) / 24
        out, _, _ = run_tc(x, w, dil=2)
        return stats(out, ref_conv(x, w, 2))
    if name == "halo_epilogue":
        x = torch.randn(1, 40, 24, 128, device=dev)
        w = torch.randn(128, 3, 3, 128, device=dev) / 34
        s_, b_ = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev)
        r = torch.randn(1, 40, 24, 128, device=dev)
        out, _, flag = run_tc(x, w, scale=s_, bias=b_, residual=r, act=1, out_split=True)
        ref = torch.relu(ref_conv(x, w) * s_.double() + b_.double() + r.double())
        return stats(out, ref)
    if name == "layer2_perf":
        x = torch.randn(1, 128, 256, 128, device=dev).relu()
        w = torch.randn(128, 3, 3, 128, device=dev) / 34
        out, ms, _ = run_tc(x, w, dil=1, reps=10)
        st = stats(out, ref_conv(x, w, 1))
        st.update(ms=ms, algorithmic_tflops=2 * 128 * 256 * 128 * 128 * 9 / ms / 1e9)
        return st
    if name == "layer3_perf":
        x = torch.randn(1, 128, 256, 256, device=dev).relu()
        w = torch.randn(256, 3, 3, 256, device=dev) / 48
        out, ms, _ = run_tc(x, w, dil=2, reps=10)
        st = stats(out, ref_conv(x, w, 2))
        st.update(ms=ms, algorithmic_tflops=2 * 128 * 256 * 256 * 256 * 9 / ms / 1e9)
        return st
    if name == "layer1_perf":
        x = torch.randn(1, 256, 512, 64, device=dev).relu()
        w = torch.randn(64, 3, 3, 64, device=dev) / 24
        out, ms, _ = run_tc(x, w, dil=1, reps=10)
        st = stats(out, ref_conv(x, w, 1))
        flop = 2 * 256 * 512 * 64 * 64 * 9
        st.update(ms=ms, algorithmic_tflops=flop / ms / 1e9)
        return st
    if name == "accum_bias_positive":
        # all-positive operands: truncating accumulation would show up as a negative mean signed error
        x = torch.rand(1, 1, 2048, 4608, device=dev) + 0.5
        w = torch.rand(128, 1, 1, 4608, device=dev) + 0.5
        out, _, _ = run_tc(x, w)
        xs = (x.half().float() + (x - x.half().float()).half().float()).double()
        ws = (w.half().float() + (w - w.half().float()).half().float()).double()
        ref_split = torch.matmul(xs.view(2048, 4608), ws.view(128, 4608).t()).view(1, 1, 2048, 128)
        ref_true = torch.matmul(x.double().view(2048, 4608), w.double().view(128, 4608).t()).view(1, 1, 2048, 128)
        torch.backends.cuda.matmul.allow_tf32 = False
        f32 = torch.matmul(x.view(2048, 4608), w.view(128, 4608).t()).view(1, 1, 2048, 128)
        return dict(tc_vs_split_inputs=stats(out, ref_split), tc_vs_true=stats(out, ref_true),
                    cublas_fp32_vs_true=stats(f32, ref_true))
    if name.startswith("attention"):
        return attention_experiment(name)
    raise KeyError(name)


def attention_experiment(name):
    """Fused attention kernel vs an fp64 softmax(QK^T/8)V + R reference."""
    import torch
    from tdnet_b200 import _cabi as cabi
    lib = cabi.load()
    dev = "cuda"
    torch.manual_seed(1)
    n, pq, pk, dv, reps = {"attention_small": (2, 300, 100, 128, 0), "attention_ragged": (1, 18721, 1225, 512, 3),
                           "attention_big": (1, 32768, 2048, 512, 10)}[name]
    pkp = (pk + 63) // 64 * 64
    q = torch.randn(n, pq, 64, device=dev) * 1.3
    k = torch.randn(n, pk, 64, device=dev) * 1.4
    v = torch.randn(n, pk, dv, device=dev) * 3
    r = torch.randn(n, pq, dv, device=dev)
    qh, ql = split(q)
    kh, kl = split(k)
    vt = torch.zeros(n, dv, pkp, device=dev)
    vt[:, :, :pk] = v.transpose(1, 2)
    vh, vl = split(vt)
    rh, rl = split(r)
    out = torch.empty(n, pq, dv, device=dev)
    d = cabi.AttentionDesc()
    d.q_hi, d.q_lo, d.q_ld, d.q_batch_stride = qh.data_ptr(), ql.data_ptr(), 64, pq * 64
    d.k_hi, d.k_lo, d.k_ld, d.k_batch_stride = kh.data_ptr(), kl.data_ptr(), 64, pk * 64
    d.vt_hi, d.vt_lo, d.vt_ld, d.vt_batch_stride = vh.data_ptr(), vl.data_ptr(), pkp, dv * pkp
    d.out = cabi.Tensor(out.data_ptr(), None, 0, n, 1, pq, dv, pq * dv, pq * dv, dv)
    d.residual = cabi.Tensor(rh.data_ptr(), rl.data_ptr(), 1, n, 1, pq, dv, pq * dv, pq * dv, dv)
    d.n, d.pq, d.pk, d.d_k, d.d_v = n, pq, pk, 64, dv
    cabi.check(lib.tdn_attention_tc(C.byref(d), None), "attention_tc")
    torch.cuda.synchronize()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            lib.tdn_attention_tc(C.byref(d), None)
        e0.record()
        for _ in range(reps):
            lib.tdn_attention_tc(C.byref(d), None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    # reference in fp64, chunked over queries to bound memory
    ref = torch.empty(n, pq, dv, dtype=torch.float64, device=dev)
    for s0 in range(0, pq, 4096):
        a = torch.softmax(torch.bmm(q[:, s0:s0 + 4096].double(), k.double().transpose(1, 2)) / 8.0, dim=2)
        ref[:, s0:s0 + 4096] = torch.bmm(a, v.double()) + r[:, s0:s0 + 4096].double()
    st = stats(out, ref)
    if ms:
        flop = 2.0 * n * pq * pk * (64 + dv)
        st.update(ms=ms, algorithmic_tflops=flop / ms / 1e9)
    return st


EXPERIMENTS = ["stem_tc_small", "stem_tc_perf"]
# (experiment, TDNET_TC_CHUNK_KB) pairs run after the default set
CHUNK_SWEEP = []
# (experiment, extra environment) pairs: the halo-region variant of the 3x3 kernel against the per-tap one
ENV_SWEEP = []   # e.g. = _HALO_SWEEP
_HALO_SWEEP = [("halo_ragged", {"TDNET_TC_HALO": "1"}), ("halo_epilogue", {"TDNET_TC_HALO": "1"}),
             ("conv3x3_d2_ragged", {"TDNET_TC_HALO": "1"}), ("layer1_perf", {"TDNET_TC_HALO": "0"}),
             ("layer1_perf", {"TDNET_TC_HALO": "1"}), ("layer2_perf", {"TDNET_TC_HALO": "0"}),
             ("layer2_perf", {"TDNET_TC_HALO": "1"}), ("layer3_perf", {"TDNET_TC_HALO": "0"}),
             ("layer3_perf", {"TDNET_TC_HALO": "1"})]

_PAIR_SWEEP = [(n, {"TDNET_TC_PAIR": v}) for n in ("layer4_perf", "layer3_perf", "wvs_perf")
               for v in ("0", "1")] + [("head_perf", {"TDNET_TC_PAIR": "0"}), ("head_perf", {"TDNET_TC_PAIR": "2"}),
                                       ("layer2_perf", {"TDNET_TC_PAIR": "2", "TDNET_TC_HALO": "0"})]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        print("RESULT " + json.dumps(experiment(sys.argv[2])))
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    lines = []
    sweep = _HALO_SWEEP if os.environ.get("TDNET_PROBE_HALO_SWEEP") else ENV_SWEEP
    if os.environ.get("TDNET_PROBE_PAIR_SWEEP"):
        sweep = _PAIR_SWEEP
    todo = [(n, None) for n in (sys.argv[1:] or EXPERIMENTS)] + ([] if sys.argv[1:] else CHUNK_SWEEP + sweep)
    if os.environ.get("TDNET_PROBE_PAIR_SWEEP"):
        todo = list(_PAIR_SWEEP)
    for name, chunk in todo:
        t0 = time.time()
        env = dict(os.environ)
        if isinstance(chunk, dict):
            env.update(chunk)
        elif chunk is not None:
            env["TDNET_TC_CHUNK_KB"] = str(chunk)
        try:
            r = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=150,
                               env=env)
            res = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            msg = res[0][7:] if res else f"FAILED rc={r.returncode} stdout={r.stdout[-600:]!r} stderr={r.stderr[-900:]!r}"
        except subprocess.TimeoutExpired:
            msg = "TIMEOUT (150 s)"
        tag = "" if chunk is None else (f" {chunk}" if isinstance(chunk, dict) else f" [chunk_kb={chunk}]")
        line = f"{name}{tag}: {msg}  [{time.time() - t0:.1f}s]"
        print(line, flush=True)
        lines.append(line)
    open(os.path.join(ROOT, "gpurun_out", "tc_probe.txt"), "w").write("\n".join(lines) + "\n")
