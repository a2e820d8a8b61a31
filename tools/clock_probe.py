"""The SM clock while the big kernels run (tdn_sm_clock_probe: one-warp CTAs that co-reside with the persistent tcgen05 kernels and
compare %clock64 with %globaltimer) next to NVML's reading: idle, under a loop of the layer-4 pair convolution, under a loop of the
big-hop attention kernel, and under running frames.

    timeout 200 python tools/clock_probe.py"""
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tdnet_b200 import _cabi as cabi  # noqa: E402


def nvml_sampler(stop, samples):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop.is_set():
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.002)


def probe(lib, label, work, warm_ms=30.0, span_ms=20.0):
    """work(stream_ms) enqueues >= warm_ms + span_ms of kernels on the current stream; the probe CTAs start after warm_ms."""
    side = torch.cuda.Stream()
    out = torch.zeros(3 * 148, dtype=torch.int64, device="cuda")
    stop, samples = threading.Event(), []
    th = threading.Thread(target=nvml_sampler, args=(stop, samples))
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = work(warm_ms)                      # warm-up part
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        cabi.check(lib.tdn_sm_clock_probe(out.data_ptr(), 148, int(span_ms * 1e6), side.cuda_stream), "probe")
    n += work(span_ms * 1.5)               # runs next to the probe CTAs
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    o = out.view(148, 3).cpu()
    ghz = sorted((o[:, 0].double() / o[:, 1].double()).tolist())
    body = samples[len(samples) // 3:] or [(0, 0.0)]
    res = {"case": label, "sm_clock_mhz_measured": {"median": round(1e3 * statistics.median(ghz), 1), "min": round(1e3 * ghz[0], 1),
                                                     "max": round(1e3 * ghz[-1], 1), "sms": len(set(o[:, 2].tolist()))},
           "nvml_sm_mhz_median": statistics.median(c for c, _ in body), "nvml_power_w_max": round(max(w for _, w in body), 1),
           "launches": n, "ms_per_launch": round(e0.elapsed_time(e1) / max(n, 1), 4)}
    print(json.dumps(res), flush=True)


def main():
    lib = cabi.load()
    g = torch.Generator().manual_seed(1)

    # layer-4 pair convolution
    cin = cout = 512
    h, w = 128, 256
    wt = (torch.randn(cout, 9 * cin, generator=g) / (9 * cin) ** 0.5).cuda()
    wh = wt.half().contiguous(); wl = (wt - wh.float()).half().contiguous()
    x = torch.randn(1, h, w, cin, generator=g).cuda()
    xh = x.half().contiguous(); xl = (x - xh.float()).half().contiguous()
    oh = torch.empty(1, h, w, cout, dtype=torch.half, device="cuda"); ol = torch.empty_like(oh)
    d = cabi.TcConvDesc()
    d.in_ = cabi.Tensor(xh.data_ptr(), xl.data_ptr(), 1, 1, h, w, cin, h * w * cin, w * cin, cin)
    d.out = cabi.Tensor(oh.data_ptr(), ol.data_ptr(), 1, 1, h, w, cout, h * w * cout, w * cout, cout)
    d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), 9 * cin
    d.cout, d.kh, d.kw, d.dilation, d.stride = cout, 3, 3, 4, 0

    def conv_work(ms):
        n = int(ms / 0.33) + 1
        for _ in range(n):
            lib.tdn_conv2d_tc(C.byref(d), None)
        return n

    # big-hop attention
    pq, pk, dv = 32768, 2048, 512
    q, k = torch.randn(1, pq, 64, generator=g) * 1.3, torch.randn(1, pk, 64, generator=g) * 1.4
    v, r = torch.randn(1, pk, dv, generator=g) * 3, torch.randn(1, pq, dv, generator=g)

    def sp(t):
        t = t.cuda(); hi = t.half().contiguous()
        return hi, (t - hi.float()).half().contiguous()
    pl = {name: sp(t) for name, t in (("q", q), ("k", k), ("vt", v.transpose(1, 2).contiguous()), ("r", r))}
    ao = torch.empty(1, pq, dv, device="cuda")
    a = cabi.AttentionDesc()
    a.q_hi, a.q_lo, a.q_ld, a.q_batch_stride = pl["q"][0].data_ptr(), pl["q"][1].data_ptr(), 64, pq * 64
    a.k_hi, a.k_lo, a.k_ld, a.k_batch_stride = pl["k"][0].data_ptr(), pl["k"][1].data_ptr(), 64, pk * 64
    a.vt_hi, a.vt_lo, a.vt_ld, a.vt_batch_stride = pl["vt"][0].data_ptr(), pl["vt"][1].data_ptr(), pk, dv * pk
    a.out = cabi.Tensor(ao.data_ptr(), None, 0, 1, 1, pq, dv, pq * dv, pq * dv, dv)
    a.residual = cabi.Tensor(pl["r"][0].data_ptr(), pl["r"][1].data_ptr(), 1, 1, 1, pq, dv, pq * dv, pq * dv, dv)
    a.n, a.pq, a.pk, a.d_k, a.d_v = 1, pq, pk, 64, dv

    def attn_work(ms):
        n = int(ms / 0.24) + 1
        for _ in range(n):
            lib.tdn_attention_tc(C.byref(a), None)
        return n

    def idle_work(ms):
        return 0

    if "--cublas" in sys.argv:
        # the yardstick: cuBLAS bf16 GEMM 8192^3 (what MEASURED_PEAKS.json's bf16_tflops times), with the clock it runs at
        n8 = 8192
        ma = torch.randn(n8, n8, device="cuda", dtype=torch.bfloat16)
        mb = torch.randn(n8, n8, device="cuda", dtype=torch.bfloat16)
        mc = torch.empty(n8, n8, device="cuda", dtype=torch.bfloat16)
        state = {"n": 0, "t0": None}

        def gemm_work(ms):
            n = int(ms / 0.66) + 1
            for _ in range(n):
                torch.matmul(ma, mb, out=mc)
            return n
        for label, warm in (("cuBLAS bf16 8192^3, back to back", 30.0), ("cuBLAS bf16 8192^3, 300 ms before the probe", 300.0)):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            probe(lib, label, gemm_work, warm_ms=warm)
            e0.record()
            k = gemm_work(20.0)
            e1.record()
            torch.cuda.synchronize()
            print(json.dumps({"case": label + " (the 20 ms right after, no probe CTAs)", "tflops": round(k * 2 * n8 ** 3 / e0.elapsed_time(e1) / 1e9, 1)}), flush=True)
        return
    probe(lib, "idle (probe CTAs only)", idle_work)
    probe(lib, "layer-4 pair convolution, back to back", conv_work)
    probe(lib, "layer-4 pair convolution, 300 ms before the probe", conv_work, warm_ms=300.0)
    probe(lib, "big-hop attention, back to back", attn_work)

    # running frames
    from tdnet_b200.model import td4_psp18
    from tdnet_b200.model.arch import feature_hw
    from tdnet_b200.synth import synth_clip, synth_state_dict
    H, W = 1024, 2048
    net = td4_psp18.td4_psp18(nclass=19, path_num=4, backbone="resnet18", ln_shape=feature_hw(H, W)).eval()
    net.load_state_dict(synth_state_dict(net.state_dict(), seed=0), strict=True)
    net.to("cuda:0")
    frames = [f.cuda() for f in synth_clip(4, H, W)]
    state = {"i": 0}
    with torch.no_grad():
        for _ in range(16):
            net(frames[state["i"] % 4], pos_id=state["i"] % 4); state["i"] += 1
        torch.cuda.synchronize()

        def frame_work(ms):
            n = int(ms / 2.9) + 1
            for _ in range(n):
                net(frames[state["i"] % 4], pos_id=state["i"] % 4); state["i"] += 1
            return n
        probe(lib, "td4-psp18 frames at 1024x2048 (CUDA graphs)", frame_work, warm_ms=60.0, span_ms=30.0)
        probe(lib, "td4-psp18 frames, 1 s before the probe", frame_work, warm_ms=1000.0, span_ms=30.0)


if __name__ == "__main__":
    main()
