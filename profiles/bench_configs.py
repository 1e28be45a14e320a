#!/usr/bin/env python
"""Device-resident timing of every BASELINE.json config on one B200 (secondary numbers for
DESIGN.md; the headline line is bench.py's).  CUDA events, inputs larger than L2 or L2 flushed,
3 warm-ups, 20 timed resizes per config.  Prints one JSON object per config.

    python profiles/bench_configs.py > gpurun_out/r01_configs.jsonl
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import avir_b200 as ab  # noqa: E402

u8, u16, f32 = np.uint8, np.uint16, np.float32
TT = {u8: torch.uint8, u16: torch.uint16, f32: torch.float32}

CONFIGS = [
    ("cfg2 1920x1080->3840x2160 RGBA u8, float4 mirror", 1, 1920, 1080, 3840, 2160, u8, u8, 8, {}),
    ("cfg3 7680x4320->3840x2160 RGBA f32, float8_dil mirror", 2, 7680, 4320, 3840, 2160, f32, f32, 16, {}),
    ("cfg3 7680x4320->3840x2160 RGBA f32, float4 mirror", 1, 7680, 4320, 3840, 2160, f32, f32, 16, {}),
    ("cfg4 16384x16384->4096x4096 RGBA u16, float4 mirror (one GPU)", 1, 16384, 16384, 4096, 4096, u16, u16, 16, {}),
    ("cfg5 7680x4320->1920x1080 RGBA u8 + sRGB gamma, float8_dil mirror", 2, 7680, 4320, 1920, 1080, u8, u8, 8,
     {"gamma": True, "alpha": 3}),
    ("8K->4K RGBA u8, float4 mirror", 1, 7680, 4320, 3840, 2160, u8, u8, 8, {}),
]


def main():
    peak = 6571.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, fp, sw, sh, nw, nh, ti, to, rb, kw in CONFIGS:
        rs = ab.CImageResizer(rb, 0, 0, fp)
        v = ab.CImageResizerVars(UseSRGBGamma=kw.get("gamma", False), AlphaIndex=kw.get("alpha", -1))
        if ti == f32:
            d_src = torch.rand((sh, sw, 4), device="cuda", dtype=torch.float32)
        else:
            hi = 256 if ti == u8 else 65536
            d_src = torch.randint(0, hi, (sh, sw, 4), device="cuda", dtype=torch.int32).to(TT[ti])
        d_dst = torch.empty((nh, nw, 4), device="cuda", dtype=TT[to])
        ws = rs.workspaceBytes((sh, sw, 4), ti, nw, nh, to, 0.0, v)
        d_ws = torch.empty(ws, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def run():
            rs.resizeImageDevice(d_src.data_ptr(), (sh, sw, 4), ti, d_dst.data_ptr(), nw, nh, to,
                                 d_ws.data_ptr(), 0.0, v, st)
        for _ in range(3):
            run()
        small = (sw * sh * 4 * np.dtype(ti).itemsize) < (200 << 20)
        times = []
        for _ in range(20):
            if small:
                flush.fill_(1)  # inputs smaller than L2: flush it between iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        bytes_ = (sw * sh * 4 * np.dtype(ti).itemsize + 2 * nw * sh * 16 + nw * nh * 4 * np.dtype(to).itemsize)
        print(json.dumps({"config": name, "ms_per_frame": ms, "src_Mpix_per_s": sw * sh / ms / 1e3,
                          "algorithmic_MB": bytes_ / 1e6, "GBps": bytes_ / ms / 1e6,
                          "frac_of_measured_hbm": bytes_ / ms / 1e6 / peak,
                          "l2": "flushed between iterations" if small else "inputs larger than L2"}))
        del d_src, d_dst, d_ws
        torch.cuda.empty_cache()
    # LANCIR 8K->4K RGBA u8 through the host API is PCIe-bound; time the device entry point
    lib = ab.lib()


if __name__ == "__main__":
    main()
