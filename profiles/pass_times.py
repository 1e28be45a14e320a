#!/usr/bin/env python
"""Device-resident timing of the two passes of one config (CUDA events, 3 warm-ups, median of
N), with a checksum of the output so that scheduling variants can be compared bit for bit.

    python profiles/pass_times.py --cfg cfg3 --var-h 1 --var-v 2 [--n 20]
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import avir_b200 as ab  # noqa: E402

CFG = {  # name: (fpclass, sw, sh, nw, nh, tin, tout, resbits, kwargs)
    "cfg3": (2, 7680, 4320, 3840, 2160, np.float32, np.float32, 16, {}),
    "cfg3f4": (1, 7680, 4320, 3840, 2160, np.float32, np.float32, 16, {}),
    "cfg2": (1, 1920, 1080, 3840, 2160, np.uint8, np.uint8, 8, {}),
    "cfg4": (1, 16384, 16384, 4096, 4096, np.uint16, np.uint16, 16, {}),
    "cfg5": (2, 7680, 4320, 1920, 1080, np.uint8, np.uint8, 8, {"gamma": True, "alpha": 3}),
    "u8k": (1, 7680, 4320, 3840, 2160, np.uint8, np.uint8, 8, {}),
    "u8kdil": (2, 7680, 4320, 3840, 2160, np.uint8, np.uint8, 8, {}),
    "rgb": (0, 5184, 3456, 1920, 1280, np.uint8, np.uint8, 8, {}),   # upstream README case (3 channels)
}
CH = {"rgb": 3}
TT = {np.uint8: torch.uint8, np.uint16: torch.uint16, np.float32: torch.float32}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg3")
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--only", default="both", choices=["both", "row", "col"])
    ap.add_argument("--var-h", type=int, default=-1, help="scheduling variant of the streaming row pass")
    ap.add_argument("--var-v", type=int, default=-1, help="scheduling variant of the streaming column pass")
    ap.add_argument("--family", type=int, default=-1, help="kernel family: 0 product order, 1 generic, 2 tile")
    ap.add_argument("--all-chains", type=int, default=-1)
    a = ap.parse_args()
    fp, sw, sh, nw, nh, ti, to, rb, kw = CFG[a.cfg]
    ch = CH.get(a.cfg, 4)
    lib = ab.lib()
    rs = ab.CImageResizer(rb, 0, 0, fp)
    v = ab.CImageResizerVars(UseSRGBGamma=kw.get("gamma", False), AlphaIndex=kw.get("alpha", -1))
    h, dp, modes = rs.descriptor((sh, sw, ch), ti, nw, nh, to, 0.0, v)
    plan = C.c_void_p()
    assert lib.avirb200_plan_create(C.c_void_p(dp), C.byref(plan)) == 0, lib.avirb200_last_error()
    lib.avirb200_plan_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
    for opt, val in ((ab.OPT_STREAM_VARIANT_H, a.var_h), (ab.OPT_STREAM_VARIANT_V, a.var_v),
                     (ab.OPT_KERNEL_FAMILY, a.family), (ab.OPT_ALL_STREAM_CHAINS, a.all_chains)):
        if val >= 0:
            assert lib.avirb200_plan_set_option(plan, opt, val) == 0
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    if ti == np.float32:
        d_src = torch.rand((sh, sw, ch), generator=g, device="cuda", dtype=torch.float32)
    else:
        hi = 256 if ti == np.uint8 else 65536
        d_src = torch.randint(0, hi, (sh, sw, ch), generator=g, device="cuda", dtype=torch.int32).to(TT[ti])
    d_dst = torch.zeros((nh, nw, ch), device="cuda", dtype=TT[to])
    wsb = C.c_size_t()
    assert lib.avirb200_plan_workspace_bytes(plan, C.byref(wsb)) == 0
    d_ws = torch.empty(wsb.value, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib.avirb200_row_pass_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.avirb200_col_pass_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def row():
        assert lib.avirb200_row_pass_device(plan, d_src.data_ptr(), sw * ch, d_ws.data_ptr(), st) == 0

    def col():
        assert lib.avirb200_col_pass_device(plan, d_ws.data_ptr(), d_dst.data_ptr(), nw * ch, st) == 0

    def med(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.n):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    row()
    rms = med(row) if a.only != "col" else None
    cms = med(col) if a.only != "row" else None
    digest = hashlib.sha1(d_dst.cpu().numpy().tobytes()).hexdigest()[:16]
    paths = lib.avirb200_plan_kernel_paths(plan)
    print(json.dumps({"cfg": a.cfg, "variant_h": a.var_h, "variant_v": a.var_v, "family": a.family,
                      "all_chains": a.all_chains,
                      "kernel_paths": paths, "row_ms": rms, "col_ms": cms,
                      "out_sha1": digest, "build_modes": list(modes)}))
    lib.avirb200_plan_destroy(plan)
    rs.free_descriptor(h)


if __name__ == "__main__":
    main()
