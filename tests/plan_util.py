"""Host planner access (product: libavirb200_host.so) and comparison with oracle plans."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "avir_b200", "libavirb200_host.so")

_host = None


def host():
    global _host
    if _host is None:
        lib = C.CDLL(HOST_SO)
        lib.avirb200_host_plan_dump.restype = C.c_long
        lib.avirb200_host_plan_dump.argtypes = [C.c_int] * 9 + [C.c_double] * 3 + [C.c_int] * 6 + [
            C.c_void_p, C.c_long]
        _host = lib
    return _host


def host_plan(mirror, sw, sh, nw, nh, ch, in_dtype, out_dtype, k=0.0, resbits=8, srcbits=0,
              ox=0.0, oy=0.0, gamma=False, buildmode=-1, params=0):
    in_dtype, out_dtype = np.dtype(in_dtype), np.dtype(out_dtype)
    args = (mirror, resbits, srcbits, params, sw, sh, nw, nh, ch, k, ox, oy,
            int(in_dtype.kind == "f"), int(out_dtype.kind == "f"), in_dtype.itemsize,
            out_dtype.itemsize, int(gamma), buildmode)
    n = host().avirb200_host_plan_dump(*args, None, 0)
    buf = np.zeros(n, dtype=np.float64)
    assert host().avirb200_host_plan_dump(*args, buf.ctypes.data, n) == n
    pos = [0]

    def take(m=1):
        v = buf[pos[0]:pos[0] + m]
        pos[0] += m
        return v

    plan = dict(zip(["out_mul", "in_gamma_mult", "out_gamma_mult", "el_count"], take(4)))
    for ax in ("H", "V"):
        mode, unsup, ns = [int(v) for v in take(3)]
        steps = []
        for _ in range(ns):
            s = dict(zip(["kind", "R", "lat", "edge", "in_len", "out_len", "ntaps", "order",
                          "upsampled", "skip_odd", "nphases", "out_prefix", "out_suffix",
                          "in_prefix", "in_suffix"], [int(v) for v in take(15)]))
            nt = int(take()[0])
            s["taps"] = take(nt).astype(np.float32)
            npos = int(take()[0])
            p = take(npos * 3).reshape(npos, 3)
            s["src_pos"] = p[:, 0].astype(np.int32)
            s["phase"] = p[:, 1].astype(np.int32)
            s["frac"] = p[:, 2].astype(np.float32)
            steps.append(s)
        plan[ax] = dict(mode=mode, unsupported=bool(unsup), steps=steps)
    assert pos[0] == n
    return plan


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def compare_axis(mine, refsteps):
    """Returns a list of mismatch descriptions between the host plan of one axis and the
    step list recorded from the upstream run."""
    bad = []
    # fold upstream's filterless upsample step into the following resize step
    folded = []
    pend = None
    for s in refsteps:
        if s["kind"] == 1 and s["FltOrigLen"] > 0:
            pend = s  # filterless 2X upsample: folded into the next (resize) step
            continue
        s = dict(s)
        s["up_in_len"] = pend["InLen"] if pend is not None else None
        pend = None
        folded.append(s)
    if len(folded) != len(mine["steps"]):
        return ["step count %d vs ref %d" % (len(mine["steps"]), len(folded))]
    for i, (m, r) in enumerate(zip(mine["steps"], folded)):
        tag = "step %d: " % i
        if r["kind"] == 1:
            if m["kind"] != 1:
                bad.append(tag + "kind")
                continue
            for a, b in (("R", "R"), ("lat", "lat"), ("in_len", "InLen"), ("out_len", "OutLen"),
                         ("out_prefix", "OutPrefix"), ("out_suffix", "OutSuffix"),
                         ("in_prefix", "InPrefix"), ("in_suffix", "InSuffix")):
                if m[a] != r[b]:
                    bad.append(tag + "%s %d vs %d" % (a, m[a], r[b]))
            if len(m["taps"]) != len(r["Flt"]) or not np.array_equal(_bits(m["taps"]), _bits(r["Flt"])):
                bad.append(tag + "upsample taps differ")
        elif r["kind"] == 0:
            if m["kind"] != 0:
                bad.append(tag + "kind")
                continue
            for a, b in (("R", "R"), ("lat", "lat"), ("edge", "edge"), ("in_len", "InLen"),
                         ("out_len", "OutLen")):
                if m[a] != r[b]:
                    bad.append(tag + "%s %d vs %d" % (a, m[a], r[b]))
            if len(m["taps"]) != len(r["Flt"]) or not np.array_equal(_bits(m["taps"]), _bits(r["Flt"])):
                bad.append(tag + "FIR taps differ")
        else:
            if m["kind"] != 2:
                bad.append(tag + "kind")
                continue
            exp_in = r["up_in_len"] if r["up_in_len"] is not None else r["InLen"]
            if m["in_len"] != exp_in:
                bad.append(tag + "in_len %d vs %d" % (m["in_len"], exp_in))
            if m["upsampled"] != int(r["up_in_len"] is not None):
                bad.append(tag + "upsampled flag")
            if m["skip_odd"] != int(r["kind"] == 3):
                bad.append(tag + "skip_odd flag")
            for a, b in (("out_len", "OutLen"), ("ntaps", "FL"), ("order", "order")):
                if m[a] != r[b]:
                    bad.append(tag + "%s %d vs %d" % (a, m[a], r[b]))
            if not np.array_equal(m["src_pos"], r["SrcPosInt"]):
                bad.append(tag + "src_pos differ")
            if not np.array_equal(_bits(m["frac"]), _bits(r["x"])):
                bad.append(tag + "frac differ")
            stride = m["ntaps"] * (m["order"] + 1)
            tp = m["taps"].reshape(-1, stride)
            for j in range(m["out_len"]):
                rt = r["bank"][int(r["fti"][j])]
                if not np.array_equal(_bits(tp[m["phase"][j]]), _bits(rt)):
                    bad.append(tag + "bank taps differ at output %d" % j)
                    break
    return bad
