#!/usr/bin/env python
"""bench.py -- headline benchmark: Mpixels/s of the AVIR separable resize hot path, 8K->4K RGBA.

  python bench.py --gpus N --steps K --warmup W            (own arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (upstream's CPU path, rank 0 only)

Workload (BASELINE.json configs[2]): CImageResizer<fpclass_float8_avx-equiv> 7680x4320 ->
3840x2160, 4-channel float.  A "step" resizes one such frame per GPU.  For N > 1 the N frames
form one 7680 x (4320*N) image that is ROW-SHARDED over the ranks: every rank runs the row
pass on its band, exchanges the filter-length halo rows with its neighbours over NCCL
(NVLink) and runs the column pass on its band -- per-GPU work is fixed (weak scaling) and the
real exchange step of the path is inside the timed region.

One JSON line is printed by rank 0; see DESIGN.md section "Measurement" for every field.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SRC_W, SRC_H, DST_W, DST_H, CH = 7680, 4320, 3840, 2160, 4
METRIC = "Mpixels/sec resize (8K->4K RGBA)"
MIRRORS = {"dil": 2, "f4": 1, "def": 0}


def algorithmic_bytes(n_frames=1):
    """src read + intermediate write + intermediate read + dst write."""
    src = SRC_W * SRC_H * CH * 4
    mid = DST_W * SRC_H * CH * 4
    dst = DST_W * DST_H * CH * 4
    return dict(row=(src + mid) * n_frames, col=(mid + dst) * n_frames,
                total=(src + 2 * mid + dst) * n_frames)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or \
               [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                              ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": sorted(reasons)}


def traffic_from_profiles():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


# ------------------------------------------------------------------------------------------------

def pick_threads(o, src, fp, cores):
    """Upstream's thread pool does not scale to every host core (each call allocates and
    first-touches its ~400 MB of scratch from all threads at once); use the thread count
    that is actually fastest on this host."""
    best, best_t = cores, None
    cand = sorted({c for c in (cores, cores // 2, cores // 4, 32, 16, 8, 4) if 1 <= c <= cores}, reverse=True)
    o.ref_resize(src, DST_W, DST_H, np.float32, fpclass=fp, resbits=16, nthreads=cand[0])
    for c in cand:
        t0 = time.perf_counter()
        o.ref_resize(src, DST_W, DST_H, np.float32, fpclass=fp, resbits=16, nthreads=c)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    return best


def run_reference(args):
    """Upstream's own CPU implementation (oracle/_ref: the unmodified headers compiled with the
    pinned flags) on all host threads, one full frame per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_ref as o
    if not o.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libavir_ref.so missing"}))
        return
    fp = MIRRORS[args.mirror]
    src = o.lcg_image(SRC_H, SRC_W, CH, np.float32, seed=12345)
    cores = pick_threads(o, src, fp, os.cpu_count() or 1)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o.ref_resize(src, DST_W, DST_H, np.float32, fpclass=fp, resbits=16, nthreads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    ms = total / len(times) * 1e3
    val = SRC_W * SRC_H * len(times) / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "reference",
                         "sample": "full 7680x4320 frame per step, %d steps, std::thread pool of %d "
                                   "workloads (fastest of a sweep up to %d host threads), pinned "
                                   "flags -O2 -mavx2 -ffp-contract=off"
                                   % (len(times), cores, os.cpu_count() or 1)},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n):
    return {"workload": "cfg3: CImageResizer<%s>(16) 7680x%d->3840x%d RGBA float32, k=2"
                        % ({"dil": "fpclass_float8_dil", "f4": "fpclass_float4",
                            "def": "fpclass_def<float>"}[args.mirror], SRC_H * n, DST_H * n),
            "mirror": args.mirror, "frames_per_step": n,
            "parallelism": "single GPU" if n == 1 else "row-sharded x%d, NCCL halo exchange" % n,
            "l2": "inputs larger than L2 (531 MB source + 265 MB intermediate per GPU per step)"}


def run_own(args):
    import torch
    import torch.distributed as dist
    import avir_b200 as ab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = ab.lib()
    hl = ab.host_lib()
    fp = MIRRORS[args.mirror]
    N = world
    stream = torch.cuda.current_stream().cuda_stream

    # ---- plan for the global image (N stacked frames)
    rs = ab.CImageResizer(16, 0, 0, fp)
    shape = (SRC_H * N, SRC_W, CH)
    h, dp, modes = rs.descriptor(shape, np.float32, DST_W, DST_H * N, np.float32, 0.0)
    plan = C.c_void_p()
    r = lib.avirb200_plan_create(C.c_void_p(dp), C.byref(plan))
    if r != 0:
        raise SystemExit("plan_create: " + lib.avirb200_last_error().decode())

    class SI(C.Structure):
        _fields_ = [(n_, C.c_int32) for n_ in ("src_row0", "src_rows", "dst_row0", "dst_rows",
                                               "need_row0", "need_rows", "halo_up", "halo_down")]
    si = SI()
    assert lib.avirb200_shard_query(plan, rank, N, C.byref(si)) == 0
    wsb = C.c_size_t()
    assert lib.avirb200_shard_workspace_bytes(plan, rank, N, C.byref(wsb)) == 0

    comm = C.c_void_p()
    if N > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_char * 128)()
            assert lib.avirb200_comm_unique_id(raw) == 0, lib.avirb200_last_error().decode()
            idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        idg = idbuf.cuda()
        dist.broadcast(idg, 0)
        raw = (C.c_char * 128).from_buffer_copy(bytes(idg.cpu().numpy().tobytes()))
        assert lib.avirb200_comm_create(raw, rank, N, C.byref(comm)) == 0, lib.avirb200_last_error().decode()

    g = torch.Generator(device="cuda")
    g.manual_seed(12345 + rank)
    d_src = torch.rand((si.src_rows, SRC_W, CH), generator=g, device="cuda", dtype=torch.float32)
    d_dst = torch.empty((si.dst_rows, DST_W, CH), device="cuda", dtype=torch.float32)
    d_ws = torch.empty(wsb.value, dtype=torch.uint8, device="cuda")
    lib.avirb200_resize_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                            C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.avirb200_resize_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_void_p]
    lib.avirb200_row_pass_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.avirb200_col_pass_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]

    def step():
        if N == 1:
            rr = lib.avirb200_resize_device(plan, d_src.data_ptr(), SRC_W * CH, d_dst.data_ptr(),
                                            DST_W * CH, d_ws.data_ptr(), stream)
        else:
            rr = lib.avirb200_resize_sharded(plan, comm, rank, N, d_src.data_ptr(), SRC_W * CH,
                                             d_dst.data_ptr(), DST_W * CH, d_ws.data_ptr(), stream)
        if rr != 0:
            raise SystemExit("resize failed: " + lib.avirb200_last_error().decode())

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if N > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t0 = time.time()
    total_ms = timed(step, args.steps, args.warmup)
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    launches_per_step = lib.avirb200_plan_last_launches(plan)
    ms_per_step = total_ms / args.steps
    value = SRC_W * SRC_H * N / (ms_per_step * 1e-3) / 1e6

    # ---- per-kernel timing for the roofline (single-GPU geometry; kernels are per-rank)
    roof = None
    if N == 1:
        row_ms = timed(lambda: lib.avirb200_row_pass_device(plan, d_src.data_ptr(), SRC_W * CH,
                                                           d_ws.data_ptr(), stream), args.steps, 2) / args.steps
        col_ms = timed(lambda: lib.avirb200_col_pass_device(plan, d_ws.data_ptr(), d_dst.data_ptr(),
                                                           DST_W * CH, stream), args.steps, 2) / args.steps
        ab_ = algorithmic_bytes()
        peak, how = peaks()
        dom = "row" if row_ms >= col_ms else "col"
        dom_ms = max(row_ms, col_ms)
        ach = ab_[dom] / (dom_ms * 1e-3) / 1e9
        tr = traffic_from_profiles()
        roof = {"bound": "hbm", "kernel": dom + "_pass", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": tr.get(dom + "_pass"), "peak_source": how,
                "algorithmic_bytes_per_launch": ab_[dom],
                "kernels": {"row_pass": {"ms": row_ms, "GBps": ab_["row"] / row_ms / 1e6,
                                         "frac": ab_["row"] / row_ms / 1e6 / peak},
                            "col_pass": {"ms": col_ms, "GBps": ab_["col"] / col_ms / 1e6,
                                         "frac": ab_["col"] / col_ms / 1e6 / peak}},
                "whole_step": {"GBps": ab_["total"] / ms_per_step / 1e6,
                               "frac": ab_["total"] / ms_per_step / 1e6 / peak}}

    # ---- end to end through the public API, host buffers, copies inside the timed region
    h_src = torch.empty((si.src_rows, SRC_W, CH), dtype=torch.float32).pin_memory()
    h_src.copy_(d_src)
    h_dst = torch.empty((si.dst_rows, DST_W, CH), dtype=torch.float32).pin_memory()
    if N == 1:
        src_np, dst_np = h_src.numpy(), h_dst.numpy()

        def e2e_step():
            rs.resizeImage(src_np, DST_W, DST_H, 0.0, NewBuf=dst_np)  # H2D + passes + D2H + sync
    else:
        def e2e_step():
            d_src.copy_(h_src, non_blocking=True)
            step()
            h_dst.copy_(d_dst, non_blocking=True)
            torch.cuda.current_stream().synchronize()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    tw0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - tw0], device="cuda")
    if N > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = SRC_W * SRC_H * N * e2e_steps / float(e2e_s.item()) / 1e6

    # ---- CPU baseline: upstream itself on the host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        import oracle_ref as o
        if o.have_ref():
            src = h_src.numpy()
            cores = pick_threads(o, src, fp, os.cpu_count() or 1)
            ts = []
            tb = time.perf_counter()
            while len(ts) < 3 or (time.perf_counter() - tb < 12 and len(ts) < 10):
                t_ = time.perf_counter()
                o.ref_resize(src, DST_W, DST_H, np.float32, fpclass=fp, resbits=16, nthreads=cores)
                ts.append(time.perf_counter() - t_)
            med = sorted(ts)[len(ts) // 2]
            cpu = {"value": SRC_W * SRC_H / med / 1e6, "unit": "Mpix/s", "cores": cores,
                   "kind": "reference", "ms_per_frame": med * 1e3,
                   "sample": "%d full 7680x4320 frames, upstream headers (-O2 -mavx2 "
                             "-ffp-contract=off) on a std::thread pool of %d workloads (fastest of a "
                             "sweep up to %d host threads)" % (len(ts), cores, os.cpu_count() or 1)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, N), build_modes=list(modes),
                           halo_rows=[si.halo_up, si.halo_down]),
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "Mpix/s",
                    "h2d_bytes_per_step": int(si.src_rows) * SRC_W * CH * 4 * N,
                    "d2h_bytes_per_step": int(si.dst_rows) * DST_W * CH * 4 * N,
                    "steps": e2e_steps},
            "gpu_launches": int(launches_per_step) * args.steps, "clocks": clocks,
        }
        print(json.dumps(line))
    if N > 1:
        lib.avirb200_comm_destroy(comm)
        dist.destroy_process_group()
    lib.avirb200_plan_destroy(plan)
    rs.free_descriptor(h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--mirror", default="dil", choices=sorted(MIRRORS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
