"""Row-sharded schedule: host-side logic on CPU.

* shard arithmetic from the descriptor (product code, avirb200_shard_query_desc);
* a world_size-2 gloo run in which each rank executes its band with the C port, exchanges
  the halo rows with its neighbour over gloo, and the concatenated result must be
  bit-identical to the unsharded run -- i.e. the halo sizes the product computes are
  sufficient and sharding changes no arithmetic.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

import avir_b200 as ab
import cases as cs


class ShardInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("src_row0", "src_rows", "dst_row0", "dst_rows",
                                          "need_row0", "need_rows", "halo_up", "halo_down")]


def shard_info(dp, rank, nranks):
    si = ShardInfo()
    r = ab.lib().avirb200_shard_query_desc(C.c_void_p(dp), rank, nranks, C.byref(si))
    return r, si


SHARD_CASES = [
    (1, 64, 256, 48, 128, 4, np.float32, np.float32, 16, {"buildmode": 0}),  # cfg3-like
    (2, 64, 256, 48, 128, 4, np.float32, np.float32, 16, {"buildmode": 1}),
    (1, 48, 512, 24, 128, 4, np.uint16, np.uint16, 16, {}),                  # cfg4-like (k=4)
    (1, 40, 96, 80, 192, 4, np.uint8, np.uint8, 8, {"buildmode": 1}),        # cfg2-like (k=0.5)
    (2, 40, 384, 20, 96, 4, np.uint8, np.uint8, 8, {"gamma": True, "alpha": 3, "buildmode": 1}),
    (0, 33, 301, 47, 177, 3, np.uint8, np.uint8, 8, {}),
]


@pytest.mark.parametrize("case", SHARD_CASES, ids=cs.case_id)
@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_shard_partition_is_consistent(case, nranks):
    fp, sw, sh, nw, nh, ch, ti, to, rb, kw = case
    rs, v = cs.resizer_and_vars(case)
    h, dp, _ = rs.descriptor((sh, sw, ch), ti, nw, nh, to, kw.get("k", 0.0), v)
    try:
        infos = []
        for r in range(nranks):
            code, si = shard_info(dp, r, nranks)
            if code != 0:
                pytest.skip("bands too small for %d ranks" % nranks)
            infos.append(si)
        assert infos[0].src_row0 == 0 and infos[0].dst_row0 == 0
        assert infos[0].halo_up == 0 and infos[-1].halo_down == 0
        for a, b in zip(infos, infos[1:]):
            assert a.src_row0 + a.src_rows == b.src_row0
            assert a.dst_row0 + a.dst_rows == b.dst_row0
        last = infos[-1]
        assert last.src_row0 + last.src_rows == sh and last.dst_row0 + last.dst_rows == nh
        for si in infos:
            assert si.need_row0 == si.src_row0 - si.halo_up
            assert si.need_rows == si.halo_up + si.src_rows + si.halo_down
            assert si.need_row0 >= 0 and si.need_row0 + si.need_rows <= sh
    finally:
        rs.free_descriptor(h)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, case_idx, q):
    try:
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port_no)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        case = SHARD_CASES[case_idx]
        fp, sw, sh, nw, nh, ch, ti, to, rb, kw = case
        src = cs.make_input(case)
        rs, v = cs.resizer_and_vars(case)
        h, dp, _ = rs.descriptor((sh, sw, ch), ti, nw, nh, to, kw.get("k", 0.0), v)
        code, si = shard_info(dp, rank, world)
        assert code == 0
        nb = {}
        for r in (rank - 1, rank + 1):
            if 0 <= r < world:
                c2, s2 = shard_info(dp, r, world)
                assert c2 == 0
                nb[r] = s2
        P = cs.port()
        P.avir_port_row_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        P.avir_port_col_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_size_t]
        rowf = nw * ch
        mid = np.full((si.need_rows, rowf), np.nan, np.float32)
        band = np.ascontiguousarray(src[si.src_row0:si.src_row0 + si.src_rows])
        own = mid[si.halo_up:si.halo_up + si.src_rows]
        assert P.avir_port_row_pass(dp, band.ctypes.data, sw * ch, si.src_rows, own.ctypes.data) == 0
        # halo exchange with the neighbours (what NCCL send/recv does on the GPUs)
        ops = []
        keep = []
        if rank > 0:
            send = torch.from_numpy(np.ascontiguousarray(own[:nb[rank - 1].halo_down]))
            recv = torch.empty((si.halo_up, rowf), dtype=torch.float32)
            keep += [send, recv]
            if send.numel():
                ops.append(dist.P2POp(dist.isend, send, rank - 1))
            if recv.numel():
                ops.append(dist.P2POp(dist.irecv, recv, rank - 1))
            up_recv = recv
        if rank + 1 < world:
            n = nb[rank + 1].halo_up
            send = torch.from_numpy(np.ascontiguousarray(own[si.src_rows - n:]))
            recv = torch.empty((si.halo_down, rowf), dtype=torch.float32)
            keep += [send, recv]
            if send.numel():
                ops.append(dist.P2POp(dist.isend, send, rank + 1))
            if recv.numel():
                ops.append(dist.P2POp(dist.irecv, recv, rank + 1))
            down_recv = recv
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if rank > 0 and si.halo_up:
            mid[:si.halo_up] = up_recv.numpy()
        if rank + 1 < world and si.halo_down:
            mid[si.halo_up + si.src_rows:] = down_recv.numpy()
        out = np.zeros((si.dst_rows, nw, ch), to)
        bad = P.avir_port_col_pass(dp, mid.ctypes.data, si.need_row0, si.need_rows, si.dst_row0,
                                   si.dst_row0 + si.dst_rows, out.ctypes.data, nw * ch)
        assert bad == 0, "band lacks rows the outputs depend on"
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
        if rank == 0:
            full = np.concatenate(gathered, axis=0)
            whole, _ = cs.port_output(case, src)
            q.put(cs.count_mismatch(whole, full))
        rs.free_descriptor(h)
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put("rank %d: %s\n%s" % (rank, e, traceback.format_exc()))
        raise


@pytest.mark.parametrize("case_idx", range(len(SHARD_CASES)), ids=lambda i: cs.case_id(SHARD_CASES[i]))
def test_two_rank_gloo_halo_exchange_is_bit_identical(case_idx):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, case_idx, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, q.get() if not q.empty() else "worker failed"
    assert q.get(timeout=5) == 0


def test_band_ranges_fuzz():
    """Seeded random sweep: for random call shapes (every chain type the planner emits) and
    random band counts, the rows avirb200_shard_query_desc says a band needs are sufficient
    (no poisoned row is read) and the banded column pass reproduces the unbanded bits.  The
    multi-GPU schedule and the pipelined host call both rest on this arithmetic."""
    P = cs.port()
    P.avir_port_row_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    P.avir_port_col_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(3)
    types = [np.uint8, np.uint16, np.float32]
    checked = 0
    for it in range(60):
        fp, ch = int(rng.integers(0, 3)), int(rng.integers(1, 5))
        sw, sh = int(rng.integers(4, 40)), int(rng.integers(40, 400))
        mode = int(rng.integers(0, 3))
        if mode == 0:
            nh = max(8, sh // int(rng.integers(1, 7)))
        elif mode == 1:
            nh = sh * int(rng.integers(1, 3))
        else:
            nh = int(rng.integers(8, 500))
        nw = int(rng.integers(2, 50))
        ti, to = types[int(rng.integers(0, 3))], types[int(rng.integers(0, 3))]
        rb = 8 if to == np.uint8 else 16
        kw = {}
        if rng.random() < 0.3:
            kw["gamma"] = True
        if rng.random() < 0.3:
            kw["buildmode"] = int(rng.integers(0, 4))
        if rng.random() < 0.2:
            kw["oy"] = float(rng.uniform(-1, 1))
        case = (fp, sw, sh, nw, nh, ch, ti, to, rb, kw)
        src = cs.make_input(case, seed=300 + it)
        whole, _ = cs.port_output(case, src)
        rs, v = cs.resizer_and_vars(case)
        h, dp, _ = rs.descriptor((sh, sw, ch), ti, nw, nh, to, 0.0, v)
        try:
            nranks = int(rng.integers(2, 10))
            for r in range(nranks):
                code, si = shard_info(dp, r, nranks)
                if code != 0:
                    break  # bands too small for this many ranks: the product refuses, too
                rowf = nw * ch
                band = np.ascontiguousarray(src[si.need_row0:si.need_row0 + si.need_rows])
                mid = np.zeros((si.need_rows, rowf), np.float32)
                assert P.avir_port_row_pass(dp, band.ctypes.data, sw * ch, si.need_rows, mid.ctypes.data) == 0
                out = np.zeros((si.dst_rows, nw, ch), to)
                bad = P.avir_port_col_pass(dp, mid.ctypes.data, si.need_row0, si.need_rows, si.dst_row0,
                                           si.dst_row0 + si.dst_rows, out.ctypes.data, nw * ch)
                assert bad == 0, (cs.case_id(case), nranks, r)
                assert cs.count_mismatch(whole[si.dst_row0:si.dst_row0 + si.dst_rows], out) == 0, \
                    (cs.case_id(case), nranks, r)
                checked += 1
        finally:
            rs.free_descriptor(h)
    assert checked >= 100
