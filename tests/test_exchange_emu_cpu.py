"""The cross-GPU gradient sum (csrc/grad_exchange.cu, include/gvd_exchange.h) executed on the HOST with one PROCESS per
rank: the kernel source as it is (tests/cuda_emu; its system-scope acquire/release flag accesses become C++ atomics on
memory the rank processes share), launched through the C ABI on every rank at once.  Checks what the N = 2 and N = 4
hardware runs could not: world sizes 3, 5 and 8 (the U = 1 unroll), payloads that do not divide by the world size,
several epochs on the same buffers, bit-identical sums on every rank -- and that a missing peer ends in the bounded
wait's status word instead of a hang is left to the hardware test (the limit is 20 s)."""
import ctypes as C
import os
import sys
from multiprocessing import get_context, shared_memory

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAG_BYTES = 256


def _rank(rank, world, names, n_floats, payload, epochs, q):
    for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cuda_emu")):
        sys.path.insert(0, p)
    import gvd_native as n
    import raster_emu

    L = raster_emu.lib()
    L.gvd_exchange_allreduce_sum.argtypes = [C.POINTER(n.ExchangeArgs), C.c_void_p]
    shms = [shared_memory.SharedMemory(name=nm) for nm in names]
    bufs = [np.ndarray((payload + FLAG_BYTES,), dtype=np.uint8, buffer=s.buf) for s in shms]
    mine = bufs[rank][:payload].view(np.float32)
    ok, sums = True, []
    for e in range(1, epochs + 1):
        rng = np.random.default_rng(1000 * e + rank)
        mine[:] = 7.0                                   # beyond n_floats: must stay untouched
        mine[:n_floats] = rng.normal(size=n_floats).astype(np.float32)
        a = n.ExchangeArgs()
        a.world, a.rank = world, rank
        for r in range(world):
            a.bufs[r] = bufs[r].ctypes.data
        a.payload_bytes, a.n_floats, a.epoch = payload, n_floats, e
        rc = L.gvd_exchange_allreduce_sum(C.byref(a), None)
        ok &= rc == 0
        want = np.zeros(n_floats, np.float32)
        for r in range(world):                          # fixed rank order, like the kernel: bit-identical
            want = want + np.random.default_rng(1000 * e + r).normal(size=n_floats).astype(np.float32)
        ok &= bool(np.array_equal(mine[:n_floats], want)) and bool((mine[n_floats:] == 7.0).all())
        sums.append(float(mine[:n_floats].astype(np.float64).sum()))
    timed_out = int(bufs[rank][payload:].view(np.uint32)[17])
    q.put((rank, bool(ok), timed_out, sums))
    for s in shms:
        s.close()


@pytest.mark.parametrize("world,n_floats", [(2, 4 * 1237), (3, 4 * 1001), (5, 4 * 777), (8, 4 * 1237), (8, 4 * 40001), (8, 8)])
def test_peer_memory_allreduce_all_world_sizes(world, n_floats):
    for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cuda_emu")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import raster_emu
    raster_emu.lib()                                    # build the host library HERE, once, not in every rank process at once
    payload = (n_floats + 64) * 4                       # a tail the reduction must not touch
    shms = [shared_memory.SharedMemory(create=True, size=payload + FLAG_BYTES) for _ in range(world)]
    try:
        for s in shms:
            s.buf[:] = bytes(len(s.buf))
        ctx = get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_rank, args=(r, world, [s.name for s in shms], n_floats, payload, 3, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=300) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert all(ok and timed_out == 0 for _, ok, timed_out, _ in res), res
        assert all(r[3] == res[0][3] for r in res)      # every rank holds the same sums, every epoch
    finally:
        for s in shms:
            s.close()
            s.unlink()
