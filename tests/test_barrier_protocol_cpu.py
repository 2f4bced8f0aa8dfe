"""Model check of the mbarrier / tcgen05.commit protocols of the tensor-core kernels (csrc/attn_tc.cu generation 7,
csrc/attn_bwd_tc.cu both forms, csrc/gemm_tc.cu one-CTA persistent and CTA-pair kernels) under random schedules.

What this is: a hand transcription of each kernel's SYNCHRONISATION -- which agent waits on which barrier with which
parity expression, which MMAs it issues, when it commits -- run under a randomised scheduler in which the tensor pipe and
the TMA engine are agents of their own (MMAs execute in issue order but arbitrarily late; a commit arrives when everything
issued before it has executed).  What it checks on every schedule:
  * no deadlock;
  * every read sees the version it was written for: S / dP of sub-block i when softmax(i) loads them, P / dS of sub-block i
    when the accumulating MMAs of sub-block i execute, the operand tile j when the MMAs of tile j execute, and -- the one
    that motivated the file -- the accumulator complete when the epilogue reads it.
mbarrier semantics modelled: a barrier has a phase counter and an arrival count per phase; `try_wait.parity P` succeeds iff
the phase of parity P has completed, i.e. iff the CURRENT phase's parity differs from P.  A parity wait is therefore only
meaningful while the waiter is at most one phase behind the barrier; two phases behind it passes on stale state, two
phases ahead it never passes.  The forward's epilogue had exactly that: it waited for the last PV on a barrier that
completes once per sub-block (`first_epilogue=True` below reproduces it and the checker finds the premature read); it now
waits on a single-use barrier.  It is a model, not the CUDA source: the parity expressions and the issue order are copied
by hand and must be kept in step with the kernels (each agent cites its lines).
"""
import random

import pytest


class Barrier:
    def __init__(self, name, count=1):
        self.name, self.count, self.arrived, self.phase = name, count, 0, 0

    def arrive(self):
        self.arrived += 1
        if self.arrived == self.count:
            self.arrived, self.phase = 0, self.phase + 1

    def passed(self, parity):
        return (self.phase & 1) != parity


class Violation(Exception):
    pass


class Machine:
    """Agents are generators yielding ('wait', barrier, parity) | ('arrive', barrier) | ('mma', fn) | ('commit', barrier) |
    ('tma', fn, barrier) | ('do', fn).  'mma' / 'commit' go through the in-order tensor pipe, 'tma' completes out of line."""

    def __init__(self, seed, lag):
        self.rng = random.Random(seed)
        self.lag = lag            # how reluctant the tensor pipe / TMA engine are to make progress (0 .. 0.95)
        self.pipe, self.tma = [], []
        self.agents = {}

    def add(self, name, gen):
        self.agents[name] = [gen, None]  # generator, pending wait

    def run(self, max_steps=200000):
        for _ in range(max_steps):
            runnable = []
            for name, (gen, pending) in self.agents.items():
                if gen is None:
                    continue
                if pending is None or pending[0].passed(pending[1]):
                    runnable.append(name)
            engines = []
            if self.pipe and self.rng.random() > self.lag:
                engines.append("<pipe>")
            if self.tma and self.rng.random() > self.lag:
                engines.append("<tma>")
            if not runnable and not self.pipe and not self.tma:
                if all(g is None for g, _ in self.agents.values()):
                    return
                stuck = {n: (p[0].name, p[1], p[0].phase) for n, (g, p) in self.agents.items() if g is not None}
                raise Violation(f"deadlock: {stuck}")
            choice = self.rng.choice(runnable + engines) if (runnable or engines) else None
            if choice is None:
                continue  # engines exist but chose to lag this tick
            if choice == "<pipe>":
                kind, arg = self.pipe.pop(0)
                if kind == "mma":
                    arg()
                else:  # a commit; a list is a multicast commit (cta_group::2: the same barrier in both CTAs of the pair)
                    for bar in (arg if isinstance(arg, list) else [arg]):
                        bar.arrive()
            elif choice == "<tma>":
                fn, bar = self.tma.pop(self.rng.randrange(len(self.tma)))
                fn()
                bar.arrive()
            else:
                slot = self.agents[choice]
                slot[1] = None
                try:
                    op = next(slot[0])
                except StopIteration:
                    slot[0] = None
                    continue
                if op[0] == "wait":
                    slot[1] = (op[1], op[2])
                elif op[0] == "arrive":
                    op[1].arrive()
                elif op[0] == "mma":
                    self.pipe.append(("mma", op[1]))
                elif op[0] == "commit":
                    self.pipe.append(("commit", op[1]))
                elif op[0] == "tma":
                    self.tma.append((op[1], op[2]))
                elif op[0] == "do":
                    op[1]()
        raise Violation("no termination")


def need(cond, msg):
    if not cond:
        raise Violation(msg)


# ------------------------------------------------------------------------------------------------------------------
# forward, generation 7 (attn_tc.cu::flash_attn7_kernel): 64-key sub-blocks, S / P double-buffered, QK one ahead
# ------------------------------------------------------------------------------------------------------------------
def forward_gen7(nsub, seed, lag, rescale_prob=0.3, first_epilogue=False, two_waits=False, softmax_warps=4):
    m = Machine(seed, lag)
    rng = random.Random(seed * 7 + 1)
    nblk = (nsub + 1) // 2
    q_full = Barrier("q_full")
    kv_full, kv_empty = [Barrier(f"kv_full{s}") for s in range(2)], [Barrier(f"kv_empty{s}") for s in range(2)]
    s_full = [Barrier(f"s_full{b}") for b in range(2)]
    p_full = [Barrier(f"p_full{b}", softmax_warps) for b in range(2)]
    o_done, o_final = Barrier("o_done"), Barrier("o_final")
    st = {"kv": [None, None], "S": [None, None], "P": [[None] * softmax_warps for _ in range(2)], "pv_done": 0, "q": False}
    grow = [rng.random() < rescale_prob for _ in range(nsub)]  # all warps see the same data, hence the same decisions

    def producer():  # warp 0
        yield ("tma", lambda: st.__setitem__("q", True), q_full)
        for j in range(nblk):
            s = j % 2
            yield ("wait", kv_empty[s], ((j // 2) & 1) ^ 1)
            yield ("tma", (lambda j=j, s=s: st["kv"].__setitem__(s, j)), kv_full[s])

    def mma():  # warp 1
        def qk(i):
            def run():
                need(st["kv"][(i >> 1) % 2] == i >> 1, f"QK_{i} read tile {st['kv'][(i >> 1) % 2]}")
                st["S"][i & 1] = i
            return run

        def pv(i):
            def run():
                need(all(v == i for v in st["P"][i & 1]), f"PV_{i} read P of {st['P'][i & 1]}")
                need(st["kv"][(i >> 1) % 2] == i >> 1, f"PV_{i} read tile {st['kv'][(i >> 1) % 2]}")
                need(st["pv_done"] == i, f"PV_{i} executed after {st['pv_done']} PVs")
                st["pv_done"] = i + 1
            return run

        def issue_qk(i):
            j, s = i >> 1, (i >> 1) % 2
            if (i & 1) == 0:
                yield ("wait", kv_full[s], (j // 2) & 1)
            yield ("mma", qk(i))
            yield ("commit", s_full[i & 1])

        yield ("wait", q_full, 0)
        yield from issue_qk(0)
        for i in range(nsub):
            if i + 1 < nsub:
                yield from issue_qk(i + 1)
            yield ("wait", p_full[i & 1], (i >> 1) & 1)
            yield ("mma", pv(i))
            yield ("commit", o_done)
            if i + 1 == nsub:
                yield ("commit", o_final)
            if (i & 1) or i + 1 == nsub:
                yield ("commit", kv_empty[(i >> 1) % 2])

    def softmax(w):  # warps 2-5
        for i in range(nsub):
            yield ("wait", s_full[i & 1], (i >> 1) & 1)
            yield ("do", lambda i=i: need(st["S"][i & 1] == i, f"softmax({i}) read S of {st['S'][i & 1]}"))
            if i > 0 and grow[i]:
                yield ("wait", o_done, (i - 1) & 1)
                yield ("do", lambda i=i: need(st["pv_done"] == i, f"rescale({i}) with {st['pv_done']} PVs executed"))
            yield ("do", lambda i=i: st["P"][i & 1].__setitem__(w, i))
            yield ("arrive", p_full[i & 1])
        if two_waits:  # the first attempted fix: phase nsub - 2, then nsub - 1, on the per-sub-block barrier
            if nsub >= 2:
                yield ("wait", o_done, (nsub - 2) & 1)
            yield ("wait", o_done, (nsub - 1) & 1)
        elif first_epilogue:
            yield ("wait", o_done, (nsub - 1) & 1)
        else:
            yield ("wait", o_final, 0)
        yield ("do", lambda: need(st["pv_done"] == nsub, f"epilogue read O after {st['pv_done']} of {nsub} PVs"))

    m.add("tma", producer())
    m.add("mma", mma())
    for w in range(softmax_warps):
        m.add(f"softmax{w}", softmax(w))
    m.run()


@pytest.mark.parametrize("nsub", [1, 2, 3, 4, 7, 16])
def test_forward_generation7_protocol(nsub):
    for seed in range(120):
        forward_gen7(nsub, seed, lag=(seed % 6) * 0.17)


def test_checker_finds_the_epilogue_wait_the_forward_used_to_have():
    """The parity wait on the per-sub-block barrier: answered two phases back when the tensor pipe lags."""
    found = 0
    for seed in range(300):
        try:
            forward_gen7(4, seed, lag=0.9, first_epilogue=True)
        except Violation as e:
            assert "epilogue read O" in str(e), e
            found += 1
    assert found > 0


def test_checker_finds_the_deadlock_of_the_first_attempted_fix():
    """Waiting for phase nsub - 2 and then nsub - 1 on the same barrier hangs whenever both are already over (a parity wait
    two phases AHEAD never passes) -- which is the normal case at speed; on hardware it cost the round's last GPU minutes."""
    found = 0
    for seed in range(100):
        try:
            forward_gen7(4, seed, lag=0.0, two_waits=True)
        except Violation as e:
            assert "deadlock" in str(e), e
            found += 1
    assert found > 0


# ------------------------------------------------------------------------------------------------------------------
# backward, first form (attn_bwd_tc.cu::flash_bwd_kernel): S / dP / P / dS double-buffered, S / dP two ahead, 3 stages
# ------------------------------------------------------------------------------------------------------------------
def backward_form1(nsub, seed, lag, kv=True, softmax_warps=8, stages=3):
    m = Machine(seed, lag)
    nblk = (nsub + 1) // 2
    x_full = Barrier("x_full")
    y_full, y_empty = [Barrier(f"y_full{s}") for s in range(stages)], [Barrier(f"y_empty{s}") for s in range(stages)]
    sdp_full = [Barrier(f"sdp_full{b}") for b in range(2)]
    pds_full = [Barrier(f"pds_full{b}", softmax_warps) for b in range(2)]
    pds_free = [Barrier(f"pds_free{b}") for b in range(2)]
    acc_done = Barrier("acc_done")
    st = {"y": [None] * stages, "SDP": [None, None], "PDS": [[None] * softmax_warps for _ in range(2)], "acc": 0, "x": False}

    def producer():
        yield ("tma", lambda: st.__setitem__("x", True), x_full)
        for j in range(nblk):
            s = j % stages
            yield ("wait", y_empty[s], ((j // stages) & 1) ^ 1)
            yield ("tma", (lambda j=j, s=s: st["y"].__setitem__(s, j)), y_full[s])

    def mma():
        def sdp(i):
            def run():
                need(st["y"][(i >> 1) % stages] == i >> 1, f"S/dP_{i} read tile {st['y'][(i >> 1) % stages]}")
                st["SDP"][i & 1] = i
            return run

        def acc(i):
            def run():
                need(all(v == i for v in st["PDS"][i & 1]), f"acc_{i} read P/dS of {st['PDS'][i & 1]}")
                need(st["y"][(i >> 1) % stages] == i >> 1, f"acc_{i} read tile {st['y'][(i >> 1) % stages]}")
                st["acc"] += 1
            return run

        def issue_sdp(i):
            j, s = i >> 1, (i >> 1) % stages
            if (i & 1) == 0:
                yield ("wait", y_full[s], (j // stages) & 1)
            yield ("mma", sdp(i))
            yield ("commit", sdp_full[i & 1])

        yield ("wait", x_full, 0)
        yield from issue_sdp(0)
        if nsub > 1:
            yield from issue_sdp(1)
        for i in range(nsub):
            yield ("wait", pds_full[i & 1], (i >> 1) & 1)
            if i + 2 < nsub:
                yield from issue_sdp(i + 2)
            yield ("mma", acc(i))
            yield ("commit", pds_free[i & 1])
            if (i & 1) or i + 1 == nsub:
                yield ("commit", y_empty[(i >> 1) % stages])
        yield ("commit", acc_done)

    def softmax(w):
        for i in range(nsub):
            if kv and (i & 1) == 0:  # the per-column statistics arrive with the tile
                yield ("wait", y_full[(i >> 1) % stages], ((i >> 1) // stages) & 1)
                yield ("do", lambda i=i: need(st["y"][(i >> 1) % stages] == i >> 1, f"softmax({i}) read statistics of tile {st['y'][(i >> 1) % stages]}"))
            yield ("wait", sdp_full[i & 1], (i >> 1) & 1)
            yield ("do", lambda i=i: need(st["SDP"][i & 1] == i, f"softmax({i}) read S/dP of {st['SDP'][i & 1]}"))
            if i >= 2:
                yield ("wait", pds_free[i & 1], ((i >> 1) - 1) & 1)
                yield ("do", lambda i=i: need(st["acc"] >= i - 1, f"softmax({i}) overwrote P/dS with {st['acc']} accumulations executed"))
            yield ("do", lambda i=i: st["PDS"][i & 1].__setitem__(w, i))
            yield ("arrive", pds_full[i & 1])
        yield ("wait", acc_done, 0)
        yield ("do", lambda: need(st["acc"] == nsub, f"epilogue after {st['acc']} of {nsub} accumulations"))

    m.add("tma", producer())
    m.add("mma", mma())
    for w in range(softmax_warps):
        m.add(f"softmax{w}", softmax(w))
    m.run()


@pytest.mark.parametrize("kv", [False, True])
@pytest.mark.parametrize("nsub", [1, 2, 3, 5, 8, 13])
def test_backward_first_form_protocol(nsub, kv):
    for seed in range(60):
        backward_form1(nsub, seed, lag=(seed % 6) * 0.17, kv=kv)


# ------------------------------------------------------------------------------------------------------------------
# backward, second form (attn_bwd_tc.cu::flash_bwd2_kernel): single-buffered, P over S and dS over dP in place.  The
# in-place overwrite by the NEXT S / dP MMAs relies on the tensor pipe executing in issue order, which the model has.
# ------------------------------------------------------------------------------------------------------------------
def backward_form2(nsub, seed, lag, kv=True, softmax_warps=4, stages=2):
    m = Machine(seed, lag)
    nblk = (nsub + 1) // 2
    x_full = Barrier("x_full")
    y_full, y_empty = [Barrier(f"y_full{s}") for s in range(stages)], [Barrier(f"y_empty{s}") for s in range(stages)]
    sdp_full, pds_full, acc_done = Barrier("sdp_full"), Barrier("pds_full", softmax_warps), Barrier("acc_done")
    # one TMEM region: ('sdp', i) after the S / dP MMAs of sub-block i, ('pds', i) once every warp has overwritten its rows
    st = {"y": [None] * stages, "region": None, "rows": [None] * softmax_warps, "acc": 0}

    def producer():
        yield ("tma", lambda: None, x_full)
        for j in range(nblk):
            s = j % stages
            yield ("wait", y_empty[s], ((j // stages) & 1) ^ 1)
            yield ("tma", (lambda j=j, s=s: st["y"].__setitem__(s, j)), y_full[s])

    def mma():
        def sdp(i):
            def run():
                need(st["y"][(i >> 1) % stages] == i >> 1, f"S/dP_{i} read tile {st['y'][(i >> 1) % stages]}")
                need(st["acc"] == i, f"S/dP_{i} overwrote P/dS with {st['acc']} accumulations executed")
                st["region"], st["rows"] = ("sdp", i), [None] * softmax_warps
            return run

        def acc(i):
            def run():
                need(all(v == i for v in st["rows"]), f"acc_{i} read P/dS rows {st['rows']}")
                st["acc"] += 1
            return run

        yield ("wait", x_full, 0)
        for i in range(nsub):
            j, s = i >> 1, (i >> 1) % stages
            if (i & 1) == 0:
                yield ("wait", y_full[s], (j // stages) & 1)
            yield ("mma", sdp(i))
            yield ("commit", sdp_full)
            yield ("wait", pds_full, i & 1)
            yield ("mma", acc(i))
            if (i & 1) or i + 1 == nsub:
                yield ("commit", y_empty[s])
        yield ("commit", acc_done)

    def softmax(w):
        for i in range(nsub):
            if kv and (i & 1) == 0:
                yield ("wait", y_full[(i >> 1) % stages], ((i >> 1) // stages) & 1)
                yield ("do", lambda i=i: need(st["y"][(i >> 1) % stages] == i >> 1, f"softmax({i}) read statistics of tile {st['y'][(i >> 1) % stages]}"))
            yield ("wait", sdp_full, i & 1)
            yield ("do", lambda i=i: need(st["region"] == ("sdp", i), f"softmax({i}) read {st['region']}"))
            yield ("do", lambda i=i: st["rows"].__setitem__(w, i))
            yield ("arrive", pds_full)
        yield ("wait", acc_done, 0)
        yield ("do", lambda: need(st["acc"] == nsub, f"epilogue after {st['acc']} of {nsub} accumulations"))

    m.add("tma", producer())
    m.add("mma", mma())
    for w in range(softmax_warps):
        m.add(f"softmax{w}", softmax(w))
    m.run()


@pytest.mark.parametrize("kv", [False, True])
@pytest.mark.parametrize("nsub", [1, 2, 3, 5, 8, 13])
def test_backward_second_form_protocol(nsub, kv):
    for seed in range(60):
        backward_form2(nsub, seed, lag=(seed % 6) * 0.17, kv=kv)


# ------------------------------------------------------------------------------------------------------------------
# GEMM, one-CTA persistent kernel (gemm_tc.cu::gemm_bf16_persistent_kernel): ST-stage operand ring fed in bursts of G
# k blocks, accumulator double-buffered in TMEM, eight epilogue warps that hand a buffer back as soon as they have read it
# ------------------------------------------------------------------------------------------------------------------
def gemm_persistent(tiles, num_kb, stages, group, seed, lag, epi_warps=8):
    m = Machine(seed, lag)
    full, empty = [Barrier(f"full{s}") for s in range(stages)], [Barrier(f"empty{s}") for s in range(stages)]
    tmem_full, tmem_empty = [Barrier(f"tmem_full{b}") for b in range(2)], [Barrier(f"tmem_empty{b}", epi_warps) for b in range(2)]
    st = {"stage": [None] * stages, "acc": [None, None], "kdone": [0, 0], "read": [[None] * epi_warps for _ in range(2)]}

    def producer():
        kc = 0
        for t in range(tiles):
            kb = 0
            while kb < num_kb:
                g_n = min(group, num_kb - kb)
                for g in range(g_n):  # the whole burst's slots are awaited first, then requested
                    yield ("wait", empty[(kc + g) % stages], (((kc + g) // stages) & 1) ^ 1)
                for g in range(g_n):
                    s = kc % stages
                    yield ("tma", (lambda s=s, t=t, k=kb + g: st["stage"].__setitem__(s, (t, k))), full[s])
                    kc += 1
                kb += g_n

    def mma():
        kc = 0
        for it in range(tiles):
            buf = it & 1
            yield ("wait", tmem_empty[buf], ((it >> 1) & 1) ^ 1)

            def check_drained(it=it, buf=buf):
                need(it < 2 or all(v == it - 2 for v in st["read"][buf]), f"tile {it} overwrote buffer {buf} read by {st['read'][buf]}")
            yield ("do", check_drained)
            for kb in range(num_kb):
                s = kc % stages
                yield ("wait", full[s], (kc // stages) & 1)

                def run(it=it, kb=kb, s=s, buf=buf):
                    need(st["stage"][s] == (it, kb), f"MMA (tile {it}, k {kb}) read stage holding {st['stage'][s]}")
                    st["acc"][buf] = it
                    st["kdone"][buf] = kb + 1
                yield ("mma", run)
                yield ("commit", empty[s])
                kc += 1
            yield ("commit", tmem_full[buf])

    def epilogue(w):
        for it in range(tiles):
            buf = it & 1
            yield ("wait", tmem_full[buf], (it >> 1) & 1)
            yield ("do", lambda it=it, buf=buf: need(st["acc"][buf] == it and st["kdone"][buf] == num_kb,
                                                     f"epilogue of tile {it} read tile {st['acc'][buf]} after {st['kdone'][buf]} k blocks"))
            yield ("do", lambda it=it, buf=buf: st["read"][buf].__setitem__(w, it))
            yield ("arrive", tmem_empty[buf])

    m.add("producer", producer())
    m.add("mma", mma())
    for w in range(epi_warps):
        m.add(f"epi{w}", epilogue(w))
    m.run()


@pytest.mark.parametrize("stages,group", [(3, 1), (6, 2), (6, 3), (8, 2)])
@pytest.mark.parametrize("tiles,num_kb", [(1, 1), (2, 5), (5, 3), (7, 20)])
def test_gemm_persistent_protocol(tiles, num_kb, stages, group):
    for seed in range(30):
        gemm_persistent(tiles, num_kb, stages, group, seed, lag=(seed % 6) * 0.17)


# ------------------------------------------------------------------------------------------------------------------
# GEMM, CTA-pair kernel (gemm_tc.cu::gemm_bf16_pair_kernel, cta_group::2): both CTAs' producers signal the LEADER's full[s]
# (one expect_tx arrival by the leader + the bytes of both CTAs' loads), the leader's MMA thread multicasts its commits to
# both CTAs' empty[s] / tmem_full[buf], and the sixteen epilogue warps of the pair arrive on the leader's tmem_empty[buf]
# ------------------------------------------------------------------------------------------------------------------
def gemm_pair(tiles, num_kb, stages, group, seed, lag, epi_warps=8):
    m = Machine(seed, lag)
    full = [Barrier(f"full{s}", 3) for s in range(stages)]  # leader only: its arrive.expect_tx + the two CTAs' transfers
    empty = [[Barrier(f"empty{r}_{s}") for s in range(stages)] for r in range(2)]
    tmem_full = [[Barrier(f"tmem_full{r}_{b}") for b in range(2)] for r in range(2)]
    tmem_empty = [Barrier(f"tmem_empty{b}", 2 * epi_warps) for b in range(2)]  # leader only
    st = {"stage": [[None] * stages for _ in range(2)], "acc": [None, None], "kdone": [0, 0],
          "read": [[None] * (2 * epi_warps) for _ in range(2)]}

    def producer(rank):
        kc = 0
        for t in range(tiles):
            kb = 0
            while kb < num_kb:
                g_n = min(group, num_kb - kb)
                for g in range(g_n):
                    yield ("wait", empty[rank][(kc + g) % stages], (((kc + g) // stages) & 1) ^ 1)
                for g in range(g_n):
                    s = kc % stages
                    if rank == 0:
                        yield ("arrive", full[s])  # arrive.expect_tx(2 x stage bytes)
                    yield ("tma", (lambda s=s, t=t, k=kb + g: st["stage"][rank].__setitem__(s, (t, k))), full[s])
                    kc += 1
                kb += g_n

    def mma():  # the leader's warp 1
        kc = 0
        for it in range(tiles):
            buf = it & 1
            yield ("wait", tmem_empty[buf], ((it >> 1) & 1) ^ 1)
            yield ("do", lambda it=it, buf=buf: need(it < 2 or all(v == it - 2 for v in st["read"][buf]),
                                                     f"tile {it} overwrote buffer {buf} read by {st['read'][buf]}"))
            for kb in range(num_kb):
                s = kc % stages
                yield ("wait", full[s], (kc // stages) & 1)

                def run(it=it, kb=kb, s=s, buf=buf):
                    need(st["stage"][0][s] == (it, kb) and st["stage"][1][s] == (it, kb),
                         f"MMA (tile {it}, k {kb}) read stages holding {st['stage'][0][s]} / {st['stage'][1][s]}")
                    st["acc"][buf], st["kdone"][buf] = it, kb + 1
                yield ("mma", run)
                yield ("commit", [empty[0][s], empty[1][s]])
                kc += 1
            yield ("commit", [tmem_full[0][buf], tmem_full[1][buf]])

    def epilogue(rank, w):
        for it in range(tiles):
            buf = it & 1
            yield ("wait", tmem_full[rank][buf], (it >> 1) & 1)
            yield ("do", lambda it=it, buf=buf: need(st["acc"][buf] == it and st["kdone"][buf] == num_kb,
                                                     f"epilogue of tile {it} read tile {st['acc'][buf]} after {st['kdone'][buf]} k blocks"))
            yield ("do", lambda it=it, buf=buf: st["read"][buf].__setitem__(rank * epi_warps + w, it))
            yield ("arrive", tmem_empty[buf])

    for r in range(2):
        m.add(f"producer{r}", producer(r))
        for w in range(epi_warps):
            m.add(f"epi{r}_{w}", epilogue(r, w))
    m.add("mma", mma())
    m.run()


@pytest.mark.parametrize("stages,group", [(5, 1), (7, 2)])
@pytest.mark.parametrize("tiles,num_kb", [(1, 1), (2, 5), (5, 3), (6, 12)])
def test_gemm_pair_protocol(tiles, num_kb, stages, group):
    for seed in range(20):
        gemm_pair(tiles, num_kb, stages, group, seed, lag=(seed % 6) * 0.17)
