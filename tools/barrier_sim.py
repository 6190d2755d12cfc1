"""Randomised discrete-event model of the mbarrier protocols of the row-block tensor-core backward kernels
(csrc/local_bwd_tcrb.cu, csrc/local_bwd_tcrb10.cu): every warp that arrives on or waits for a barrier is an agent, the
scheduler interleaves them at random, `tcgen05.commit` arrivals are delayed until the issuer's earlier MMAs have
"completed" (in issue order, after a random delay).  It looks for deadlocks (nobody can run, not everybody is done) and
for ring-slot races (a slot rewritten before its readers are done, read before it is written).

mbarrier semantics modelled: an expected arrival count per phase; the phase completes when that many arrivals have
been counted -- whichever phase the arriving agent *meant*; `wait(parity)` succeeds when the barrier's current phase
parity differs from `parity` (so a waiter that falls two phases behind aliases).

    python tools/barrier_sim.py            # all configurations, 300 random schedules each

No GPU needed; `tests/test_host_cpu.py` runs a short version.
"""
from __future__ import annotations

import random
import sys


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def ready(self, parity):
        return (self.phase & 1) != parity


class Deadlock(Exception):
    pass


class Race(Exception):
    pass


class Sim:
    """Agents are generators yielding ('wait', bar, parity) | ('arrive', bar) | ('commit', issuer_id, bar) |
    ('mma', issuer_id) | ('own', slot, who) | ('release', slot, who) | ('work',)."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.agents = []           # [name, generator, pending_request]
        self.inflight = {}         # issuer -> list of ['mma' | ('commit', bar), remaining_delay]
        self.slots = {}            # slot -> (state, who)
        self.done = set()          # events that have happened: ('transformed', t, warp), ('mma_done', t, issuer)

    def add(self, name, gen):
        self.agents.append([name, gen, None])

    def run(self, max_steps=2_000_000):
        live = list(self.agents)
        for ag in live:
            ag[2] = next(ag[1], None)
        steps = 0
        while True:
            steps += 1
            if steps > max_steps:
                raise Deadlock("step limit")
            live = [a for a in live if a[2] is not None]
            # tensor pipe: retire in-flight work of every issuer in order, with random progress
            progressed = False
            for iss, q in self.inflight.items():
                if q and self.rng.random() < 0.5:
                    q[0][1] -= 1
                    if q[0][1] <= 0:
                        kind = q.pop(0)[0]
                        if kind != "mma":
                            kind[1].arrive()
                    progressed = True
            if not live and not any(self.inflight.values()):
                return steps
            runnable = [a for a in live if a[2][0] != "wait" or a[2][1].ready(a[2][2])]
            if not runnable:
                if any(self.inflight.values()):
                    continue
                raise Deadlock("; ".join(f"{a[0]} waits {a[2][1].name} p{a[2][2]} (phase {a[2][1].phase})" for a in live))
            ag = self.rng.choice(runnable)
            req = ag[2]
            if req[0] == "arrive":
                req[1].arrive()
            elif req[0] == "mma":
                self.inflight.setdefault(req[1], []).append(["mma", self.rng.randint(1, 4)])
            elif req[0] == "commit":
                q = self.inflight.setdefault(req[1], [])
                if q:
                    q.append([("commit", req[2]), 1])
                else:
                    req[2].arrive()
            elif req[0] == "own":
                st = self.slots.get(req[1], ("free", None))
                if st[0] != "free":
                    raise Race(f"{ag[0]} takes {req[1]} while {st}")
                self.slots[req[1]] = ("held", req[2])
            elif req[0] == "release":
                self.slots[req[1]] = ("free", None)
            elif req[0] == "done":
                self.done.add(req[1])
            elif req[0] == "need":
                if req[1] not in self.done:
                    raise Race(f"{ag[0]} passed its wait for {req[1]} before it happened")
            ag[2] = next(ag[1], None)


def rowblock(sim, *, nit, ns, nq_of_item, na, nraw, groups, split_issuers, ntile=2, tmem_bufs=1, warps_per_group=4,
             split_by="q", cont=None):
    """The protocol of local_bwd_tcrb*.cu.  nq_of_item(i) = source rows (stages) of item i per channel slice.
    cont(i) -> item i continues item i - 1 in the same image (local_bwd_tcrb10h.cu): its first two source rows are the
    last two of item i - 1 and are staged once -- the issuers add them into item i's accumulator buffer while they finish
    item i - 1 (after waiting for that buffer's tmem_ready), item i starts at q = 2, and every issuing warp commits
    accum_full once per item after its last MMA."""
    if cont is not None:
        return rowblock_shared(sim, nit=nit, nq_of_item=nq_of_item, na=na, nraw=nraw, ntile=ntile, tmem_bufs=tmem_bufs,
                               warps_per_group=warps_per_group, cont=cont)
    raw_full = [Bar(f"raw_full{s}", 1) for s in range(nraw)]
    raw_empty = [Bar(f"raw_empty{s}", warps_per_group) for s in range(nraw)]
    a_full = [Bar(f"a_full{s}", warps_per_group) for s in range(na)]
    a_empty = [Bar(f"a_empty{s}", 2) for s in range(na)]
    n_iss = 2 * (2 if split_issuers else 1)
    accum_full = [Bar(f"accum_full{b}", n_iss) for b in range(tmem_bufs)]
    tmem_ready = [Bar(f"tmem_ready{b}", 4) for b in range(tmem_bufs)]
    stages = [(i, js, q) for i in range(nit) for js in range(ns) for q in range(nq_of_item(i))]

    def producer():
        s, sph = 0, 0
        for t, _ in enumerate(stages):
            if t >= nraw:
                yield ("wait", raw_empty[s], sph ^ 1)
            yield ("own", ("raw", s), "tma")
            yield ("release", ("raw", s), "tma")
            yield ("arrive", raw_full[s])
            s += 1
            if s == nraw:
                s, sph = 0, sph ^ 1

    def transform(grp, warp):
        a = s = aph = sph = 0
        for t, _ in enumerate(stages):
            if groups == 1 or (t & 1) == grp:
                if t >= na:
                    yield ("wait", a_empty[a], aph ^ 1)
                yield ("wait", raw_full[s], sph)
                yield ("work",)
                yield ("arrive", raw_empty[s])
                yield ("done", ("transformed", t, warp))
                yield ("arrive", a_full[a])
            a += 1
            if a == na:
                a, aph = 0, aph ^ 1
            s += 1
            if s == nraw:
                s, sph = 0, sph ^ 1

    def issuer(mt, par, iid):
        a = aph = 0
        t = -1
        for i in range(nit):
            buf = i % tmem_bufs
            yield ("wait", tmem_ready[buf], (i // tmem_bufs) & 1)
            for js in range(ns):
                nq = nq_of_item(i)
                for q in range(nq):
                    t += 1
                    if not split_issuers or ((q if split_by == "q" else t) & 1) == par:
                        yield ("wait", a_full[a], aph)
                        for w in range(warps_per_group):
                            yield ("need", ("transformed", t, w))
                        if mt < ntile:
                            yield ("mma", iid)
                        yield ("commit", iid, a_empty[a])
                        last = (q >= nq - 2) if split_issuers else (q == nq - 1)
                        if last and js == ns - 1:
                            yield ("commit", iid, accum_full[buf])
                    a += 1
                    if a == na:
                        a, aph = 0, aph ^ 1

    def epilogue(warp):
        for b in range(tmem_bufs):
            yield ("arrive", tmem_ready[b])
        for i in range(nit):
            buf = i % tmem_bufs
            yield ("wait", accum_full[buf], (i // tmem_bufs) & 1)
            yield ("work",)
            if tmem_bufs > 1 or i + 1 < nit:
                yield ("arrive", tmem_ready[buf])

    sim.add("producer", producer())
    for g in range(groups):
        for w in range(warps_per_group):
            sim.add(f"transform g{g} w{w}", transform(g, w))
    iid = 0
    for mt in range(2):
        for par in range(2 if split_issuers else 1):
            sim.add(f"issuer mt{mt} par{par}", issuer(mt, par, iid))
            iid += 1
    for w in range(4):
        sim.add(f"epilogue w{w}", epilogue(w))


def rowblock_shared(sim, *, nit, nq_of_item, na, nraw, ntile, tmem_bufs, warps_per_group, cont):
    raw_full = [Bar(f"raw_full{s}", 1) for s in range(nraw)]
    raw_empty = [Bar(f"raw_empty{s}", warps_per_group) for s in range(nraw)]
    a_full = [Bar(f"a_full{s}", warps_per_group) for s in range(na)]
    a_empty = [Bar(f"a_empty{s}", 2) for s in range(na)]
    accum_full = [Bar(f"accum_full{b}", 4) for b in range(tmem_bufs)]
    tmem_ready = [Bar(f"tmem_ready{b}", 4) for b in range(tmem_bufs)]
    q0 = [2 if (i > 0 and cont(i)) else 0 for i in range(nit)]
    stages = [(i, q) for i in range(nit) for q in range(q0[i], nq_of_item(i))]

    def producer():
        s, sph = 0, 0
        for t, _ in enumerate(stages):
            if t >= nraw:
                yield ("wait", raw_empty[s], sph ^ 1)
            yield ("arrive", raw_full[s])
            s += 1
            if s == nraw:
                s, sph = 0, sph ^ 1

    def transform(grp, warp):
        a = s = aph = sph = 0
        for t, _ in enumerate(stages):
            if (t & 1) == grp:
                if t >= na:
                    yield ("wait", a_empty[a], aph ^ 1)
                yield ("wait", raw_full[s], sph)
                yield ("work",)
                yield ("arrive", raw_empty[s])
                yield ("done", ("transformed", t, warp))
                yield ("arrive", a_full[a])
            a += 1
            if a == na:
                a, aph = 0, aph ^ 1
            s += 1
            if s == nraw:
                s, sph = 0, sph ^ 1

    def issuer(mt, par, iid):
        a = aph = 0
        t = -1
        for i in range(nit):
            buf = i % tmem_bufs
            yield ("wait", tmem_ready[buf], (i // tmem_bufs) & 1)
            yield ("need", ("zeroed", i))
            nq = nq_of_item(i)
            nxt = i + 1 < nit and cont(i + 1)
            for q in range(q0[i], nq):
                t += 1
                if (t & 1) == par:
                    yield ("wait", a_full[a], aph)
                    for w in range(warps_per_group):
                        yield ("need", ("transformed", t, w))
                    if mt < ntile:
                        yield ("mma", iid)
                    if nxt and q >= nq - 2:
                        yield ("wait", tmem_ready[(i + 1) % tmem_bufs], ((i + 1) // tmem_bufs) & 1)
                        yield ("need", ("zeroed", i + 1))
                        if mt < ntile:
                            yield ("mma", iid)
                    yield ("commit", iid, a_empty[a])
                a += 1
                if a == na:
                    a, aph = 0, aph ^ 1
            yield ("commit", iid, accum_full[buf])

    def epilogue(warp):
        for b in range(tmem_bufs):
            yield ("done", ("zeroed", b))
            yield ("arrive", tmem_ready[b])
        for i in range(nit):
            buf = i % tmem_bufs
            yield ("wait", accum_full[buf], (i // tmem_bufs) & 1)
            yield ("work",)
            yield ("done", ("zeroed", i + tmem_bufs))
            yield ("arrive", tmem_ready[buf])

    sim.add("producer", producer())
    for g in range(2):
        for w in range(warps_per_group):
            sim.add(f"transform g{g} w{w}", transform(g, w))
    iid = 0
    for mt in range(2):
        for par in range(2):
            sim.add(f"issuer mt{mt} par{par}", issuer(mt, par, iid))
            iid += 1
    for w in range(4):
        sim.add(f"epilogue w{w}", epilogue(w))


def forward(sim, *, nkb, seg, nop, nraw, n_tw, n_iss, n_drain):
    """The protocol of the joint kernels (local_fwd_tc.cu: n_tw = 8 transform warps that also drain, one issuer;
    local_fwd_tcp.cu: 12 transform warps of which 8 drain, two issuers that both take every k-block)."""
    raw_full = [Bar(f"raw_full{s}", 1) for s in range(nraw)]
    raw_empty = [Bar(f"raw_empty{s}", n_tw) for s in range(nraw)]
    op_full = [Bar(f"op_full{s}", n_tw) for s in range(nop)]
    op_empty = [Bar(f"op_empty{s}", n_iss) for s in range(nop)]
    accum = Bar("accum", n_iss)
    drained = Bar("drained", n_tw)
    nseg = (nkb + seg - 1) // seg

    def producer():
        for k in range(nkb):
            s = k % nraw
            if k >= nraw:
                yield ("wait", raw_empty[s], ((k // nraw) & 1) ^ 1)
            yield ("arrive", raw_full[s])

    def issuer(iid):
        for k in range(nkb):
            o, kin = k % nop, k % seg
            if kin == 0 and k > 0:
                yield ("wait", drained, (k // seg - 1) & 1)
            yield ("wait", op_full[o], (k // nop) & 1)
            for w in range(n_tw):
                yield ("need", ("transformed", k, w))
            yield ("mma", iid)
            yield ("commit", iid, op_empty[o])
            if kin == seg - 1 or k == nkb - 1:
                yield ("commit", iid, accum)

    def transform(w):
        for k in range(nkb):
            s, o = k % nraw, k % nop
            yield ("wait", raw_full[s], (k // nraw) & 1)
            if k >= nop:
                yield ("wait", op_empty[o], ((k // nop) & 1) ^ 1)
            yield ("work",)
            yield ("done", ("transformed", k, w))
            yield ("arrive", op_full[o])
            yield ("arrive", raw_empty[s])
            if (k % seg) == seg - 1 or k == nkb - 1:
                sg = k // seg
                yield ("wait", accum, sg & 1)
                yield ("work",)
                if sg + 1 < nseg:
                    yield ("arrive", drained)

    sim.add("producer", producer())
    for i in range(n_iss):
        sim.add(f"issuer {i}", issuer(i))
    for w in range(n_tw):
        sim.add(f"transform w{w}", transform(w))


class Mark:
    """Stands in for a barrier in a ('commit', issuer, bar) request: records that the issuer's MMAs up to here are complete."""

    def __init__(self, sim, key):
        self.sim, self.key = sim, key

    def arrive(self):
        self.sim.done.add(self.key)


def joint10(sim, *, chunks, nxp=4, nyp=2, nset=2, n_stage=7, x_release="x_free"):
    """The protocol of the config-2 tensor-core joint (local_fwd_tcj10.cu).  `chunks` = y row pairs per chunk (a chunk has
    one x pair more); x pair P of a chunk is read by its y pairs P - 1 and P.  x slots are released by the issuers' x_free
    commits (two per x pair), y slots by y_done, accumulator sets by the four drain warps.  x_release = "y_done" is the
    first version of the ring-4 kernel, which waited for the y pair that last read the slot on the y_done barrier of THAT
    pair's slot: with a ring deeper than nyp + 1 the waiter can fall two phases behind and never wakes (seen on the GPU)."""
    x_full = [Bar(f"x_full{s}", n_stage) for s in range(nxp)]
    x_free = [Bar(f"x_free{s}", 2) for s in range(nxp)]
    y_full = [Bar(f"y_full{s}", n_stage) for s in range(nyp)]
    y_done = [Bar(f"y_done{s}", 1) for s in range(nyp)]
    set_free = [Bar(f"set_free{s}", 4) for s in range(nset)]
    # global numbering
    xs, ys = [], []            # xs[g] = (chunk, local P, readers as global y indices); ys[q] = (chunk, local Q, xg0)
    for c, npy in enumerate(chunks):
        xb, yb = len(xs), len(ys)
        for P in range(npy + 1):
            xs.append((c, P, [yb + Q for Q in (P - 1, P) if 0 <= Q < npy]))
        for Q in range(npy):
            ys.append((c, Q, xb + Q))

    def x_stage(w):
        for g, (c, P, readers) in enumerate(xs):
            if g >= nxp:
                if x_release == "x_free":
                    yield ("wait", x_free[g % nxp], (g // nxp - 1) & 1)
                else:
                    # y pair that last read the previous occupant, or the last y pair of the chunk before
                    yb = sum(chunks[:c])
                    wy = yb + P - nxp if P >= nxp else yb - 1
                    if wy >= 0:
                        yield ("wait", y_done[wy % nyp], (wy // nyp) & 1)
                for q in xs[g - nxp][2]:
                    yield ("need", ("mma_done", q))
            yield ("work",)
            yield ("done", ("x", g, w))
            yield ("arrive", x_full[g % nxp])

    def y_stage(w):
        for q in range(len(ys)):
            if q >= nyp:
                yield ("wait", y_done[q % nyp], ((q - nyp) // nyp) & 1)
                yield ("need", ("mma_done", q - nyp))
            yield ("work",)
            yield ("done", ("y", q, w))
            yield ("arrive", y_full[q % nyp])

    def issuer(i):
        for q, (c, Q, xg0) in enumerate(ys):
            if (q & 1) != i:
                continue
            xg1 = xg0 + 1
            yield ("wait", y_full[q % nyp], (q // nyp) & 1)
            yield ("wait", x_full[xg1 % nxp], (xg1 // nxp) & 1)
            if q >= nset:
                yield ("wait", set_free[q % nset], ((q - nset) // nset) & 1)
                for d in range(4):
                    yield ("need", ("drained", q - nset, d))
            for w in range(n_stage):
                yield ("need", ("y", q, w))
                yield ("need", ("x", xg0, w))
                yield ("need", ("x", xg1, w))
            yield ("mma", i)
            yield ("commit", i, Mark(sim, ("mma_done", q)))
            yield ("commit", i, y_done[q % nyp])
            yield ("commit", i, x_free[xg0 % nxp])
            if Q == 0:
                yield ("commit", i, x_free[xg0 % nxp])
            yield ("commit", i, x_free[xg1 % nxp])
            if Q == chunks[c] - 1:
                yield ("commit", i, x_free[xg1 % nxp])

    def drain(d):
        for q in range(len(ys)):
            yield ("wait", y_done[q % nyp], (q // nyp) & 1)
            yield ("need", ("mma_done", q))
            yield ("work",)
            yield ("done", ("drained", q, d))
            yield ("arrive", set_free[q % nset])

    for w in range(n_stage):
        sim.add(f"x stage w{w}", x_stage(w))
        sim.add(f"y stage w{w}", y_stage(w))
    for i in range(2):
        sim.add(f"issuer {i}", issuer(i))
    for d in range(4):
        sim.add(f"drain {d}", drain(d))


JOINT10_CONFIGS = {
    "tcj10 (adopted): x ring 4 released by x_free": dict(chunks=[12, 13, 1, 5, 2, 9]),
    "tcj10, single-pair chunks": dict(chunks=[1, 1, 1, 2, 1, 1, 3, 1]),
    "tcj10, x ring 3": dict(chunks=[7, 6, 9], nxp=3),
    "tcj10 ring 4 released through y_done (bug)": dict(chunks=[12, 13, 7], x_release="y_done"),
}


FORWARD_CONFIGS = {
    "joint K=128 (local_fwd_tc.cu)": dict(nkb=70, seg=32, nop=2, nraw=4, n_tw=8, n_iss=1, n_drain=8),
    "packed joint K=20, T=7 (local_fwd_tcp.cu)": dict(nkb=150, seg=64, nop=3, nraw=4, n_tw=12, n_iss=2, n_drain=8),
    "packed joint K=20, T=3": dict(nkb=150, seg=64, nop=4, nraw=4, n_tw=12, n_iss=2, n_drain=8),
}


CONFIGS = {
    # the K = 10 kernel: two transform groups, four issuers, double-buffered accumulators, even ring depths 4 / 6, and
    # issuers split by the GLOBAL stage parity, so that a ring slot always belongs to the same group and issuer pair
    "tcrb10 (adopted)": dict(nit=7, ns=1, nq_of_item=lambda i: [10, 9, 10, 3, 10, 5, 10][i], na=4, nraw=6, groups=2,
                             split_issuers=True, tmem_bufs=2, split_by="t"),
    "tcrb10, one pixel tile": dict(nit=5, ns=1, nq_of_item=lambda i: [10, 9, 7, 10, 4][i], na=4, nraw=6, groups=2,
                                   split_issuers=True, tmem_bufs=2, ntile=1, split_by="t"),
    # the kernel as it ships since late round 2: boundary source rows of consecutive chunks staged once (7-row chunks of
    # one image, then a new image, single-row and two-row items)
    "tcrb10h, shared boundary rows (adopted)": dict(nit=9, ns=1, nq_of_item=lambda i: [9, 9, 9, 8, 9, 9, 4, 3, 9][i], na=6, nraw=6,
                                                    groups=2, split_issuers=True, tmem_bufs=2, split_by="t",
                                                    cont=lambda i: i not in (0, 4, 7)),
    # its first version split the issuers by the row index inside a chunk: after a chunk with an odd number of rows a
    # slot alternates between issuer pairs, and an issuer that skipped a phase of a_full passes its parity wait early
    "tcrb10 with issuers split by row-in-chunk (bug)": dict(nit=7, ns=1, nq_of_item=lambda i: [10, 9, 10, 3, 10, 5, 10][i], na=4,
                                                            nraw=6, groups=2, split_issuers=True, tmem_bufs=2),
    # the K = 20 kernel as built: one transform group, four issuers
    "tcrb T=7 (adopted)": dict(nit=3, ns=3, nq_of_item=lambda i: 27, na=3, nraw=3, groups=1, split_issuers=True),
    "tcrb T=3 (adopted)": dict(nit=3, ns=3, nq_of_item=lambda i: 12, na=6, nraw=8, groups=1, split_issuers=True),
    # the reverted experiment: two transform groups with the odd ring depths of T = 7
    "tcrb T=7, two transform groups (reverted)": dict(nit=3, ns=3, nq_of_item=lambda i: 27, na=3, nraw=3, groups=2,
                                                      split_issuers=True, ntile=1),
}


def check(name, cfg, runs, seed0=0, builder=None):
    bad = None
    builder = builder or (forward if name in FORWARD_CONFIGS else joint10 if name in JOINT10_CONFIGS else rowblock)
    for r in range(runs):
        sim = Sim(seed0 + r)
        builder(sim, **cfg)
        try:
            sim.run()
        except (Deadlock, Race) as e:
            bad = f"seed {seed0 + r}: {type(e).__name__}: {e}"
            break
    return bad


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    for name, cfg in list(CONFIGS.items()) + list(FORWARD_CONFIGS.items()) + list(JOINT10_CONFIGS.items()):
        bad = check(name, cfg, runs)
        print(f"{name:45s} {'OK (' + str(runs) + ' schedules)' if bad is None else bad}")
