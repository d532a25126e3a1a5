"""Discrete-event model of the barrier protocol of the persistent backward kernel (csrc/attn_bwd_v3.cu).

The kernel's 28 warps talk through ~50 mbarriers whose waits are PARITY probes: `try_wait.parity P` succeeds as soon as the
barrier's current phase has the other parity -- it cannot tell "the phase I want has completed" from "the barrier is still
one phase BEHIND what I think" or "two phases ahead".  Both bugs round 2 hit on hardware were of that family (a producer lane
probing all_done two phases late; a second arrive.expect_tx on a k_full phase that was still open).  This model restates the
roles' loops with the kernel's own index formulas and checks, over random schedules and random item lengths (1 .. 8 tiles, as
with a causal mask):
  * no deadlock (every role terminates),
  * no arrival on a barrier phase that is already complete (arrival-count underflow = the hardware fault),
  * every probe is UNAMBIGUOUS: when a wait succeeds, the barrier's true phase is exactly the one the waiter meant + 1,
  * every buffer is read with the contents the reader expects and overwritten only after its readers are done
    (ring slots, S/dP TMEM buffers, P/dS columns + boxes, K shared-memory buffers, K/V TMEM copies, dQ and dV/dK accumulators).
It is a model of the protocol, not of the arithmetic; the GPU tests cover the kernel itself."""
import heapq
import random

import pytest

NSUB = 4


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        assert self.pending > 0, f"arrival on completed phase of {self.name} (count underflow)"
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count


class Sim:
    """Roles are generators yielding ('wait', bar, phase_wanted) or ('delay', cycles); everything else happens inline."""

    def __init__(self, n_iters, slots, seed, stall_p=0.0):
        self.rng = random.Random(seed)
        self.stall_p = stall_p
        self.now = 0
        self.events = []        # (time, seq, fn): asynchronous completions (TMA loads, MMA commits)
        self.seq = 0
        self.n_iters = n_iters  # tiles per (non-empty) item of this CTA
        self.S = slots
        B = lambda n, c: Bar(n, c)      # noqa: E731
        self.k_full, self.kt_ready = B("k_full", 1), B("kt_ready", 4)
        self.qdo_full = [B(f"qdo_full{i}", 1) for i in range(slots)]
        self.qdo_empty = [B(f"qdo_empty{i}", 2) for i in range(slots)]
        self.sdp_full = [B(f"sdp_full{j}", 1) for j in range(NSUB)]
        self.pds_full = [B(f"pds_full{i}", 4) for i in range(2 * NSUB)]      # (model: 4 warps instead of 128 threads)
        self.pds_free = [B(f"pds_free{j}", 1) for j in range(NSUB)]
        self.s_empty = [B(f"s_empty{j}", 4) for j in range(NSUB)]
        self.dq_full, self.dq_empty = B("dq_full", 1), B("dq_empty", 4)
        self.all_done, self.sdp_done = B("all_done", 3), B("sdp_done", 4)
        self.box_free = [B("box_free0", 2), B("box_free1", 2)]
        self.b_turn = [B("b_turn0", 1), B("b_turn1", 1)]
        self.kbuf_free = [B("kbuf_free0", 1), B("kbuf_free1", 1)]
        # resources: contents tags
        self.ring = [None] * slots           # ('V', it) or ('H', it, u)
        self.kbuf = [None, None]             # item whose K sits in the shared-memory buffer
        self.kv_tmem = None                  # item whose K, V sit in TMEM
        self.sbuf = [None, None]             # (it, k, j) whose S^T/dP^T sit in TMEM buffer
        self.pcols = [None] * NSUB           # (it, k) whose P^T/dS^T sit in warpgroup j's columns
        self.box = [[None] * NSUB, [None] * NSUB]
        self.dq_tmem = None
        self.acc = ("zero", 0)               # dV/dK accumulators: ('zero', it) | ('acc', it, n_subtiles)
        self.k_in_flight = 0

    # ---- helpers -------------------------------------------------------------------------------------------------
    def later(self, lo, hi, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + self.rng.randint(lo, hi), self.seq, fn))

    def commit(self, order_state, bars, fn=None):
        """tcgen05.commit: the arrivals happen when every MMA this issuer has issued so far completed (in issue order)."""
        t = max(order_state[0], self.now) + self.rng.randint(50, 400)
        order_state[0] = t
        self.seq += 1

        def done():
            if fn:
                fn()
            for b in bars:
                b.arrive()
        heapq.heappush(self.events, (t, self.seq, done))

    # ---- roles (index formulas as in the kernel) -----------------------------------------------------------------------
    def producer(self):
        it = kb = 0
        n_items = len(self.n_iters)
        for item, n_iter in enumerate(self.n_iters):
            if item == 0:
                self.load_v(0, 0, 0)
                self.load_k(0)
            for u in range(2 * n_iter):
                ug = 2 * kb + it + 1 + u
                s = ug % self.S
                yield ("wait", self.qdo_empty[s], ug // self.S - 1)
                assert self.ring[s] is None, ("ring slot overwritten", s, self.ring[s])
                self.ring[s] = ("loading",)
                self.later(300, 1500, lambda s=s, it=it, u=u: (self.ring.__setitem__(s, ("H", it, u)), self.qdo_full[s].arrive()))
                yield ("delay", self.rng.randint(100, 500))
            if item + 1 < n_items:
                pv = 2 * (kb + n_iter) + it + 1
                sv = pv % self.S
                yield ("wait", self.qdo_empty[sv], pv // self.S - 1)
                self.load_v(sv, it + 1, pv)
                # (a lane can be descheduled for a long time anywhere: the cross-role waits must hold up under that)
                yield ("delay", self.rng.choice([0, 0, 0, 4000, 40000]))
                if it > 0:
                    yield ("wait", self.kbuf_free[(it + 1) & 1], (it - 1) >> 1)
                yield ("wait", self.k_full, it)            # one arrival per phase: K(it) must have landed
                yield ("wait", self.kt_ready, it)          # ... and every warp of warpgroup 0 must have seen that phase
                self.load_k(it + 1)
            kb += n_iter
            it += 1

    def load_v(self, sv, item_no, pv):
        assert self.ring[sv] is None, ("ring slot overwritten by V", sv, self.ring[sv])
        self.ring[sv] = ("loading",)
        self.later(300, 1500, lambda: (self.ring.__setitem__(sv, ("V", item_no)), self.qdo_full[sv].arrive()))

    def load_k(self, item_no):
        buf = item_no & 1
        assert self.kbuf[buf] is None or self.kbuf[buf] == ("free",), ("K buffer overwritten", buf, self.kbuf[buf])
        # arrive.expect_tx is THE arrival of a one-arrival barrier: a second one while the bytes of the first are still in flight
        # lands on the same open phase (the hardware fault of round 2)
        assert self.k_in_flight == 0, "k_full armed while its phase is open (arrival-count underflow)"
        self.k_in_flight += 1
        self.kbuf[buf] = ("loading",)

        def landed():
            self.kbuf[buf] = item_no
            self.k_in_flight -= 1
            self.k_full.arrive()
        self.later(300, 1500, landed)

    def mma_a(self, j):
        it = kb = 0
        order = [0]
        for n_iter in self.n_iters:
            T = NSUB * n_iter
            yield ("wait", self.kt_ready, it)
            for t in range(j, T, NSUB):
                tg = 4 * kb + t
                u = (tg >> 1) + it + 1
                yield ("wait", self.qdo_full[u % self.S], u // self.S)
                assert self.ring[u % self.S] in (("H", it, t >> 1), ("H1", it, t >> 1)), ("A reads wrong ring contents", self.ring[u % self.S], it, t)
                if tg >= 2:
                    yield ("wait", self.s_empty[(j + 2) & 3], (tg - 2) >> 2)
                assert self.kv_tmem == it, ("S MMA with another item's K/V in TMEM", self.kv_tmem, it)
                assert self.sbuf[t & 1] is None, ("S buffer overwritten", self.sbuf[t & 1])
                self.sbuf[t & 1] = ("mma",)
                bars = [self.sdp_full[j]] + ([self.sdp_done] if t + NSUB >= T else [])
                self.commit(order, bars, lambda b=t & 1, tag=(it, t >> 2, j): self.sbuf.__setitem__(b, tag))
                yield ("delay", self.rng.randint(100, 400))
            kb += n_iter
            it += 1

    def mma_b(self, w):
        it = kb = 0
        order = [0]
        for n_iter in self.n_iters:
            T = NSUB * n_iter
            for t in range(w, T, 2):
                j = t & 3
                tg, kg = 4 * kb + t, kb + (t >> 2)
                yield ("wait", self.pds_full[(kg & 1) * NSUB + j], kg >> 1)
                assert self.pcols[j] == (it, t >> 2), ("B reads wrong P/dS", self.pcols[j], it, t)
                if tg > 0:
                    yield ("wait", self.b_turn[tg & 1], (tg - 1) >> 1)
                slot = ((tg >> 1) + it + 1) % self.S
                assert self.ring[slot] in (("H", it, t >> 1), ("H1", it, t >> 1)), ("B reads wrong ring contents", self.ring[slot], it, t)
                if t == 0:
                    assert self.acc == ("zero", it), ("dV/dK accumulators not zeroed for this item", self.acc, it)
                    self.acc = ("acc", it, 0)
                assert self.acc[0] == "acc" and self.acc[1] == it and self.acc[2] == t, ("dV/dK issue order", self.acc, it, t)
                self.acc = ("acc", it, t + 1)

                def released(j=j, slot=slot):
                    self.pcols[j] = None

                self.commit(order, [self.pds_free[j]], released)
                self.commit(order, [self.qdo_empty[slot]], lambda slot=slot, odd=t & 1: self.release_half(slot, odd))
                if t >= T - 2:
                    self.commit(order, [self.all_done])
                self.b_turn[(tg + 1) & 1].arrive()
                yield ("delay", self.rng.randint(100, 500))
            kb += n_iter
            it += 1

    def release_half(self, slot, odd):
        tag = self.ring[slot]
        if isinstance(tag, tuple) and len(tag) == 3 and tag[0] == "H":
            self.ring[slot] = ("H1", tag[1], tag[2])       # first of the two releases
        else:
            self.ring[slot] = None

    def warp_c(self):
        it = kb = 0
        order = [0]
        for n_iter in self.n_iters:
            for k in range(n_iter):
                kg = kb + k
                if kg > 0:
                    self.box_free[(kg - 1) & 1].arrive()
                for j in range(NSUB):
                    yield ("wait", self.pds_full[(kg & 1) * NSUB + j], kg >> 1)
                    assert self.box[kg & 1][j] == (it, k), ("C reads wrong dS box", self.box[kg & 1][j], it, k)
                if kg > 0:
                    yield ("wait", self.dq_empty, kg - 1)
                assert self.kbuf[it & 1] == it, ("dQ MMA with another item's K in shared memory", self.kbuf, it)
                assert self.dq_tmem is None, ("dQ accumulator overwritten", self.dq_tmem)
                self.dq_tmem = ("mma",)

                def done(kg=kg, tag=(it, k)):
                    self.dq_tmem = tag
                    for j in range(NSUB):
                        self.box[kg & 1][j] = None

                self.commit(order, [self.dq_full, self.box_free[kg & 1]], done)
                if k == n_iter - 1:
                    self.commit(order, [self.all_done, self.kbuf_free[it & 1]], lambda b=it & 1: self.kbuf.__setitem__(b, ("free",)))
                yield ("delay", self.rng.randint(100, 600))
            kb += n_iter
            it += 1

    def drain(self, w):
        it = kb = 0
        for n_iter in self.n_iters:
            for k in range(n_iter):
                yield ("wait", self.dq_full, kb + k)
                assert self.dq_tmem == (it, k), ("drain reads wrong dQ", self.dq_tmem, it, k)
                yield ("delay", self.rng.randint(50, 300))
                if self.dq_empty.pending == 1:
                    self.dq_tmem = None
                self.dq_empty.arrive()
                yield ("delay", self.rng.randint(200, 3000))
            kb += n_iter
            it += 1

    def compute(self, wg, w):
        """One of the four warps of compute warpgroup wg (w = 0 is the one that stands for 'thread 0' actions)."""
        it = kb = 0
        n_items = len(self.n_iters)
        for item, n_iter in enumerate(self.n_iters):
            if it == 0 and wg == 0:
                yield from self.kv_to_tmem(0, 0, w)
            for k in range(n_iter):
                kg = kb + k
                yield ("wait", self.sdp_full[wg], kg)
                assert self.sbuf[wg & 1] == (it, k, wg), ("compute reads wrong S buffer", self.sbuf[wg & 1], it, k, wg)
                yield ("delay", self.rng.randint(50, 200))
                if self.s_empty[wg].pending == 1:
                    self.sbuf[wg & 1] = None
                self.s_empty[wg].arrive()
                yield ("delay", self.rng.randint(300, 1500))
                if kg > 0:
                    yield ("wait", self.pds_free[wg], kg - 1)
                    if kg >= 2:
                        yield ("wait", self.box_free[kg & 1], (kg >> 1) - 1)
                assert self.pcols[wg] in (None, (it, k)), ("P/dS columns overwritten", self.pcols[wg], it, k)
                assert self.box[kg & 1][wg] in (None, (it, k)), ("dS box overwritten", self.box[kg & 1][wg], it, k)
                self.pcols[wg] = (it, k)
                self.box[kg & 1][wg] = (it, k)
                yield ("delay", self.rng.randint(300, 1500))
                self.pds_full[(kg & 1) * NSUB + wg].arrive()
            if wg == 0 and item + 1 < n_items:
                yield ("wait", self.sdp_done, it)
                yield from self.kv_to_tmem(it + 1, 2 * (kb + n_iter) + it + 1, w)
            yield ("wait", self.all_done, it)
            assert self.acc == ("acc", it, NSUB * n_iter), ("epilogue reads unfinished dV/dK", self.acc, it, n_iter)
            yield ("delay", self.rng.randint(100, 1500))
            yield ("barrier", "epilogue")                     # named_bar_sync(6, 512): every compute warp has read dV / dK
            if wg == 0 and w == 0:
                self.acc = ("zero", it + 1)
            kb += n_iter
            it += 1

    def kv_to_tmem(self, item_no, pv, w):
        sv = pv % self.S
        yield ("wait", self.k_full, item_no)
        yield ("wait", self.qdo_full[sv], pv // self.S)
        assert self.kbuf[item_no & 1] == item_no, ("K buffer holds another item", self.kbuf, item_no)
        assert self.ring[sv] == ("V", item_no), ("V slot holds something else", self.ring[sv], item_no)
        if w == 0:
            assert all(x is None or x == ("mma",) or x[0] != item_no - 1 or True for x in self.sbuf)
            self.kv_tmem = item_no
        yield ("delay", self.rng.randint(100, 400))
        self.kt_ready.arrive()
        yield ("barrier", "kv")                               # named_bar_sync(7, 128)
        if w == 0:
            self.ring[sv] = None
            self.qdo_empty[sv].arrive()
            self.qdo_empty[sv].arrive()

    # ---- scheduler ---------------------------------------------------------------------------------------------------
    def run(self):
        roles = {"producer": self.producer(), "C": self.warp_c()}
        for j in range(NSUB):
            roles[f"A{j}"] = self.mma_a(j)
        for w in range(2):
            roles[f"B{w}"] = self.mma_b(w)
        for w in range(4):
            roles[f"drain{w}"] = self.drain(w)
        for wg in range(NSUB):
            for w in range(4):
                roles[f"wg{wg}.{w}"] = self.compute(wg, w)
        state = {n: ("ready", None) for n in roles}           # ready | ('wait', bar, phase) | ('sleep', t) | ('barrier', id)
        barrier_groups = {"epilogue": 16, "kv": 4}
        at_barrier = {"epilogue": [], "kv": []}
        live = set(roles)
        guard = 0
        while live:
            guard += 1
            assert guard < 2_000_000, "model did not terminate"
            progressed = False
            names = list(live)
            self.rng.shuffle(names)
            for n in names:
                st = state[n]
                if st[0] == "wait":
                    bar, want = st[1], st[2]
                    if want < 0:
                        ok = True                               # the parity of "phase -1": a fresh barrier passes
                    else:
                        ok = (bar.phase & 1) != (want & 1)      # what the hardware probe sees
                        if ok:
                            assert bar.phase == want + 1, f"{n}: ambiguous probe on {bar.name}: wanted phase {want}, barrier is in {bar.phase}"
                    if not ok:
                        continue
                elif st[0] == "stalled":
                    if self.now < st[1]:
                        continue
                    state[n] = ("wait", st[2], st[3])
                    progressed = True
                    continue
                elif st[0] == "sleep":
                    if self.now < st[1]:
                        continue
                elif st[0] == "barrier":
                    continue
                try:
                    req = next(roles[n])
                except StopIteration:
                    live.discard(n)
                    progressed = True
                    continue
                progressed = True
                if req[0] == "wait":
                    state[n] = ("wait", req[1], req[2])
                    if self.stall_p and self.rng.random() < self.stall_p:
                        # any warp can lose the scheduler for a long time right before a probe: hold the probe back
                        state[n] = ("stalled", self.now + self.rng.choice([3000, 20000, 60000]), req[1], req[2])
                elif req[0] == "delay":
                    state[n] = ("sleep", self.now + req[1])
                elif req[0] == "barrier":
                    state[n] = ("barrier", req[1])
                    at_barrier[req[1]].append(n)
                    if len(at_barrier[req[1]]) == barrier_groups[req[1]]:
                        for m in at_barrier[req[1]]:
                            state[m] = ("ready", None)
                        at_barrier[req[1]] = []
            # advance time: fire the next asynchronous completion, or move to the next wake-up
            if self.events and (not progressed or self.events[0][0] <= self.now):
                t, _, fn = heapq.heappop(self.events)
                self.now = max(self.now, t)
                fn()
                progressed = True
            if not progressed:
                sleepers = [st[1] for st in state.values() if st[0] in ("sleep", "stalled")]
                if sleepers:
                    self.now = max(self.now + 1, min(sleepers))
                elif not self.events:
                    blocked = {n: (state[n][1].name, state[n][2], state[n][1].phase) if state[n][0] == "wait" else state[n] for n in live}
                    raise AssertionError(f"deadlock: {blocked}")
            else:
                self.now += self.rng.randint(1, 50)


@pytest.mark.parametrize("slots", [5, 6])
def test_uniform_items_random_schedules(slots):
    for seed in range(15):
        Sim([8] * 4, slots, seed).run()


@pytest.mark.parametrize("slots", [5, 6])
def test_short_and_mixed_items_as_with_a_causal_mask(slots):
    rng = random.Random(1234)
    for seed in range(60):
        n_iters = [rng.randint(1, 8) for _ in range(rng.randint(1, 7))]
        Sim(n_iters, slots, seed).run()
    for seed in range(25):
        Sim([1] * 9, slots, 1000 + seed).run()              # one-tile items only: the producer runs furthest ahead


@pytest.mark.parametrize("slots", [5, 6])
def test_any_warp_may_stall_before_any_probe(slots):
    """The parity-lag hazard in general: a waiter that loses the scheduler for tens of thousands of cycles right before a probe
    must still find its barrier in the phase it expects (i.e. the barrier's next phase must depend on the waiter)."""
    rng = random.Random(99)
    for seed in range(16):
        n_iters = [rng.randint(1, 6) for _ in range(rng.randint(2, 5))]
        Sim(n_iters, slots, 5000 + seed, stall_p=0.03).run()
    for seed in range(12):
        Sim([1] * 7, slots, 7000 + seed, stall_p=0.05).run()


def test_the_model_sees_the_two_bugs_found_on_hardware():
    """Remove each fix in turn: the model must object (else it would not have caught them)."""
    class ArmsKFullEarly(Sim):
        def producer(self):
            for step in Sim.producer(self):
                if step[0] == "wait" and (step[1] is self.k_full or step[1] is self.kt_ready):
                    continue                                   # skip "K(it) has landed and been seen" before arming the next phase
                yield step

    with pytest.raises(AssertionError):
        for seed in range(60):
            ArmsKFullEarly([1] * 9, 6, seed).run()

    class NoKtReadyWait(Sim):
        def producer(self):
            for step in Sim.producer(self):
                if step[0] == "wait" and step[1] is self.kt_ready:
                    continue                                   # found by THIS model: a warp of warpgroup 0 that stalls before its
                yield step                                     # k_full probe finds the barrier two phases on

    with pytest.raises(AssertionError):
        for seed in range(40):
            NoKtReadyWait([1] * 8, 5, 7000 + seed, stall_p=0.05).run()

    class WaitsAllDoneLate(Sim):
        def producer(self):
            for step in Sim.producer(self):
                if step[0] == "wait" and step[1] in self.kbuf_free:
                    it_prev = None
                    # the first version waited on all_done(it - 1) here instead of a barrier of the K buffer's own
                    b = step[1]
                    idx = self.kbuf_free.index(b)
                    it_prev = 2 * step[2] + idx                # the item whose release is awaited
                    yield ("wait", self.all_done, it_prev)
                    continue
                yield step

    with pytest.raises(AssertionError):
        for seed in range(200):
            WaitsAllDoneLate([1] * 12, 6, seed).run()
