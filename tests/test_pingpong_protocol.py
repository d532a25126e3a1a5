"""Protocol model of the two-query-tile forward kernel (csrc/attn_fwd_pingpong.cu), run on the CPU.

The kernel was written without a GPU at hand, so its synchronisation -- mbarrier phases derived from running counters, a
K/V ring shared by two MMA-issuing warps (released by two arrivals, also for key tiles one of the query tiles does not
need), the exp token passed through two named barriers, work items of unequal tile counts -- is restated here as a
discrete-event model with the hardware semantics that matter:

  * an mbarrier completes a phase after `count` arrivals; `wait(parity)` passes once the phase of that parity has completed
    (i.e. the current phase has the other parity), exactly like mbarrier.try_wait.parity -- including its blindness to
    being two phases late, which is why every buffer also carries a tag (work item, tile) that readers check;
  * TMA loads land after a latency and complete the `full` barrier; tcgen05.mma of one issuing warp execute in issue
    order, take time, read their operands while they run and signal `tcgen05.commit` barriers when everything issued
    before the commit has finished;
  * `bar.sync` / `bar.arrive` on a named barrier of 256 threads: one warpgroup blocks, the other's arrival releases it.

Random latencies (seeded) shake the interleavings.  The test fails on a deadlock, on a reader seeing the wrong tag
(a buffer overwritten too early or read too early), and on two exp phases overlapping.
"""
import heapq
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):                  # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class NamedBar:
    """bar.sync by one warpgroup + bar.arrive by the other (count 256 = 128 + 128): a counting hand-over."""
    def __init__(self):
        self.arrivals = 0

    def arrive(self):
        self.arrivals += 1

    def try_sync(self):
        if self.arrivals > 0:
            self.arrivals -= 1
            return True
        return False


class Sim:
    def __init__(self, items, seed, kv_stages=2):
        self.rng = random.Random(seed)
        self.items = items                   # list of (nt0, nt1) for the work items of this CTA
        self.now = 0
        self.events = []                     # (time, seq, fn)
        self.seq = 0
        self.q_full = [MBar(1), MBar(1)]
        self.q_empty = [MBar(1), MBar(1)]
        self.k_full = [MBar(1) for _ in range(kv_stages)]
        self.k_empty = [MBar(2) for _ in range(kv_stages)]
        self.v_full = [MBar(1) for _ in range(kv_stages)]
        self.v_empty = [MBar(2) for _ in range(kv_stages)]
        self.s_full = [MBar(1), MBar(1)]
        self.s_empty = [MBar(4), MBar(4)]
        self.p_full = [MBar(4), MBar(4)]
        self.pv_done = [MBar(1), MBar(1)]
        self.b_full = [MBar(1) for _ in range(4)]
        self.b_empty = [MBar(4) for _ in range(4)]
        self.token = [NamedBar(), NamedBar()]
        self.kv_stages = kv_stages
        # buffers carry the tag of what they hold
        self.Q = [None, None]
        self.K = [None] * kv_stages
        self.V = [None] * kv_stages
        self.S = [None, None]
        self.P = [None, None]
        self.O = [None, None]                # (item, number of tiles accumulated)
        self.Bias = [None] * 4
        self.mma_busy_until = [0, 0]         # per issuing warp: in-order execution
        self.in_exp = [False, False]
        self.exp_overlaps = 0
        self.finished = set()

    # ---- event plumbing ----
    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + dt, self.seq, fn))

    def lat(self, lo, hi):
        return self.rng.randint(lo, hi)

    def tma(self, setter, bar):
        def land():
            setter()
            bar.arrive()
        self.at(self.lat(300, 1500), land)

    def mma(self, warp, dur, run, commits):
        """issue an MMA group on `warp`: runs after everything issued before it, `run()` at its end, then the commits"""
        start = max(self.now, self.mma_busy_until[warp])
        end = start + dur
        self.mma_busy_until[warp] = end

        def fin():
            run()
            for b in commits:
                b.arrive()
        self.seq += 1
        heapq.heappush(self.events, (end, self.seq, fin))

    # ---- agents: generators yielding a predicate to wait for, or an int delay ----
    def producer(self):
        Tk, W = 0, [0, 0]
        for it, nts in enumerate(self.items):
            nt_max = max(nts)
            if nt_max == 0:
                continue
            for i in (0, 1):
                if nts[i] == 0:
                    continue
                if W[i] > 0:
                    yield lambda i=i, par=(W[i] - 1) & 1: self.q_empty[i].done(par)
                self.tma(lambda i=i, it=it: self.Q.__setitem__(i, it), self.q_full[i])
                W[i] += 1
            for j in range(nt_max):
                s, par = Tk % self.kv_stages, ((Tk // self.kv_stages) & 1) ^ 1
                yield lambda s=s, par=par: self.k_empty[s].done(par)
                self.tma(lambda s=s, it=it, j=j: self.K.__setitem__(s, (it, j)), self.k_full[s])
                yield lambda s=s, par=par: self.v_empty[s].done(par)
                self.tma(lambda s=s, it=it, j=j: self.V.__setitem__(s, (it, j)), self.v_full[s])
                Tk += 1
        self.finished.add("producer")

    def bias_producer(self, i):
        I = 0
        for it, nts in enumerate(self.items):
            for i2 in range(2 * nts[i]):
                s = i * 2 + (I & 1)
                yield lambda s=s, par=((I >> 1) & 1) ^ 1: self.b_empty[s].done(par)
                self.tma(lambda s=s, it=it, i2=i2: self.Bias.__setitem__(s, (it, i2)), self.b_full[s])
                I += 1
        self.finished.add(f"bias{i}")

    def mma_warp(self, i):
        Tk = Ti = Wi = 0
        for it, nts in enumerate(self.items):
            nt, nt_max = nts[i], max(nts)

            def issue_s(tk, tile, last, it=it):
                s = tk % self.kv_stages

                def run():
                    assert self.Q[i] == it, ("S read a stale Q", i, it, self.Q[i])
                    assert self.K[s] == (it, tile), ("S read the wrong K tile", i, it, tile, self.K[s])
                    self.S[i] = (it, tile)
                commits = [self.s_full[i], self.k_empty[s]] + ([self.q_empty[i]] if last else [])
                self.mma(i, 256, run, commits)

            if nt > 0:
                yield lambda par=Wi & 1: self.q_full[i].done(par)
                yield lambda s=Tk % self.kv_stages, par=(Tk // self.kv_stages) & 1: self.k_full[s].done(par)
                if Ti > 0:
                    yield lambda par=(Ti - 1) & 1: self.s_empty[i].done(par)
                issue_s(Tk, 0, nt == 1)
                Wi += 1
            for j in range(nt_max):
                s = Tk % self.kv_stages
                if j < nt:
                    if j + 1 < nt:
                        tn = Tk + 1
                        yield lambda s2=tn % self.kv_stages, par=(tn // self.kv_stages) & 1: self.k_full[s2].done(par)
                        yield lambda par=Ti & 1: self.s_empty[i].done(par)
                        issue_s(tn, j + 1, j + 2 == nt)
                    yield lambda s=s, par=(Tk // self.kv_stages) & 1: self.v_full[s].done(par)
                    yield lambda par=Ti & 1: self.p_full[i].done(par)

                    def run(s=s, j=j, it=it):
                        assert self.V[s] == (it, j), ("PV read the wrong V tile", i, it, j, self.V[s])
                        assert self.P[i] == (it, j), ("PV read the wrong P", i, it, j, self.P[i])
                        if j == 0:
                            self.O[i] = (it, 1)
                        else:
                            assert self.O[i] == (it, j), ("O accumulates out of order", i, self.O[i], (it, j))
                            self.O[i] = (it, j + 1)
                    self.mma(i, 256, run, [self.pv_done[i], self.v_empty[s]])
                    Ti += 1
                else:
                    yield lambda s=s, par=(Tk // self.kv_stages) & 1: self.k_full[s].done(par)
                    yield lambda s=s, par=(Tk // self.kv_stages) & 1: self.v_full[s].done(par)
                    self.k_empty[s].arrive()
                    self.v_empty[s].arrive()
                Tk += 1
        self.finished.add(f"mma{i}")

    def softmax_wg(self, qi):
        T = I = 0
        if qi == 1:
            self.token[0].arrive()                               # pre-arm: the token starts with warpgroup 0
        for it, nts in enumerate(self.items):
            nt, nt_max = nts[qi], max(nts)
            for j in range(nt_max):
                if j >= nt:
                    yield lambda: self.token[qi].try_sync()
                    self.token[qi ^ 1].arrive()
                    continue
                for hh in (0, 1):                                # bias halves -> registers, then release both
                    s = qi * 2 + ((I + hh) & 1)
                    yield lambda s=s, par=((I + hh) >> 1) & 1: self.b_full[s].done(par)
                    assert self.Bias[s] == (it, 2 * j + hh), ("wrong bias half", qi, it, j, hh, self.Bias[s])
                for hh in (0, 1):
                    for _ in range(4):
                        self.b_empty[qi * 2 + ((I + hh) & 1)].arrive()
                I += 2
                yield lambda par=T & 1: self.s_full[qi].done(par)
                assert self.S[qi] == (it, j), ("softmax read the wrong S", qi, it, j, self.S[qi])
                yield self.lat(100, 300)                         # tcgen05.ld
                for _ in range(4):
                    self.s_empty[qi].arrive()
                yield self.lat(200, 900)                         # bias add, max
                yield lambda: self.token[qi].try_sync()
                if self.in_exp[qi ^ 1]:
                    self.exp_overlaps += 1
                self.in_exp[qi] = True
                yield self.lat(1000, 1200)                       # exp phase
                self.in_exp[qi] = False
                self.token[qi ^ 1].arrive()
                if j > 0:
                    yield lambda par=(T - 1) & 1: self.pv_done[qi].done(par)
                    assert self.O[qi] == (it, j), ("O rescale before the previous P V finished", qi, self.O[qi], (it, j))
                self.P[qi] = (it, j)
                yield self.lat(50, 150)                          # tcgen05.st
                for _ in range(4):
                    self.p_full[qi].arrive()
                T += 1
            if nt > 0:                                           # epilogue
                yield lambda par=(T - 1) & 1: self.pv_done[qi].done(par)
                assert self.O[qi] == (it, nt), ("epilogue read an incomplete O", qi, self.O[qi], (it, nt))
                yield self.lat(300, 1200)
        if qi == 0:
            yield lambda: self.token[0].try_sync()               # balance the last arrival of warpgroup 1
        self.finished.add(f"wg{qi}")

    def run(self):
        agents = {"producer": self.producer(), "bias0": self.bias_producer(0), "bias1": self.bias_producer(1),
                  "mma0": self.mma_warp(0), "mma1": self.mma_warp(1), "wg0": self.softmax_wg(0), "wg1": self.softmax_wg(1)}
        waiting = {}                                             # name -> predicate
        sleeping = {}                                            # name -> wake time

        def step(name):
            gen = agents[name]
            while True:
                try:
                    y = next(gen)
                except StopIteration:
                    agents.pop(name)
                    return
                if callable(y):
                    if y():
                        continue
                    waiting[name] = y
                    return
                sleeping[name] = self.now + y
                return

        for name in list(agents):
            step(name)
        guard = 0
        while agents:
            guard += 1
            assert guard < 2_000_000, "simulation does not terminate"
            progressed = False
            for name in list(waiting):
                if waiting[name]():
                    waiting.pop(name)
                    step(name)
                    progressed = True
            for name in list(sleeping):
                if sleeping[name] <= self.now:
                    sleeping.pop(name)
                    step(name)
                    progressed = True
            if progressed:
                continue
            nxt = []
            if self.events:
                nxt.append(self.events[0][0])
            if sleeping:
                nxt.append(min(sleeping.values()))
            assert nxt, ("deadlock", sorted(waiting), {k: v for k, v in vars(self).items() if k in ("S", "P", "O", "K", "V")})
            self.now = max(self.now, min(nxt))
            while self.events and self.events[0][0] <= self.now:
                _, _, fn = heapq.heappop(self.events)
                fn()
        while self.events:                                       # drain outstanding completions
            self.now, _, fn = heapq.heappop(self.events)
            fn()
        assert self.finished == {"producer", "bias0", "bias1", "mma0", "mma1", "wg0", "wg1"}
        assert self.token[0].arrivals == 0 and self.token[1].arrivals == 0, "unbalanced exp token"
        assert self.exp_overlaps == 0, "two exp phases ran at the same time"
        return self.now


ITEM_SETS = {
    "non_causal_8_tiles": [(8, 8)] * 3,
    "causal_pairs": [(7, 8), (5, 6), (1, 2), (3, 4)],                    # tile 1 sees one more key tile than tile 0
    "odd_number_of_query_blocks": [(8, 0), (8, 8), (8, 8)],             # the first pair hangs over the end: tile 1 has no work
    "single_tile_items": [(1, 1), (1, 1), (1, 2), (1, 1)],
    "causal_m_gt_n_with_empty_items": [(0, 0), (0, 1), (1, 2), (2, 3)],
    "long_then_short": [(32, 32), (1, 1), (2, 2)],
}


@pytest.mark.parametrize("name", sorted(ITEM_SETS))
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_protocol_has_no_deadlock_and_no_stale_reads(name, seed):
    Sim(ITEM_SETS[name], seed).run()


def test_token_alternation_costs_less_than_serialising_everything():
    """sanity of the model itself: with the token the tile-pair period is about max(2E, E + R), not 2(E + R)"""
    t = Sim([(8, 8)] * 4, 5).run()
    assert t < 32 * 2 * (1200 + 1200) * 0.75
