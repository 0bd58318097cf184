"""CPU model of the barrier protocol of the short-key (cross-attention) attention kernel, fmha_fwd_kernel<4, 2, 2>
(univid_b200/csrc/fmha_fwd_sm100.cuh): TMA producer, MMA warp, two softmax groups, mbarriers with the hardware's
phase-parity semantics, the tensor pipe as an in-order queue, TMA loads and stores as delayed completions.  Every
buffer carries a tag saying what it holds; each read asserts the tag, each overwrite asserts that the previous content
was consumed.  The actors below are line-by-line restatements of the kernel's control flow (waits, arrives, commits,
parities) -- including the cross-unit score issue of round 2 (the MMA warp issues the next unit's first Q K^T inside
the current unit's last step) -- run under randomised latencies over random unit sequences (0, 1, 2, 3, 4, 16 key
tiles per unit).  A protocol error shows up as a deadlock, a stale tag or a clobbered buffer.

This is host-side test infrastructure (no GPU, no CUDA): it checks the protocol's logic, the `-m gpu` tests check the
kernel.  The CTA pair is collapsed into one CTA: both CTAs run the same protocol against the leader's barriers.
"""
import heapq
import random

import pytest

STAGES = 4          # K/V ring slots of the pair variant
QBUFS = 2


class Deadlock(Exception):
    pass


class Bar:
    """mbarrier: `count` arrivals complete the current phase; try_wait.parity(p) is true iff the phase of parity p has
    completed, i.e. p differs from the parity of the current (incomplete) phase."""

    def __init__(self, name, count=1):
        self.name, self.count, self.k, self.pending = name, count, 0, 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending = 0
            self.k += 1

    def test(self, parity):
        return (self.k & 1) != parity


class Sim:
    def __init__(self, units, seed, hoist=True, hoist_min_steps=2):
        self.units, self.rng, self.hoist, self.hoist_min = units, random.Random(seed), hoist, hoist_min_steps
        self.now, self.events, self.seq = 0, [], 0
        self.q_full = [[Bar(f"q_full{b}{t}") for t in range(2)] for b in range(QBUFS)]
        self.q_empty = [[Bar(f"q_empty{b}{t}") for t in range(2)] for b in range(QBUFS)]
        self.kv_full = [Bar(f"kv_full{s}") for s in range(STAGES)]
        self.kv_empty = [Bar(f"kv_empty{s}") for s in range(STAGES)]
        self.s_full = [Bar(f"s_full{t}") for t in range(2)]
        self.p_full = [[Bar(f"p_full{t}{h}") for h in range(2)] for t in range(2)]
        self.pv_done = [Bar(f"pv_done{t}") for t in range(2)]
        self.o_full = [Bar(f"o_full{t}") for t in range(2)]
        # buffers: tag = what they hold, None = being written
        self.smem_q = [[None, None] for _ in range(QBUFS)]
        self.kv = [None] * STAGES
        self.S = [("free",), ("free",)]          # ("S", si, step) | ("held", si, step) once in the softmax registers
        self.p_tm = [None, None]                 # first half of P_t (lives in S_t's columns)
        self.p_tm_read = [True, True]
        self.p_sm = [None, None]                 # second half (shared-memory panel)
        self.p_sm_read = [True, True]
        self.O = [("drained",), ("drained",)]    # ("acc", si, steps accumulated)
        self.pipe_busy_until = 0                 # tensor pipe: ops run in issue order, one after the other
        self.done_units = [0, 0]
        self.stores_done = [True, True]          # the TMA store of group t has finished reading its staging tile

    # ---- time / events
    def at(self, delay, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + delay, self.seq, fn))

    def lat(self, lo, hi):
        return self.rng.randint(lo, hi)

    # ---- tensor pipe: ops execute in issue order; a commit is an op that arrives on a barrier
    def issue(self, duration, fn):
        start = max(self.now, self.pipe_busy_until)
        self.pipe_busy_until = start + duration
        self.seq += 1
        heapq.heappush(self.events, (self.pipe_busy_until, self.seq, fn))

    def commit(self, bar):
        self.issue(self.lat(1, 30), bar.arrive)

    # ---- actors (generators yielding ("wait", bar, parity) | ("delay", n))
    def producer(self):
        ring = 0
        for si, n in enumerate(self.units):
            qb, qpar = si % QBUFS, (si // QBUFS) & 1
            for t in range(2):
                yield ("wait", self.q_empty[qb][t], qpar ^ 1)
                self.smem_q[qb][t] = None

                def landed(qb=qb, t=t, si=si):
                    self.smem_q[qb][t] = ("Q", si, t)
                    self.q_full[qb][t].arrive()
                self.at(self.lat(50, 1500), landed)
            for i in range(2 * n):
                stage = ring % STAGES
                yield ("wait", self.kv_empty[stage], ((ring // STAGES) & 1) ^ 1)
                self.kv[stage] = None

                def landed(stage=stage, tag=("V" if i & 1 else "K", si, i >> 1)):
                    self.kv[stage] = tag
                    self.kv_full[stage].arrive()
                self.at(self.lat(50, 3000), landed)
                ring += 1
                yield ("delay", self.lat(1, 40))

    def mma(self):
        ring = gstep = 0
        scores_ahead = False

        def qk(si, qb, t, k_ring, step):
            def run():
                assert self.smem_q[qb][t] == ("Q", si, t), ("QK reads a stale Q", si, t, self.smem_q[qb][t])
                assert self.kv[k_ring % STAGES] == ("K", si, step), ("QK reads a stale K", si, step, self.kv[k_ring % STAGES])
                assert self.S[t][0] in ("free", "held"), ("QK overwrites unread scores", si, step, self.S[t])
                assert self.p_tm_read[t], ("QK overwrites the unread first half of P", si, step)
                self.S[t] = ("S", si, step)
            self.issue(self.lat(100, 600), run)

        def pv(si, t, v_ring, step, half):
            def run():
                assert self.kv[v_ring % STAGES] == ("V", si, step), ("PV reads a stale V", si, step, self.kv[v_ring % STAGES])
                if half == 0:
                    assert self.p_tm[t] == (si, step), ("PV reads a stale P (TMEM half)", si, step, self.p_tm[t])
                    self.p_tm_read[t] = True
                    if step == 0:
                        assert self.O[t] == ("drained",), ("PV overwrites an undrained O", si, self.O[t])
                        self.O[t] = ("acc", si, 0)
                else:
                    assert self.p_sm[t] == (si, step), ("PV reads a stale P (smem half)", si, step, self.p_sm[t])
                    self.p_sm_read[t] = True
                    assert self.O[t] == ("acc", si, step), ("PV accumulates out of order", si, step, self.O[t])
                    self.O[t] = ("acc", si, step + 1)
            self.issue(self.lat(50, 300), run)

        for si, n in enumerate(self.units):
            qb, qpar = si % QBUFS, (si // QBUFS) & 1
            primed, scores_ahead = scores_ahead, False
            nqb = nqpar = 0
            if self.hoist and n >= self.hoist_min and si + 1 < len(self.units):
                scores_ahead = self.units[si + 1] > 0
                nqb, nqpar = (si + 1) % QBUFS, ((si + 1) // QBUFS) & 1
            yield ("wait", self.q_full[qb][0], qpar)
            if n == 0:
                yield ("wait", self.q_full[qb][1], qpar)
                self.commit(self.o_full[0])
                self.commit(self.o_full[1])
                continue
            if not primed:
                yield ("wait", self.kv_full[ring % STAGES], (ring // STAGES) & 1)
                qk(si, qb, 0, ring, 0)
                self.commit(self.s_full[0])
                yield ("wait", self.q_full[qb][1], qpar)
                qk(si, qb, 1, ring, 0)
                self.commit(self.s_full[1])
                self.commit(self.kv_empty[ring % STAGES])
            for step in range(n):
                par = (gstep + step) & 1
                v_ring = ring + 2 * step + 1
                k_ring = v_ring + 1
                more = step + 1 < n
                ahead = more or scores_ahead
                yield ("wait", self.kv_full[v_ring % STAGES], (v_ring // STAGES) & 1)
                for t in range(2):
                    yield ("wait", self.p_full[t][0], par)
                    pv(si, t, v_ring, step, 0)
                    if ahead:
                        if t == 0:
                            yield ("wait", self.kv_full[k_ring % STAGES], (k_ring // STAGES) & 1)
                        if not more:
                            yield ("wait", self.q_full[nqb][t], nqpar)
                            qk(si + 1, nqb, t, k_ring, 0)
                        else:
                            qk(si, qb, t, k_ring, step + 1)
                        self.commit(self.s_full[t])
                    yield ("wait", self.p_full[t][1], par)
                    pv(si, t, v_ring, step, 1)
                    self.commit(self.pv_done[t])
                    if not more:
                        self.commit(self.o_full[t])
                    yield ("delay", self.lat(1, 60))
                self.commit(self.kv_empty[v_ring % STAGES])
                if ahead:
                    self.commit(self.kv_empty[k_ring % STAGES])
            ring += 2 * n
            gstep += n

    def softmax(self, t):
        gstep = 0
        pending = [-1]

        def release_pending():
            if pending[0] >= 0:
                yield ("waitfn", lambda: self.stores_done[t])
                self.q_empty[pending[0]][t].arrive()
                pending[0] = -1

        for si, n in enumerate(self.units):
            qb = si % QBUFS
            for step in range(n):
                par = (gstep + step) & 1
                yield ("wait", self.s_full[t], par)
                assert self.S[t] == ("S", si, step), ("softmax reads the wrong score tile", t, si, step, self.S[t])
                yield ("delay", self.lat(20, 200))                         # tcgen05.ld of the tile
                self.S[t] = ("held", si, step)
                if step > 0 and self.rng.random() < 0.3:                   # lazy rescale of O
                    yield ("wait", self.pv_done[t], par ^ 1)
                    assert self.O[t] == ("acc", si, step), ("rescale sees an O that is still accumulating", self.O[t])
                yield ("delay", self.lat(200, 1500))                       # first 64 keys
                assert self.p_tm_read[t], ("P (TMEM half) overwritten before the previous P V read it", t, si, step)
                self.p_tm[t], self.p_tm_read[t] = (si, step), False
                self.p_full[t][0].arrive()
                yield ("delay", self.lat(200, 1500))                       # second 64 keys
                if gstep + step > 0:
                    yield ("wait", self.pv_done[t], par ^ 1)
                assert self.p_sm_read[t], ("P panel overwritten before the previous P V read it", t, si, step)
                self.p_sm[t], self.p_sm_read[t] = (si, step), False
                self.p_full[t][1].arrive()
                if step == 0:
                    yield from release_pending()
            yield from release_pending()
            gstep += n
            yield ("wait", self.o_full[t], si & 1)
            if n > 0:
                assert self.O[t] == ("acc", si, n), ("epilogue reads an incomplete O", t, si, self.O[t])
            yield ("delay", self.lat(100, 1200))                           # drain O, scale, pack
            self.O[t] = ("drained",)
            assert self.smem_q[qb][t] == ("Q", si, t), ("O staging overwrites a live Q", t, si, self.smem_q[qb][t])
            self.smem_q[qb][t] = ("O", si, t)
            self.stores_done[t] = False

            def store_read_done(t=t):
                self.stores_done[t] = True
            self.at(self.lat(50, 2500), store_read_done)                   # TMA store has read the staging tile
            pending[0] = qb
            self.done_units[t] += 1
        yield from release_pending()

    # ---- scheduler
    def run(self):
        actors = {"producer": self.producer(), "mma": self.mma(), "sm0": self.softmax(0), "sm1": self.softmax(1)}
        blocked = {}
        ready = [(0, name) for name in actors]

        def step_actor(name):
            gen = actors[name]
            try:
                op = gen.send(None)
            except StopIteration:
                del actors[name]
                return
            if op[0] == "delay":
                self.at(op[1], lambda name=name: step_actor(name))
            else:
                blocked[name] = op
                poll(name)

        def poll(name):
            op = blocked.get(name)
            if op is None:
                return
            ok = op[1].test(op[2]) if op[0] == "wait" else op[1]()
            if ok:
                del blocked[name]
                self.at(self.lat(1, 20), lambda name=name: step_actor(name))

        for _, name in ready:
            step_actor(name)
        guard = 0
        while self.events:
            self.now, _, fn = heapq.heappop(self.events)
            fn()
            for name in list(blocked):
                poll(name)
            guard += 1
            assert guard < 5_000_000
        if actors:
            raise Deadlock({n: (op[1].name, op[2], op[1].k) if op[0] == "wait" else "store" for n, op in blocked.items()})
        assert self.done_units == [len(self.units)] * 2


def _sequences():
    rng = random.Random(7)
    fixed = [[4] * 12, [1] * 9, [2] * 9, [4, 1, 0, 3, 4, 4, 0, 0, 2, 1, 1, 4], [16, 16, 1, 16], [0, 0, 4, 0], [3]]
    return fixed + [[rng.choice([0, 1, 2, 3, 4, 4, 4, 16]) for _ in range(rng.randint(1, 14))] for _ in range(20)]


@pytest.mark.parametrize("hoist", [True, False])
def test_short_key_attention_protocol_is_deadlock_free_and_never_reads_stale_data(hoist):
    for units in _sequences():
        for seed in range(12):
            Sim(units, seed, hoist=hoist).run()


def test_scores_may_also_cross_the_boundary_behind_one_step_units():
    """The kernel issues the next unit's scores only from units of >= 2 steps (the buffer the next Q loads into is
    released one softmax step into the current unit).  The model says the rule is conservative -- the release does
    not wait on anything the MMA warp still has to do -- so it is a margin, not a necessity."""
    for units in _sequences():
        for seed in range(6):
            Sim(units, seed, hoist=True, hoist_min_steps=1).run()


def _without_waits(actor, prefix):
    """A Sim whose `actor` skips every wait on barriers whose name starts with `prefix`."""
    class Broken(Sim):
        pass

    def patched(self, *args):
        for op in getattr(Sim, actor)(self, *args):
            if op[0] == "wait" and op[1].name.startswith(prefix):
                continue
            yield op
    setattr(Broken, actor, patched)
    return Broken


@pytest.mark.parametrize("actor,prefix", [("mma", "q_full"), ("mma", "p_full00"), ("mma", "kv_full"), ("softmax", "pv_done"),
                                          ("softmax", "s_full"), ("producer", "kv_empty"), ("producer", "q_empty")])
def test_the_model_catches_a_broken_protocol(actor, prefix):
    """Negative controls: drop one kind of wait from one actor and some schedule reads stale data, clobbers a live
    buffer or deadlocks -- i.e. every wait of the protocol is load-bearing and the model's checks are live."""
    broken = _without_waits(actor, prefix)
    failures = 0
    for units in ([4] * 8, [2] * 8, [16, 16]):
        for seed in range(10):
            try:
                broken(units, seed).run()
            except (AssertionError, Deadlock):
                failures += 1
    assert failures > 0
