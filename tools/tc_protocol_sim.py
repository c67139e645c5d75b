"""Randomised interleaving simulator of the tensor-core samplers' intra-CTA hand-off protocol (tc_sampler.cu): eight row warps,
the MMA-issuing warp and the tensor pipe, with the kernel's mbarriers (count/phase/parity semantics), tensor-memory regions
(accumulators D0/D1, the A operand in halves) and named barriers.  It checks, under adversarial scheduling (some warps made
arbitrarily slow), that
  * no agent deadlocks (including parity aliasing: a waiter that falls two phases behind would block for ever),
  * every MMA group reads exactly the operand version it is meant to read (all writers done, no writer of the next version started),
  * an accumulator is only overwritten after all eight row warps have read the previous unit out of it,
  * a row warp never overwrites an A-operand region that an issued, not yet completed MMA group still reads.
No GPU, no product code: a model of the protocol, kept next to the kernel so that a change of the protocol can be tried here first.
    python tools/tc_protocol_sim.py [runs] [steps] [--old-a-ready] [--explicit-acc-waits] [--late-read]
--old-a-ready models the single 8-count bar_a_ready the kernel had before (both halves arriving on one barrier): the simulator
finds the early-release race that motivated the split within a few hundred schedules.
Default = the shipped issuer (late round 2; the group fusion is the f16x2 form, the three-product form takes a slot wait where
the fusion is — slot waits are not modelled): NO accumulator waits — every operand announcement of the row warps (operand halves,
x) follows their last read of the accumulator the announced MMAs overwrite — and fused issue groups (layer 1: first half | second
half + unit b; heads: first half | second half + the 64-column unit).  --explicit-acc-waits is the protocol before that (one
bar_acc_empty wait per unit, one group per wait).  --late-read is a negative control: the row warps announce the second operand
half BEFORE reading unit b's accumulator, which makes the implied release false; the model reports the overwrite.
Round 2 note: the kernel now publishes layer 0's output in FOUR pieces on four barriers (bar_h1_ready[0..3]) and layer 1's in two
(bar_a_ready[0..1]) — the same one-barrier-per-piece rule this model established for halves; the model still simulates the
two-halves form for both hand-offs, and keeps x inside the first operand half (the kernel moved it to the end of A_lo so that the
first h1 quarters can be stored while layer 0's second unit still reads x: disjoint columns, not a protocol matter)."""
import random
import sys


class Violation(Exception):
    pass


class MBar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count
        elif self.pending < 0:
            raise Violation(f"{self.name}: more arrivals than its count")

    def passed(self, parity):        # try_wait.parity(P): the phase with parity P has completed
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, steps, old_a_ready, rng, implied=True, late_read=False):
        self.T, self.old, self.rng = steps, old_a_ready, rng
        self.implied, self.late_read = implied and not old_a_ready, late_read
        self.acc_full = [MBar(f"acc_full{b}", 1) for b in range(2)]
        self.acc_empty = [MBar(f"acc_empty{b}", 8) for b in range(2)]
        self.a_ready = [MBar("a_ready0", 8), MBar("a_ready1", 8)]
        self.x_ready = MBar("x_ready", 4)
        self.named = {1: set(), 3: set()}          # named barriers (256 threads = all 8 row warps): arrival sets per generation
        self.named_gen = {1: 0, 3: 0}
        # tensor memory: A[(q, half, cs)] = content tag, D[b][(q, cs)] = tag of the unit it holds or None (consumed)
        self.A = {(q, h, cs): None for q in range(4) for h in range(2) for cs in range(2)}
        self.D = [{(q, cs): None for q in range(4) for cs in range(2)} for _ in range(2)]
        self.read_locks = {}                       # A region -> number of issued, incomplete MMA groups reading it
        self.pipe = []                             # issued work in order: ("mma", reads, expect, dwrite, unit) | ("commit", bar)
        self.trace = []

    # ---- tensor pipe: executes issued work strictly in order ----
    def pipe_step(self):
        kind, *rest = self.pipe.pop(0)
        if kind == "commit":
            rest[0].arrive()
            return
        reads, expect, b, unit, first = rest
        for r in reads:
            if self.A[r] != expect:
                raise Violation(f"MMA of {unit} reads A{r} = {self.A[r]}, expected {expect}")
            self.read_locks[r] -= 1
        if first:
            for k, v in self.D[b].items():
                if v is not None:
                    raise Violation(f"unit {unit} overwrites D{b}{k} which still holds unread {v}")
        for k in self.D[b]:
            self.D[b][k] = ("partial", unit)
        if rest[-1] == "last" or True:
            pass

    def issue(self, reads, expect, b, unit, first):
        for r in reads:
            self.read_locks[r] = self.read_locks.get(r, 0) + 1
        self.pipe.append(("mma", reads, expect, b, unit, first))

    def write_A(self, region, tag, who):
        if self.read_locks.get(region, 0) > 0:
            raise Violation(f"{who} overwrites A{region} (-> {tag}) while an issued MMA group still reads it")
        self.A[region] = tag

    def read_D(self, b, key, unit, who):
        v = self.D[b][key]
        if v != ("partial", unit):
            raise Violation(f"{who} reads D{b}{key} = {v}, expected unit {unit}")
        self.D[b][key] = None

    # ---- agents (generators yielding a wait predicate or None) ----
    def row_warp(self, w):
        q, cs = w & 3, w >> 2
        who = f"row{w}"
        u = 0
        if cs == 0:
            self.write_A((q, 0, 0), ("x", 0), who)
            yield None
            self.x_ready.arrive()
        for t in range(self.T):
            for layer in range(2):
                tag = ("h1", t) if layer == 0 else ("pf", t)
                b, n = u & 1, u >> 1                                   # unit a
                yield (lambda b=b, n=n: self.acc_full[b].passed(n & 1))
                self.read_D(b, (q, cs), (t, layer, "a"), who)
                yield None
                self.acc_empty[b].arrive()
                yield None
                u += 1
                b, n = u & 1, u >> 1                                   # unit b
                yield (lambda b=b, n=n: self.acc_full[b].passed(n & 1))
                self.write_A((q, 0, cs), tag, who)                     # first half of the new operand (held in registers until now)
                yield None
                if not self.old:                                       # (round 2: announced before the next accumulator load)
                    self.a_ready[0].arrive()
                    yield None
                if not self.late_read:
                    self.read_D(b, (q, cs), (t, layer, "b"), who)
                    yield None
                if self.old:
                    self.a_ready[0].arrive()
                    yield None
                self.acc_empty[b].arrive()
                yield None
                self.write_A((q, 1, cs), tag, who)
                yield None
                self.a_ready[0 if self.old else 1].arrive()
                yield None
                if self.late_read:                                     # negative control: the read follows the announcement
                    self.read_D(b, (q, cs), (t, layer, "b"), who)
                    yield None
                u += 1
            yield from self.named_sync(3, w)
            for unit in ("h128", "h64"):
                b, n = u & 1, u >> 1
                yield (lambda b=b, n=n: self.acc_full[b].passed(n & 1))
                self.read_D(b, (q, cs), (t, 2, unit), who)
                yield None
                self.acc_empty[b].arrive()
                yield None
                u += 1
            yield from self.named_sync(1, w)
            if cs == 1:
                for _ in range(self.rng.randrange(0, 4)):              # bias-table refresh etc.: arbitrary delay
                    yield None
                continue
            for _ in range(self.rng.randrange(0, 6)):                  # exchange, norm, grid word, update: arbitrary delay
                yield None
            if t + 1 < self.T:
                self.write_A((q, 0, 0), ("x", t + 1), who)
                yield None
                self.x_ready.arrive()
                yield None

    def named_sync(self, bid, w):
        gen = self.named_gen[bid]
        self.named[bid].add(w)
        if len(self.named[bid]) == 8:
            self.named[bid] = set()
            self.named_gen[bid] += 1
        yield (lambda: self.named_gen[bid] > gen)

    def mma_warp(self):
        u = xr = ar = 0
        allq = range(4)
        first_half = [(q, 0, cs) for q in allq for cs in range(2)]
        second_half = [(q, 1, cs) for q in allq for cs in range(2)]
        for t in range(self.T):
            yield (lambda xr=xr: self.x_ready.passed(xr & 1))
            xr += 1
            if self.implied:
                # shipped f16x2 issuer: no accumulator waits, fused groups
                for half in ("a", "b"):                                # layer 0 (x implies both accumulators free)
                    b = u & 1
                    self.issue([(q, 0, 0) for q in allq], ("x", t), b, (t, 0, half), True)
                    self.pipe.append(("commit", self.acc_full[b]))
                    u += 1
                yield None
                for layer, names, expect in ((1, ("a", "b"), ("h1", t)), (2, ("h128", "h64"), ("pf", t))):
                    ba, bb = u & 1, (u + 1) & 1
                    yield (lambda ar=ar: self.a_ready[0].passed((ar >> 1) & 1))
                    ar += 1
                    self.issue(first_half, expect, ba, (t, layer, names[0]), True)
                    yield None
                    yield (lambda ar=ar: self.a_ready[1].passed((ar >> 1) & 1))
                    ar += 1
                    self.issue(second_half, expect, ba, (t, layer, names[0]), False)
                    self.pipe.append(("commit", self.acc_full[ba]))
                    self.issue(first_half + second_half, expect, bb, (t, layer, names[1]), True)   # same issue group, no wait
                    self.pipe.append(("commit", self.acc_full[bb]))
                    yield None
                    u += 2
                continue
            for half in ("a", "b"):                                    # layer 0: both units read x
                b, n = u & 1, u >> 1
                yield (lambda b=b, n=n: self.acc_empty[b].passed((n & 1) ^ 1))
                self.issue([(q, 0, 0) for q in allq], ("x", t), b, (t, 0, half), True)
                self.pipe.append(("commit", self.acc_full[b]))
                yield None
                u += 1
            for unit in range(4):
                split = unit in (0, 2)
                layer, name = (1, "ab"[unit]) if unit < 2 else (2, ("h128", "h64")[unit - 2])
                expect = ("h1", t) if unit < 2 else ("pf", t)
                b, n = u & 1, u >> 1
                groups = 2 if split else 1
                for g in range(groups):
                    if split:
                        if self.old:
                            yield (lambda ar=ar: self.a_ready[0].passed(ar & 1))
                        else:
                            yield (lambda ar=ar, g=g: self.a_ready[g].passed((ar >> 1) & 1))
                        ar += 1
                    if g == 0:
                        yield (lambda b=b, n=n: self.acc_empty[b].passed((n & 1) ^ 1))
                    reads = (first_half if g == 0 else second_half) if split else first_half + second_half
                    self.issue(reads, expect, b, (t, layer, name), g == 0)
                    yield None
                self.pipe.append(("commit", self.acc_full[b]))
                yield None
                u += 1

    # ---- scheduler ----
    def run(self):
        agents = {f"row{w}": self.row_warp(w) for w in range(8)}
        agents["mma"] = self.mma_warp()
        waiting = {k: None for k in agents}
        weights = {k: self.rng.choice([1, 1, 1, 5, 25]) for k in agents}      # some agents much slower than others
        pipe_w = self.rng.choice([1, 3, 10])
        live = set(agents)
        while live:
            runnable = [k for k in live if waiting[k] is None or waiting[k]()]
            cands = runnable + (["pipe"] if self.pipe else [])
            if not cands:
                st = {k: "blocked" for k in live}
                raise Violation(f"deadlock: {sorted(live)} blocked, phases "
                                f"acc_full {[b.phase for b in self.acc_full]} acc_empty {[b.phase for b in self.acc_empty]} "
                                f"a_ready {[b.phase for b in self.a_ready]} x_ready {self.x_ready.phase}")
            k = self.rng.choices(cands, weights=[(1.0 / pipe_w if c == "pipe" else 1.0 / weights[c]) for c in cands])[0]
            if k == "pipe":
                self.pipe_step()
                continue
            try:
                waiting[k] = next(agents[k])
            except StopIteration:
                live.discard(k)
        while self.pipe:
            self.pipe_step()


class TxBar:
    """mbarrier with a transaction count: the phase completes when the pending arrivals AND the outstanding bytes reach zero."""

    def __init__(self, name):
        self.name, self.pending, self.tx, self.phase = name, 1, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = 1

    def arrive_expect_tx(self, n):
        if self.pending == 0:
            raise Violation(f"{self.name}: second arrival inside one phase")
        self.tx += n
        self.pending -= 1
        self._check()

    def complete_tx(self, n):
        self.tx -= n
        self._check()

    def passed(self, parity):
        return (self.phase & 1) != parity


class TeamSim:
    """The inter-CTA side: `teams` tile teams of four ranks.  Per step every rank sends its partial score to its three peers
    (st.async: asynchronous delivery, completes bytes on the RECEIVER's bar_mail[step & 1]), waits for its own mailbox, reads the
    four slots, the team leader adds the tile norm to the step's grid word, every rank polls the word, updates.  Checked: no
    deadlock, every rank reads exactly this step's partials from every peer (no stale or future mail), the grid word of a step is
    complete when read."""

    def __init__(self, steps, teams, rng):
        self.T, self.n, self.rng = steps, teams, rng
        self.bar = {(tm, r, par): TxBar(f"mail[{tm},{r},{par}]") for tm in range(teams) for r in range(4) for par in range(2)}
        self.mail = {(tm, r, par, src): None for tm in range(teams) for r in range(4) for par in range(2) for src in range(4)}
        self.net = []                 # in flight: (team, dst, par, src, step)
        self.word = {}

    def rank(self, tm, r):
        for t in range(self.T):
            par, ph = t & 1, (t >> 1) & 1
            for _ in range(self.rng.randrange(0, 5)):          # the score evaluation: arbitrary time
                yield None
            self.bar[(tm, r, par)].arrive_expect_tx(3)
            yield None
            self.mail[(tm, r, par, r)] = t                     # own slot
            for d in range(1, 4):
                self.net.append((tm, (r + d) & 3, par, r, t))
                yield None
            yield (lambda par=par, ph=ph: self.bar[(tm, r, par)].passed(ph))
            for src in range(4):
                if self.mail[(tm, r, par, src)] != t:
                    raise Violation(f"team {tm} rank {r} step {t}: mailbox slot of rank {src} holds step {self.mail[(tm, r, par, src)]}")
            yield None
            if r == 0:
                self.word[t] = self.word.get(t, 0) + 1
                yield None
            yield (lambda t=t: self.word.get(t, 0) >= self.n)
            if self.word[t] != self.n:
                raise Violation(f"grid word of step {t} read as {self.word[t]} of {self.n}")
            for _ in range(self.rng.randrange(0, 3)):          # update
                yield None

    def deliver(self):
        i = self.rng.randrange(len(self.net))                  # st.async completions are not ordered between senders
        tm, dst, par, src, t = self.net.pop(i)
        self.mail[(tm, dst, par, src)] = t
        self.bar[(tm, dst, par)].complete_tx(1)

    def run(self):
        agents = {(tm, r): self.rank(tm, r) for tm in range(self.n) for r in range(4)}
        waiting = {k: None for k in agents}
        weights = {k: self.rng.choice([1, 1, 1, 5, 25]) for k in agents}
        net_w = self.rng.choice([1, 3, 20])
        live = set(agents)
        while live:
            runnable = [k for k in live if waiting[k] is None or waiting[k]()]
            cands = runnable + (["net"] if self.net else [])
            if not cands:
                raise Violation(f"deadlock: ranks {sorted(live)} blocked")
            k = self.rng.choices(cands, weights=[(1.0 / net_w if c == "net" else 1.0 / weights[c]) for c in cands])[0]
            if k == "net":
                self.deliver()
                continue
            try:
                waiting[k] = next(agents[k])
            except StopIteration:
                live.discard(k)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    runs = int(args[0]) if args else 2000
    steps = int(args[1]) if len(args) > 1 else 3
    old = "--old-a-ready" in sys.argv
    implied, late = "--explicit-acc-waits" not in sys.argv, "--late-read" in sys.argv
    bad = 0
    for seed in range(runs):
        try:
            Sim(steps, old, random.Random(seed), implied=implied, late_read=late).run()
        except Violation as e:
            bad += 1
            if bad <= 3:
                print(f"schedule {seed}: {e}")
    name = "single a_ready barrier (old)" if old else ("one a_ready barrier per half, " + ("implied accumulator release + fused groups (shipped)"
                                                                                        if implied else "explicit accumulator waits"))
    print(f"{runs} random schedules x {steps} steps, protocol = {name}{', announcements before the read (negative control)' if late else ''}: "
          f"{bad} violations")
    bad_team = 0
    for seed in range(runs):
        try:
            TeamSim(steps + 3, 3, random.Random(seed)).run()
        except Violation as e:
            bad_team += 1
            if bad_team <= 3:
                print(f"team schedule {seed}: {e}")
    print(f"{runs} random schedules x {steps + 3} steps, team exchange + grid word (3 teams of 4 ranks): {bad_team} violations")
    return 1 if ((bad and not old and not late) or bad_team) else 0


if __name__ == "__main__":
    sys.exit(main())
