"""Discrete-event check of the synchronisation protocol of csrc/attention_pipe.cu (persistent ping-pong attention).

The kernel could not be run on hardware when it was written, so its mbarrier protocol is transcribed here agent by agent
(TMA producer, two MMA issuer chains, the cls warp, 2 x 4 row warps) and executed under thousands of random interleavings
with random completion delays for the asynchronous engines (TMA transactions, tcgen05 commits).  Checked:
  * no deadlock (every agent finishes its job list);
  * no mbarrier phase aliasing (a waiter never misses a phase because the barrier ran two phases ahead);
  * no buffer hazard: every read sees the data of the job it expects, and no buffer is overwritten while a reader of
    the previous job may still touch it (Q tiles, K/V stages, P buffers, the S/O TMEM regions).

    python tools/sim_attn_pipe.py [--seeds 2000] [--jobs 5]

mbarrier model: `pending` arrivals left in the current phase; completing a phase flips `phase` and re-arms the count.
wait(parity) succeeds iff the phase with that parity has completed, i.e. `phase != parity` — and raises if the barrier
is found TWO completions ahead of the waiter (aliasing), which real hardware would turn into a hang or a stale read."""
import argparse
import random


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase, self.completed = name, count, count, 0, 0

    def arrive(self, n=1):
        self.pending -= n
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending, self.phase, self.completed = self.count, self.phase ^ 1, self.completed + 1


class Sim:
    def __init__(self, njobs, prefix, rng, broken=""):
        self.njobs, self.prefix, self.rng, self.broken = njobs, prefix, rng, broken
        B = lambda n, c: Bar(n, c)
        self.kv_full = [B(f"kv_full{s}", 1) for s in range(2)]
        self.kv_empty = [B(f"kv_empty{s}", 3 if prefix else 2) for s in range(2)]
        self.q_full = B("q_full", 1)
        self.q_empty = B("q_empty", 2 + 8)            # two S chains + 8 row warps (each warp stands for its 32 threads)
        self.s_full = [B(f"s_full{w}", 1) for w in range(2)]
        self.s_empty = [B(f"s_empty{w}", 4) for w in range(2)]
        self.p_full = [[B(f"p_full{w}{h}", 4) for h in range(2)] for w in range(2)]
        self.pv_done = [[B(f"pv_done{w}{h}", 1) for h in range(2)] for w in range(2)]
        # buffer contents: which job's data they hold (None = nothing valid), and who may still be reading
        self.kv = [None, None]
        self.q = None
        self.s_tmem = [None, None]          # job whose S_w / O_w lives in the TMEM region
        self.o_ready = [None, None]
        self.p = [[None, None], [None, None]]  # p[w] = (job, half) currently in warpgroup w's P buffer
        self.pending_async = []             # (due_time, fn): TMA transaction completions, tcgen05 commit arrivals
        self.time = 0
        self.waits = {}                     # agent -> number of completed waits per barrier (for the aliasing check)

    # asynchronous engines -------------------------------------------------------------------------------------
    def later(self, fn, lo=1, hi=30):
        self.pending_async.append((self.time + self.rng.randint(lo, hi), fn))

    def pump(self):
        due = [x for x in self.pending_async if x[0] <= self.time]
        self.pending_async = [x for x in self.pending_async if x[0] > self.time]
        for _, fn in sorted(due, key=lambda x: x[0]):
            fn()

    def wait(self, agent, bar, parity):
        """generator: blocks until the phase with `parity` has completed; detects aliasing via completion counts"""
        key = (agent, bar.name)
        while bar.phase == parity:
            done = self.waits.get(key, 0)
            assert bar.completed <= done + 1, f"{agent} waiting on {bar.name}: barrier is {bar.completed - done} phases ahead (aliasing)"
            yield
        done = self.waits.get(key, 0)
        assert bar.completed == done + 1, f"{agent} on {bar.name}: completed={bar.completed}, waiter consumed {done} (phase aliasing)"
        self.waits[key] = done + 1

    # agents (transcribed from the kernel) -----------------------------------------------------------------------
    def tma(self):
        for it in range(self.njobs):
            s, u = it & 1, it >> 1
            if u >= 1 and self.broken != "no_kv_empty_wait":
                yield from self.wait("tma", self.kv_empty[s], (u - 1) & 1)
            assert self.kv[s] is None or self.kv[s] == it - 2, f"K/V stage {s} overwritten while holding job {self.kv[s]}"
            self.kv[s] = ("loading", it)

            def kv_done(s=s, it=it):
                self.kv[s] = it
                self.kv_full[s].arrive()
            self.later(kv_done)
            if it >= 1 and self.broken != "no_q_empty_wait":
                yield from self.wait("tma", self.q_empty, (it - 1) & 1)
            self.q = ("loading", it)

            def q_done(it=it):
                self.q = it
                self.q_full.arrive()
            self.later(q_done)
            yield

    def issuer(self, w):
        me = f"mma{w}"
        commits = []                         # in-order completion of this thread's MMAs

        def commit(fn):                      # tcgen05.commit: arrives after ALL earlier MMAs of this thread completed
            commits.append(fn)
            due = max([t for t, _ in self.pending_async if getattr(_, "issuer", None) == me] + [self.time]) + self.rng.randint(1, 25)

            def fire(fn=fn):
                fn()
            fire.issuer = me
            self.pending_async.append((due, fire))
        for it in range(self.njobs):
            s, u = it & 1, it >> 1
            yield from self.wait(me, self.q_full, it & 1)
            yield from self.wait(me, self.kv_full[s], u & 1)
            if it >= 1 and self.broken != "no_s_empty_wait":
                yield from self.wait(me, self.s_empty[w], (it - 1) & 1)
            assert self.q == it and self.kv[s] == it, f"{me}: S MMA of job {it} reads Q={self.q} KV={self.kv[s]}"
            assert self.s_tmem[w] in (None, ("drained", it - 1)), f"{me}: S_{w} of job {it} overwrites live TMEM {self.s_tmem[w]}"
            self.s_tmem[w] = ("S", it)

            def s_done(w=w, s=s, it=it):     # the S MMA reads Q_w and K until it completes
                assert self.q == it and self.kv[s] == it, f"{me}: S MMA of job {it} still running, Q={self.q} KV={self.kv[s]}"
                self.s_full[w].arrive()
            commit(s_done)
            commit(lambda: self.q_empty.arrive())
            for half in range(2):
                yield from self.wait(me, self.p_full[w][half], it & 1)
                assert self.p[w] == [it, half], f"{me}: P.V job {it} half {half} reads P buffer {self.p[w]}"
                assert self.kv[s] == it, f"{me}: P.V of job {it} reads K/V stage holding {self.kv[s]}"

                def pv(w=w, half=half, it=it, s=s):
                    assert self.p[w] == [it, half] and self.kv[s] == it, \
                        f"{me}: P.V job {it} half {half} still running, P={self.p[w]} KV={self.kv[s]}"
                    if half == 1:
                        self.o_ready[w] = it
                    self.pv_done[w][half].arrive()
                commit(pv)
            commit(lambda s=s: self.kv_empty[s].arrive())
            yield

    def cls(self):
        if not self.prefix:
            return
        for it in range(self.njobs):
            s, u = it & 1, it >> 1
            yield from self.wait("cls", self.kv_full[s], u & 1)
            for _ in range(self.rng.randint(1, 6)):      # scores + P.V over the smem K/V stage
                assert self.kv[s] == it, f"cls warp of job {it} reads K/V stage holding {self.kv[s]}"
                yield
            self.kv_empty[s].arrive()

    def row_warp(self, w, q):
        me = f"row{w}{q}"
        for it in range(self.njobs):
            ph = it & 1
            yield from self.wait(me, self.q_full, ph)
            for _ in range(self.rng.randint(0, 3)):
                assert self.q == it, f"{me}: reads q row of job {it} but the buffer holds {self.q}"
                yield
            self.q_empty.arrive()
            yield from self.wait(me, self.s_full[w], ph)
            for _ in range(self.rng.randint(1, 4)):      # pass 1: the whole score row
                assert self.s_tmem[w] == ("S", it), f"{me}: pass 1 of job {it} reads TMEM {self.s_tmem[w]}"
                yield
            for half in range(2):
                if half == 1 and self.broken != "no_pv0_wait":
                    yield from self.wait(me, self.pv_done[w][0], ph)
                for _ in range(self.rng.randint(1, 4)):  # pass 2 of this half: reads S cols of the half, writes the P buffer
                    assert self.s_tmem[w][1] == it, f"{me}: pass 2 of job {it} reads TMEM {self.s_tmem[w]}"
                    yield
                # P buffer write: the previous contents must have been consumed (P.V of the previous half / job committed)
                prev = self.p[w]
                if prev != [it, half]:
                    if half == 0:
                        assert prev == [None, None] or (prev == [it - 1, 1] and self.pv_done[w][1].completed >= it), \
                            f"{me}: P buffer {prev} overwritten before its P.V completed"
                    else:
                        assert prev == [it, 0] and self.pv_done[w][0].completed >= it + 1, f"{me}: P half 0 overwritten too early ({prev})"
                    self.p[w] = [it, half]
                self.p_full[w][half].arrive()
            yield from self.wait(me, self.pv_done[w][1], ph)
            assert self.o_ready[w] == it, f"{me}: epilogue of job {it} reads O of job {self.o_ready[w]}"
            yield
            self.s_empty[w].arrive()
            if self.s_empty[w].pending == self.s_empty[w].count:     # the last warp of the group drained O_w
                self.s_tmem[w] = ("drained", it)
            for _ in range(self.rng.randint(0, 3)):      # global stores from registers
                yield

    def run(self):
        agents = {"tma": self.tma(), "mma0": self.issuer(0), "mma1": self.issuer(1), "cls": self.cls()}
        for w in range(2):
            for q in range(4):
                agents[f"row{w}{q}"] = self.row_warp(w, q)
        live = dict(agents)
        idle_rounds = 0
        while live:
            self.time += 1
            self.pump()
            name = self.rng.choice(list(live))
            before = (self.time, len(self.pending_async))
            try:
                next(live[name])
            except StopIteration:
                del live[name]
            idle_rounds = idle_rounds + 1 if self.time > 200000 else 0
            assert self.time < 200000, f"deadlock / livelock: still running {sorted(live)} at t={self.time}"
        # drain the async queue (final commits)
        while self.pending_async:
            self.time += 1
            self.pump()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=2000)
    ap.add_argument("--jobs", type=int, default=5)
    ap.add_argument("--selftest", action="store_true", help="also check that deliberately broken protocols ARE caught")
    a = ap.parse_args()
    if a.selftest:
        for broken in ("no_q_empty_wait", "no_kv_empty_wait", "no_s_empty_wait", "no_pv0_wait"):
            caught = 0
            for seed in range(200):
                try:
                    Sim(4, 1, random.Random(seed), broken=broken).run()
                except AssertionError:
                    caught += 1
            print(f"selftest {broken}: caught in {caught}/200 schedules")
            assert caught > 0, f"the simulator does not notice a protocol without {broken}"
    n = 0
    for seed in range(a.seeds):
        for prefix in (0, 1):
            for njobs in range(1, a.jobs + 1):
                Sim(njobs, prefix, random.Random(seed * 131 + prefix * 17 + njobs)).run()
                n += 1
    print(f"attention_pipe protocol: {n} random schedules, no deadlock, no phase aliasing, no buffer hazard")


if __name__ == "__main__":
    main()
