"""Randomised interleaving model of the synchronisation protocol of csrc/attn_fwd2.cuh / attn_fwd3.cuh (the opt-in
forward kernels; attn_fwd2 ran correctly on the device the first time it was launched).

The kernel could not be run on a device when it was written, so the part that cannot be checked by compiling — who
waits for whom, on which mbarrier phase, and which buffer may be overwritten when — is restated here statement by
statement and executed under random schedules:

  * issuer thread, four softmax warps, the tensor pipe (an in-order FIFO: MMAs and commits) and the TMA unit are
    coroutines; a seeded scheduler picks any runnable one at every step, and asynchronous units take random delays;
  * mbarriers are modelled with phase counters; a wait names the phase it expects — waking up on a different phase,
    or waiting for a parity that can no longer be observed (the hardware hang), is an error;
  * every buffer carries a tag of what it holds (K_t / V_t in the three rotating smem buffers, P_t per warp, S_t in
    TMEM, the number of PV products folded into O and the rescales applied by every warp); every consumer asserts the
    tag it needs at the start AND at the end of its operation, so write-after-read and read-before-write hazards
    surface as assertion failures under some schedule.

`run(n_tiles, seed)` returns the number of scheduler steps; tests/test_attn_v2_protocol_cpu.py runs many seeds.
The model follows the kernel's statement order; keep the two in step when either changes."""
from __future__ import annotations

import random


class Deadlock(Exception):
    pass


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier was initialised for"
        if self.pending == 0:
            self.completed += 1
            self.pending = self.count

    def ready(self, phase):
        """mbarrier.try_wait.parity(phase & 1): true iff the most recently completed phase has that parity"""
        last = self.completed - 1                      # -1: nothing completed yet == "previous phase" parity 1
        return (last & 1) == (phase & 1)


class Sim:
    def __init__(self, n_tiles, seed, skip=(), n_warps=4):
        """`skip`: names of waits to leave out ("sfree", "o_before_k", "o_before_p", "s_before_v", "pair") — used by
        the test to prove that the model notices a broken protocol.  n_warps = 4 models attn_fwd2 (one thread per
        row), n_warps = 8 models attn_fwd3: warps w and w + 4 share rows and exchange the row max through a
        two-slot (tile parity) buffer and a named barrier."""
        self.n, self.rng, self.skip, self.W = n_tiles, random.Random(seed), set(skip), n_warps
        self.bar = {k: Barrier(k, c) for k, c in (("q", 1), ("k0", 1), ("k1", 1), ("v0", 1), ("v1", 1), ("s", 1),
                                                  ("sfree", n_warps), ("p", n_warps), ("o", 1))}
        for pr in range(4):
            self.bar[f"pair{pr}"] = Barrier(f"pair{pr}", 2)      # hardware named barrier: counts, no parity aliasing
        self.xbuf = {}                                 # (parity, warp) -> tile whose max that slot holds
        self.kv = [None, None, None]                   # tags of the three rotating buffers
        self.q_loaded = False
        self.p_tag = [None] * n_warps                  # per softmax warp: tile whose P rows it wrote last
        self.s_tag = None                              # tile whose scores sit in TMEM
        self.s_read = [None] * n_warps                 # per warp: last tile copied to registers
        self.o_products = 0                            # number of P_t V_t folded into O
        self.o_rescaled = [0] * n_warps                # per warp: rescales applied (tile index of the last one)
        self.pipe = []                                 # in-order tensor pipe: ("mma", kind, t) / ("commit", bar)
        self.pipe_busy = None
        self.tma = []                                  # [remaining delay, action]
        self.steps = 0

    # ---- asynchronous units -------------------------------------------------------------------------------
    def tma_issue(self, buf, tag, bar):
        # the buffer may not be read by anything still queued in or running on the tensor pipe
        for op in self.pipe + ([self.pipe_busy] if self.pipe_busy else []):
            if op[0] == "mma":
                assert self._operand_buffer(op) != buf, f"TMA {tag} overwrites buffer {buf} still needed by {op}"
        self.kv[buf] = ("loading", tag)

        def done():
            self.kv[buf] = tag
            self.bar[bar].arrive()
        self.tma.append([self.rng.randint(1, 40), done])

    def _operand_buffer(self, op):
        kind, t = op[1], op[2]
        if self.W == 8:                                # attn_fwd3: K_0 -> 0, K_t -> 2, V_t -> t odd ? 0 : 1
            return (0 if t == 0 else 2) if kind == "qk" else (0 if t & 1 else 1)
        return (2 * t) % 3 if kind == "qk" else (2 * t + 1) % 3

    def _mma_check(self, op, starting):
        kind, t = op[1], op[2]
        buf = self._operand_buffer(op)
        if kind == "qk":
            assert self.q_loaded and self.kv[buf] == ("K", t), f"QK_{t}: buffer {buf} holds {self.kv[buf]}"
            if t > 0:
                assert all(r == t - 1 for r in self.s_read), f"QK_{t} overwrites S_{t - 1} before it was copied: {self.s_read}"
        else:
            assert self.kv[buf] == ("V", t), f"PV_{t}: buffer {buf} holds {self.kv[buf]}"
            assert all(x == t for x in self.p_tag), f"PV_{t}: P buffer holds {self.p_tag}"
            assert self.o_products == t, f"PV_{t}: O holds {self.o_products} products"
            if t > 0:
                assert all(x == t for x in self.o_rescaled), f"PV_{t}: rescale state {self.o_rescaled}"

    def unit_step(self):
        """advance the TMA unit and the tensor pipe by one (random) step"""
        if self.tma and self.rng.random() < 0.7:
            item = self.rng.choice(self.tma)
            item[0] -= 1
            if item[0] <= 0:
                self.tma.remove(item)
                item[1]()
        if self.pipe_busy is None and self.pipe and self.rng.random() < 0.6:
            op = self.pipe.pop(0)
            if op[0] == "commit":
                self.bar[op[1]].arrive()
            else:
                self._mma_check(op, True)
                self.pipe_busy = [op[0], op[1], op[2], self.rng.randint(1, 12)]
        elif self.pipe_busy is not None:
            self.pipe_busy[3] -= 1
            if self.pipe_busy[3] <= 0:
                op = tuple(self.pipe_busy[:3])
                self._mma_check(op, False)
                if op[1] == "qk":
                    self.s_tag = op[2]
                else:
                    self.o_products += 1
                self.pipe_busy = None

    # ---- threads (generators yield ("wait", barrier, phase) or None for a plain scheduling point) -----------
    def issuer(self):
        if self.W == 8:
            yield from self.issuer_v3()
            return
        n = self.n
        # prologue (before the CTA barrier): Q, K_0, V_0 and K_1
        self.tma.append([self.rng.randint(1, 40), lambda: (setattr(self, "q_loaded", True), self.bar["q"].arrive())])
        self.tma_issue(0, ("K", 0), "k0")
        self.tma_issue(1, ("V", 0), "v0")
        if n > 1:
            self.tma_issue(2, ("K", 1), "k1")
        yield None
        yield ("wait", "q", 0)
        yield ("wait", "k0", 0)
        self.pipe += [("mma", "qk", 0), ("commit", "s")]
        for t in range(n):
            if t + 1 < n:
                if "s_before_v" not in self.skip:
                    yield ("wait", "s", t)
                self.tma_issue((2 * t) % 3, ("V", t + 1), f"v{(t + 1) & 1}")
                if t >= 1:
                    if "o_before_k" not in self.skip:
                        yield ("wait", "o", t - 1)
                    self.tma_issue((2 * t + 2) % 3, ("K", t + 1), f"k{(t + 1) & 1}")
                yield ("wait", f"k{(t + 1) & 1}", (t + 1) >> 1)
                if "sfree" not in self.skip:
                    yield ("wait", "sfree", t)
                self.pipe += [("mma", "qk", t + 1), ("commit", "s")]
            yield ("wait", f"v{t & 1}", t >> 1)
            yield ("wait", "p", t)
            self.pipe += [("mma", "pv", t), ("commit", "o")]
            yield None

    def issuer_v3(self):
        """attn_fwd3: K_{t+2} is fetched into buffer 2 as soon as S_{t+1} has been computed from it; V_{t+1} takes
        V_{t-1}'s (or K_0's) buffer"""
        n = self.n
        self.tma.append([self.rng.randint(1, 40), lambda: (setattr(self, "q_loaded", True), self.bar["q"].arrive())])
        self.tma_issue(0, ("K", 0), "k0")
        self.tma_issue(1, ("V", 0), "v0")
        if n > 1:
            self.tma_issue(2, ("K", 1), "k1")
        yield None
        yield ("wait", "q", 0)
        yield ("wait", "k0", 0)
        self.pipe += [("mma", "qk", 0), ("commit", "s")]
        for t in range(n):
            if t + 1 < n:
                yield ("wait", f"k{(t + 1) & 1}", (t + 1) >> 1)
                if "sfree" not in self.skip:
                    yield ("wait", "sfree", t)
                self.pipe += [("mma", "qk", t + 1), ("commit", "s")]
                if t + 2 < n:
                    if "s_before_k" not in self.skip:
                        yield ("wait", "s", t + 1)
                    self.tma_issue(2, ("K", t + 2), f"k{t & 1}")
                if t >= 1 and "o_before_v" not in self.skip:
                    yield ("wait", "o", t - 1)
                self.tma_issue(0 if (t + 1) & 1 else 1, ("V", t + 1), f"v{(t + 1) & 1}")
            yield ("wait", f"v{t & 1}", t >> 1)
            yield ("wait", "p", t)
            self.pipe += [("mma", "pv", t), ("commit", "o")]
            yield None

    def softmax(self, w):
        n = self.n
        for t in range(n):
            yield ("wait", "s", t)
            assert self.s_tag == t, f"warp {w}: expected S_{t} in TMEM, found S_{self.s_tag}"
            yield None                                     # tcgen05.ld in flight
            assert self.s_tag == t, f"warp {w}: S_{t} overwritten while it was being read"
            self.s_read[w] = t
            self.bar["sfree"].arrive()
            yield None                                     # bias, max
            if self.W == 8:
                yield from self._pair_exchange(w, t, t & 1)
            if t > 0:
                if "o_before_p" not in self.skip:
                    yield ("wait", "o", t - 1)
                assert self.o_products == t, f"warp {w}: rescale for tile {t} sees {self.o_products} products"
                busy = self.pipe_busy is not None and self.pipe_busy[1] == "pv"
                assert not busy, f"warp {w}: rescales O while a PV product is running"
                yield None                                 # ld, mul, st
                assert self.o_products == t
                self.o_rescaled[w] = t
            yield None                                     # exp2, dropout, pack
            busy = self.pipe_busy is not None and self.pipe_busy[1] == "pv"
            assert not busy, f"warp {w}: writes P_{t} while a PV product reads the buffer"
            assert all(op[:2] != ("mma", "pv") for op in self.pipe), f"warp {w}: writes P_{t} with a PV product queued"
            self.p_tag[w] = t
            self.bar["p"].arrive()
        if self.W == 8:
            yield from self._pair_exchange(w, n, n & 1)     # the two half-row sums, in the slot the last tile left free
        yield ("wait", "o", n - 1)
        assert self.o_products == n, f"warp {w}: epilogue reads O with {self.o_products}/{n} products"

    def _pair_exchange(self, w, t, parity):
        """write own slot, named barrier with the partner warp, read the partner's slot"""
        partner = w ^ 4
        self.xbuf[(parity, w)] = t
        yield None
        if "pair" not in self.skip:
            b = self.bar[f"pair{w & 3}"]
            gen = b.completed
            b.arrive()
            while b.completed == gen:                      # bar.sync: block until the partner has arrived too
                yield None
        assert self.xbuf.get((parity, partner)) == t, (f"warp {w}: reads the exchange slot of tile "
                                                       f"{self.xbuf.get((parity, partner))} instead of {t}")
        yield None

    # ---- scheduler ------------------------------------------------------------------------------------------
    def run(self, max_steps=200000):
        threads = {"issuer": self.issuer()}
        threads.update({f"w{w}": self.softmax(w) for w in range(self.W)})
        blocked = {}
        while threads:
            self.steps += 1
            if self.steps > max_steps:
                raise Deadlock(f"no progress: blocked = {blocked}")
            self.unit_step()
            runnable = []
            for name in threads:
                if name in blocked:
                    bar, phase = blocked[name]
                    b = self.bar[bar]
                    if b.ready(phase):
                        assert b.completed - 1 == phase, (f"{name}: waited for phase {phase} of '{bar}' and woke on "
                                                          f"phase {b.completed - 1}")
                        runnable.append(name)
                    else:
                        assert b.completed - 1 < phase, (f"{name}: phase {phase} of '{bar}' can no longer be observed "
                                                         f"(barrier is at {b.completed - 1}): hardware hang")
                else:
                    runnable.append(name)
            if not runnable:
                if not self.tma and not self.pipe and self.pipe_busy is None:
                    raise Deadlock(f"all threads blocked with idle units: {blocked}")
                continue
            name = self.rng.choice(runnable)
            blocked.pop(name, None)
            try:
                req = next(threads[name])
            except StopIteration:
                del threads[name]
                continue
            if req is not None:
                blocked[name] = (req[1], req[2])
        assert not self.pipe and self.pipe_busy is None and not self.tma, "asynchronous work left behind at exit"
        return self.steps


def run(n_tiles, seed, skip=(), n_warps=4):
    return Sim(n_tiles, seed, skip, n_warps).run()


if __name__ == "__main__":
    import sys
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    for n_warps in (4, 8):
        for n_tiles in (1, 2, 3, 4, 6, 9):
            total = sum(run(n_tiles, s, n_warps=n_warps) for s in range(n_seeds))
            print(f"{n_warps} softmax warps, n_tiles={n_tiles}: {n_seeds} random schedules ok "
                  f"({total / n_seeds:.0f} steps on average)")
