#!/usr/bin/env python
"""Discrete simulation of the mbarrier / named-barrier protocol of attention3.cuh (CPU only, no GPU needed).

Every role of one persistent CTA (TMA producer, the two MMA issuers, the two slots' softmax groups) is a Python
generator that yields the barrier it waits on; mbarriers follow the hardware rule (a parity wait succeeds when the
barrier's current phase parity differs from the waited parity), the named barriers of the MUFU turns and of the LONE
merge are two-party barriers with a sync side and an arrive side, TMA loads and MMA commits complete after a random
delay.  The simulation checks that every item list drains (no deadlock), that no ring stage / Q buffer / TMEM region is
overwritten before its readers are done, that every consumer reads the tile it expects, that the two slots are never
inside their exponentials at the same time, and that slot 0 merges the partial that belongs to its item.

    python tools/attn3_protocol_sim.py            # sweep of shapes / grids / seeds
"""
import random
import sys

ST = 4


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):
        return (self.phase & 1) != (parity & 1)


class NBar:
    """Named barrier shared by two parties (all warps of a slot collapsed into one agent): completes a phase when both
    the bar.sync side and the bar.arrive side have reached it."""

    def __init__(self):
        self.count, self.phase = 0, 0

    def arrive(self):
        self.count += 1
        assert self.count <= 2, "named barrier over-subscribed"
        if self.count == 2:
            self.count, self.phase = 0, self.phase + 1

    def test(self, target_phase):
        return self.phase >= target_phase


class Geo:
    def __init__(self, n, H, B):
        self.ntiles = (n + 127) // 128
        self.nblk = (n + 127) // 128
        self.npairs = self.ntiles >> 1
        self.p_total = B * H * self.npairs
        self.total = self.p_total + (B * H if self.ntiles & 1 else 0)
        self.H = H

    def item(self, idx):
        if idx < self.p_total:
            bh, tile0, lone = idx // self.npairs, 2 * (idx % self.npairs), False
        else:
            bh, tile0, lone = idx - self.p_total, self.ntiles - 1, True
        jbB = (self.nblk + 1) // 2 if lone else 0
        nsA = (self.nblk + 1) // 2 if lone else self.nblk
        return dict(bh=bh, q0=(tile0, tile0 if lone else tile0 + 1), jb=(0, jbB), ns=(nsA, self.nblk - jbB), lone=lone)

    def ring_pos(self, it, i, t):
        if it["lone"]:
            return 2 * i + t if i < it["ns"][1] else it["ns"][1] + i
        return i


class Sim:
    def __init__(self, n, H, B, G, cta, seed):
        self.rng = random.Random(seed)
        self.geo = Geo(n, H, B)
        self.G, self.cta = G, cta
        self.q_full = [MBar(1), MBar(1)]
        self.q_empty = [MBar(1), MBar(1)]
        self.kv_full = [MBar(1) for _ in range(ST)]
        self.kv_empty = [MBar(2) for _ in range(ST)]
        self.s_full = [MBar(1), MBar(1)]
        self.s_free = [MBar(1), MBar(1)]   # the 8 warp arrivals collapsed into one agent
        self.p_full = [MBar(1), MBar(1)]
        self.pv_done = [MBar(1), MBar(1)]
        self.turn = [NBar(), NBar()]      # EXP_TURN + slot: that slot may run its exponentials
        self.xo_full, self.xo_free = NBar(), NBar()
        self.exp_owner = None             # slot inside its exponentials (the MUFU token: never both)
        self.xo = None                    # item whose slot-1 partial sits in the exchange buffer
        self.events = []  # (time, fn) asynchronous completions
        self.now = 0
        # data model: what each buffer currently holds / who still reads it
        self.ring = [None] * ST      # (item idx, key block)
        self.ring_readers = [0] * ST
        self.qbuf = [None, None]
        self.S = [None, None]        # (item, block) computed
        self.P = [None, None]
        self.O = [None, None]        # (item, blocks accumulated)
        self.done_items = {0: [], 1: []}

    def later(self, fn, lo=1, hi=30):
        self.events.append((self.now + self.rng.randint(lo, hi), fn))

    def items(self):
        idx, ord_ = self.cta, 0
        while idx < self.geo.total:
            yield idx, ord_, self.geo.item(idx)
            idx += self.G
            ord_ += 1

    # ---- roles ----
    def producer(self):
        qn = [0, 0]
        for idx, ord_, it in self.items():
            for t in range(2):
                if it["ns"][t] == 0:
                    continue
                yield (self.q_empty[t], (qn[t] & 1) ^ 1)

                def land(t=t, idx=idx, it=it):
                    self.qbuf[t] = (idx, it["q0"][t])
                    self.q_full[t].arrive()
                self.later(land)
                qn[t] += 1
            c = ord_ * self.geo.nblk
            for i in range(it["ns"][0]):
                for t in range(2 if it["lone"] else 1):
                    if i >= it["ns"][t]:
                        continue
                    j = it["jb"][t] + i
                    s = c % ST
                    yield (self.kv_empty[s], ((c // ST) & 1) ^ 1)
                    assert self.ring_readers[s] == 0, "ring stage overwritten while MMAs still read it"

                    def land(s=s, idx=idx, j=j):
                        self.ring[s] = (idx, j)
                        self.kv_full[s].arrive()
                    self.ring[s] = "in flight"
                    self.later(land)
                    c += 1

    def mma(self, t):
        geo = self.geo
        st = dict(k_qk=0, k_pv=0, qn=0, seen=0)
        cur = dict(it=None, idx=self.cta, ord=0, i=0)

        def seek():
            while cur["idx"] < geo.total:
                cur["it"] = geo.item(cur["idx"])
                if cur["it"]["ns"][t] > 0:
                    break
                cur["idx"] += self.G
                cur["ord"] += 1
            cur["i"] = 0

        def issue_next_qk():
            if cur["idx"] >= geo.total:
                return
            it = cur["it"]
            if cur["i"] == 0:
                yield (self.q_full[t], st["qn"] & 1)
                st["qn"] += 1
            c = cur["ord"] * geo.nblk + geo.ring_pos(it, cur["i"], t)
            while st["seen"] <= c:
                yield (self.kv_full[st["seen"] % ST], (st["seen"] // ST) & 1)
                st["seen"] += 1
            if st["k_qk"] > 0:
                yield (self.s_free[t], (st["k_qk"] - 1) & 1)
            s = c % ST
            want = (cur["idx"], it["jb"][t] + cur["i"])
            assert self.ring[s] == want, f"QK slot {t}: stage {s} holds {self.ring[s]}, wanted {want}"
            assert self.qbuf[t] == (cur["idx"], it["q0"][t]), f"QK slot {t}: Q buffer holds {self.qbuf[t]}"
            last = cur["i"] == it["ns"][t] - 1

            def done(want=want, last=last):
                self.S[t] = want
                self.s_full[t].arrive()
                if last:
                    self.q_empty[t].arrive()
            self.later(done)
            st["k_qk"] += 1
            cur["i"] += 1
            if cur["i"] == it["ns"][t]:
                cur["idx"] += self.G
                cur["ord"] += 1
                seek()

        seek()
        yield from issue_next_qk()
        for idx, ord_, it in self.items():
            for i in range(it["ns"][t]):
                yield from issue_next_qk()
                yield (self.p_full[t], st["k_pv"] & 1)
                c = ord_ * geo.nblk + geo.ring_pos(it, i, t)
                s = c % ST
                want = (idx, it["jb"][t] + i)
                assert self.ring[s] == want, f"PV slot {t}: stage {s} holds {self.ring[s]}, wanted {want}"
                assert self.P[t] == want, f"PV slot {t}: P holds {self.P[t]}, wanted {want}"
                self.ring_readers[s] += 1

                def done(s=s, idx=idx, i=i, lone=it["lone"]):
                    self.ring_readers[s] -= 1
                    self.O[t] = (idx, i + 1) if i else (idx, 1)
                    self.kv_empty[s].arrive()
                    if lone:
                        self.kv_empty[s].arrive()
                    self.pv_done[t].arrive()
                self.later(done)
                st["k_pv"] += 1

    def nb_sync(self, bar):
        """bar.sync: arrive, then wait for the phase this arrival belongs to."""
        target = bar.phase + 1
        bar.arrive()
        yield (bar, target)

    def softmax(self, t):
        k = 0
        first_merge = True
        if t == 1:
            self.turn[0].arrive()  # slot 0 goes first
        for idx, ord_, it in self.items():
            ns = it["ns"][t]
            for i in range(ns):
                yield (self.s_full[t], k & 1)
                want = (idx, it["jb"][t] + i)
                assert self.S[t] == want, f"softmax slot {t}: S holds {self.S[t]}, wanted {want}"
                self.s_free[t].arrive()
                if i > 0:
                    yield (self.pv_done[t], (k - 1) & 1)
                    assert self.O[t] == (idx, i), f"softmax slot {t}: O is {self.O[t]} before block {i}"
                yield from self.nb_sync(self.turn[t])          # the MUFU turn
                assert self.exp_owner is None, "both slots inside their exponentials"
                self.exp_owner = t
                yield (NBar(), 0)                               # let the other roles run while this slot "computes"
                self.exp_owner = None
                self.turn[t ^ 1].arrive()
                self.P[t] = want
                self.p_full[t].arrive()
                k += 1
            if t == 1:
                for _ in range(it["ns"][0] - it["ns"][1]):     # turns slot 1 does not use
                    yield from self.nb_sync(self.turn[1])
                    self.turn[0].arrive()
            if ns == 0:
                continue
            yield (self.pv_done[t], (k - 1) & 1)
            assert self.O[t] == (idx, ns), f"epilogue slot {t}: O is {self.O[t]}, wanted {(idx, ns)}"
            merge = it["lone"] and it["ns"][1] > 0
            if merge and t == 1:
                if not first_merge:
                    yield from self.nb_sync(self.xo_free)
                first_merge = False
                assert self.xo is None, "slot 1 overwrote a partial that slot 0 has not read"
                self.xo = idx
                self.xo_full.arrive()
            elif merge:
                yield from self.nb_sync(self.xo_full)
                assert self.xo == idx, f"slot 0 merged the partial of item {self.xo} into item {idx}"
                self.xo = None
                self.xo_free.arrive()
            self.done_items[t].append(idx)

    def run(self):
        roles = {"producer": self.producer(), "mma0": self.mma(0), "mma1": self.mma(1), "sm0": self.softmax(0),
                 "sm1": self.softmax(1)}
        waiting = {}
        for name, g in roles.items():
            try:
                waiting[name] = next(g)
            except StopIteration:
                pass
        steps = 0
        while waiting or self.events:
            steps += 1
            assert steps < 5_000_000, "runaway"
            progressed = False
            names = list(waiting)
            self.rng.shuffle(names)
            for name in names:
                bar, par = waiting[name]
                if bar.test(par):
                    progressed = True
                    try:
                        waiting[name] = next(roles[name])
                    except StopIteration:
                        del waiting[name]
            if not progressed:
                if not self.events:
                    raise RuntimeError(f"deadlock: waiting roles {list(waiting)}")
                self.events.sort(key=lambda e: e[0])
                tm, fn = self.events.pop(0)
                self.now = max(self.now, tm)
                fn()
        want = [idx for idx, _, it in self.items()]
        assert self.done_items[0] == want, "slot 0 did not finish every item"
        assert self.done_items[1] == [idx for idx, _, it in self.items() if it["ns"][1] > 0]


def main():
    cases = 0
    for n in (40, 128, 129, 256, 321, 361, 513, 553, 681, 1193):
        for BH in (1, 3, 24):
            for G in (1, 2, 5, 148):
                geo = Geo(n, BH, 1)
                for cta in sorted({0, min(G, geo.total) - 1}):
                    if cta >= geo.total:
                        continue
                    for seed in range(4):
                        Sim(n, BH, 1, min(G, geo.total), cta, seed).run()
                        cases += 1
    print(f"attention3 protocol simulation: {cases} cases drained without deadlock or buffer hazard")


if __name__ == "__main__":
    sys.exit(main())
