"""CPU suite: discrete-event model of the experimental FP16-path GEMM's mbarrier protocol
(atom_b200/csrc/gemm_f16path_sm100.cuh).  Every warp role is a coroutine that performs the kernel's waits and arrivals
(transcribed from the source, including the parities); the scheduler interleaves them in random orders.  Checked:
the CTA always runs to completion (no deadlock), no barrier ever receives more arrivals than its phase expects, no
waiter is ever overtaken by two phase flips (parity aliasing), and every stage's data dependencies hold:
a slot is never overwritten before the MMAs that read it completed, and never read before it was fully written."""
import random

import pytest


class MBar:
    def __init__(self, count, name):
        self.count, self.pending, self.phase, self.tx, self.name = count, count, 0, 0, name

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, n=1):
        assert self.pending >= n, f"{self.name}: more arrivals than the phase expects"
        self.pending -= n
        self._maybe_flip()

    def expect_tx(self, b):
        self.tx += b

    def complete_tx(self, b):
        self.tx -= b
        assert self.tx >= 0, f"{self.name}: tx underflow"
        self._maybe_flip()

    def done(self, parity):           # try_wait.parity: has the phase with this parity completed?
        return (self.phase & 1) != parity


def run_cta(G, kPack, kRing, kConv, direct, seed):
    rng = random.Random(seed)
    SS = 4
    nunits, nstages = G + 2, 2 * G + 2
    pack_full = [MBar(1, f"pack_full{i}") for i in range(kPack)]
    pack_empty = [MBar(kConv, f"pack_empty{i}") for i in range(kPack)]
    exp_full = [MBar(kConv, f"exp_full{i}") for i in range(kRing)]
    slot_free = [MBar(1, f"slot_free{i}") for i in range(kRing)]
    scale_full = [MBar(32, f"scale_full{i}") for i in range(SS)]
    scale_empty = [MBar(kConv, f"scale_empty{i}") for i in range(SS)]
    acc_ready = MBar(1, "acc_ready")
    inflight = []                                         # async completions (TMA bytes, tcgen05.commit arrivals)
    state = {"packed_unit": [None] * kPack, "exp_writes": [dict() for _ in range(kRing)], "exp_stage": [None] * kRing,
             "mma_done_stage": -1, "scale_group": [None] * SS}

    def wait(bar, parity, expect_phase):
        # expect_phase: the phase index the waiter means; passing on a later phase with the same parity would be aliasing
        while not bar.done(parity):
            yield
        assert bar.phase == expect_phase + 1, f"{bar.name}: waiter for phase {expect_phase} woke in phase {bar.phase} (aliasing)"

    def producer():
        issued = min(kPack, nunits)
        for u in range(nunits):
            ps = u % kPack
            if u >= issued:
                yield from wait(pack_empty[ps], ((u // kPack) & 1) ^ 1, u // kPack - 1)
            assert state["packed_unit"][ps] is None or state["packed_unit"][ps] == u - kPack
            pack_full[ps].expect_tx(1)                     # mbarrier.arrive.expect_tx is one atomic operation:
            pack_full[ps].arrive()                         # the tx count is raised before the arrival is counted
            inflight.append(("tma_pack", ps, u))
            yield

    def mma():
        for t in range(nstages):
            es = t % kRing
            yield from wait(exp_full[es], (t // kRing) & 1, t // kRing)
            assert state["exp_stage"][es] == t and len(state["exp_writes"][es]) == kConv, "MMA read an incomplete slot"
            inflight.append(("commit_slot", es, t))
            yield
        inflight.append(("commit_acc", 0, nstages - 1))

    def scale_loader():
        for g in range(G + 1):
            ss = g % SS
            if g >= SS:
                yield from wait(scale_empty[ss], ((g // SS) - 1) & 1, g // SS - 1)
            state["scale_group"][ss] = g
            scale_full[ss].arrive(32)                     # cp.async noinc arrivals of the 32 lanes
            yield

    def converter(cw):
        for t in range(nstages):
            es = t % kRing
            is_i4 = t < 2 * G
            u = (t >> 1) if is_i4 else G + (t - 2 * G)
            ps = u % kPack
            half = (t & 1) if is_i4 else 0
            sg = u if is_i4 else G
            ss = sg % SS
            first_unit, last_unit = (not is_i4) or half == 0, (not is_i4) or half == 1
            first_group = (half == 0) if is_i4 else (u == G)
            last_group = (half == 1) if is_i4 else (u == G + 1)
            if t >= kRing:
                yield from wait(slot_free[es], ((t // kRing) - 1) & 1, t // kRing - 1)
                assert state["mma_done_stage"] >= t - kRing, "slot overwritten before its MMAs completed"
            if direct and cw == 0:
                exp_full[es].expect_tx(1)
                inflight.append(("tma_exp", es, t))
            if first_unit:
                yield from wait(pack_full[ps], (u // kPack) & 1, u // kPack)
            assert state["packed_unit"][ps] == u, "converter read the wrong packed unit"
            if first_group:
                yield from wait(scale_full[ss], (sg // SS) & 1, sg // SS)
            assert state["scale_group"][ss] == sg, "converter read the wrong group's scales"
            if state["exp_stage"][es] != t:
                state["exp_stage"][es], state["exp_writes"][es] = t, {}
            yield
            state["exp_writes"][es][cw] = True
            exp_full[es].arrive()
            if last_unit:
                pack_empty[ps].arrive()
            if last_group:
                scale_empty[ss].arrive()
            yield
        yield from wait(acc_ready, 0, 0)

    tasks = [producer(), mma(), scale_loader()] + [converter(c) for c in range(kConv)]
    alive = list(range(len(tasks)))
    idle_rounds = 0
    while alive:
        progressed = False
        rng.shuffle(alive)
        for i in list(alive):
            before = (tuple(b.phase for b in pack_full + pack_empty + exp_full + slot_free + scale_full + scale_empty), len(inflight))
            try:
                next(tasks[i])
            except StopIteration:
                alive.remove(i)
                progressed = True
                continue
            after = (tuple(b.phase for b in pack_full + pack_empty + exp_full + slot_free + scale_full + scale_empty), len(inflight))
            progressed |= before != after
        # asynchronous completions land at random times
        rng.shuffle(inflight)
        for _ in range(rng.randint(0, len(inflight))):
            kind, idx, tag = inflight.pop()
            progressed = True
            if kind == "tma_pack":
                state["packed_unit"][idx] = tag
                pack_full[idx].complete_tx(1)
            elif kind == "tma_exp":
                exp_full[idx].complete_tx(1)
            elif kind == "commit_slot":
                state["mma_done_stage"] = max(state["mma_done_stage"], tag)
                slot_free[idx].arrive()
            else:
                acc_ready.arrive()
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 200, f"deadlock: G={G} kPack={kPack} kRing={kRing} alive={alive}"
    assert not inflight or all(k == "commit_slot" for k, _, _ in inflight)


@pytest.mark.parametrize("G,kPack,kRing,kConv,direct", [(1, 3, 3, 16, False), (1, 4, 4, 16, False), (3, 3, 3, 16, False),
                                                        (7, 4, 4, 16, False), (31, 3, 3, 16, False), (5, 3, 4, 16, True),
                                                        (2, 4, 5, 16, True), (9, 2, 2, 8, False)])
def test_protocol_terminates_and_respects_dependencies(G, kPack, kRing, kConv, direct):
    for seed in range(12):
        run_cta(G, kPack, kRing, kConv, direct, seed)
