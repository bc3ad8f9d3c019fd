"""Host models of the ordered commit's step-wise consumers: (1) the tag-array step classification, (2) further down, the turn
protocol (sets of warps, named barriers, end markers).

(1) Tag-array step classification of the ordered commit (fastglobetrotter_b200/csrc/kernels.cu,
k_commit_ordered<STEPWISE>, lambda `classify`): which lanes of a 32-record step share a tile cell, found without match.any.

The kernel's bit-exactness tests (`-m gpu`) see whatever write order the hardware picks when several lanes store to one tag
byte.  This model replays the same protocol on the CPU with EVERY choice left open -- which of the competing writers of a slot
"wins" (last store), and what older steps left behind in the two arrays -- and checks the property the fold relies on:

  kind 0  => all valid cells of the step are distinct;
  kind 1  => every lane's partner is exactly the other lane with its cell (or itself when it is alone), and no cell is held
             by three or more lanes;
  kind 2  (fall back to match.any) only when some slot is shared by three or more lanes.

The fold then adds, per cell, the lower lane's value before the higher lane's: the reference's order (C:155 / C:161)."""
import numpy as np
import pytest

SLOTS = 512            # kTagSlots


def classify(cells, tA, tB, rng):
    """one step; cells[32] (-1 = no record); tA / tB = the warp's two tag arrays as older steps left them"""
    lanes = np.arange(32)
    valid = cells >= 0
    slot = cells & (SLOTS - 1)
    # round 1: every valid lane stores its number at its slot; the stores of one instruction land in an arbitrary order
    for l in rng.permutation(32):
        if valid[l]:
            tA[slot[l]] = l
    t = np.where(valid, tA[slot], lanes)
    lost = valid & (t != lanes)
    if not lost.any():
        return 0, lanes.copy()
    # round 2: the losers store their number into the second array
    for l in rng.permutation(32):
        if lost[l]:
            tB[slot[l]] = l
    u = np.where(valid, tB[slot], lanes)
    bad = lost & (u != lanes)
    cand = np.where(lost, t, u) & 31
    c_cell = cells[cand]                      # the two shuffles
    c_lost = lost[cand]
    paired = np.where(lost, c_cell == cells, valid & c_lost & (c_cell == cells))
    if bad.any():
        return 2, None
    return 1, np.where(paired, cand, lanes)


def check(cells, kind, partner):
    valid = cells >= 0
    groups = {}
    for l in np.nonzero(valid)[0]:
        groups.setdefault(int(cells[l]), []).append(int(l))
    slots = {}
    for l in np.nonzero(valid)[0]:
        slots.setdefault(int(cells[l]) & (SLOTS - 1), []).append(int(l))
    if kind == 0:
        assert all(len(g) == 1 for g in groups.values())
    elif kind == 1:
        assert all(len(g) <= 2 for g in groups.values())
        for g in groups.values():
            if len(g) == 1:
                assert partner[g[0]] == g[0]
            else:
                assert partner[g[0]] == g[1] and partner[g[1]] == g[0]
        for l in np.nonzero(~valid)[0]:
            assert partner[l] == l
    else:
        assert any(len(s) >= 3 for s in slots.values()), "fell back although no slot holds three lanes"


@pytest.mark.parametrize("n_cells", [8, 40, 300, 4650, 40000])
def test_step_classification_is_exact_for_every_write_order(n_cells):
    rng = np.random.default_rng(n_cells)
    kinds = [0, 0, 0]
    tA = rng.integers(0, 256, SLOTS).astype(np.int64)      # uninitialised shared memory: any byte
    tB = rng.integers(0, 256, SLOTS).astype(np.int64)
    for trial in range(3000):
        n_valid = int(rng.integers(1, 33))
        cells = np.full(32, -1, dtype=np.int64)
        cells[:n_valid] = rng.integers(0, n_cells, n_valid)   # a partial last step leaves the high lanes empty
        if trial % 7 == 0:                                     # slot aliases: different cells, same slot
            cells[:n_valid] = (cells[:n_valid] % 4) * SLOTS + rng.integers(0, 6, n_valid)
        kind, partner = classify(cells, tA, tB, rng)          # the arrays carry over from step to step, as in the kernel
        check(cells, kind, partner)
        kinds[kind] += 1
    if n_cells >= 4650:
        assert kinds[0] > 0 and kinds[1] > 0                   # both fast classes occur at the bench workload's tile size


def test_fold_order_of_a_couple_is_lane_order():
    """kind 1: the lower lane of a couple adds its own value first, then its partner's (kernels.cu, lambda `fold1`)."""
    rng = np.random.default_rng(1)
    tA = np.zeros(SLOTS, dtype=np.int64); tB = np.zeros(SLOTS, dtype=np.int64)
    for _ in range(500):
        cells = rng.integers(0, 600, 32).astype(np.int64)
        vals = rng.random(32) * 10.0 ** rng.integers(-8, 8, 32)
        kind, partner = classify(cells, tA, tB, rng)
        if kind != 1:
            continue
        tile = {int(c): 0.1 for c in cells}
        ref = dict(tile)
        for l in range(32):                                    # the reference: one += per record, in stream (= lane) order
            ref[int(cells[l])] = ref[int(cells[l])] + vals[l]
        for l in range(32):
            p = int(partner[l])
            if p >= l:
                acc = tile[int(cells[l])] + vals[l]
                if p != l:
                    acc = acc + vals[p]
                tile[int(cells[l])] = acc
        assert tile == ref


# ---------------------------------------------------------------------------------------------------------------------------
# The turn protocol of the step-wise consumers (kernels.cu, k_commit_ordered<STEPWISE>): two sets of two warps, set q takes the
# stages k = q (mod 2) of the ring, warp h of a set the steps 2h and 2h+1; the fold turn travels w0 -> w1 -> w2 -> w3 -> w0
# through named barriers (bar.sync by the waiter, bar.arrive by its predecessor); the producer ends every task with one end
# marker per set; a __syncthreads separates tasks.  Replayed here as a random interleaving of five cooperating "warps":
# no schedule may deadlock, no named barrier may receive a second arrive before its waiter passed, and the folds must come
# out in stream order -- also across tasks without records, tasks of one stage and partly filled stages.
RS, STAGE = 4, 128


def _run_protocol(task_stage_counts, rng):
    """task_stage_counts: per task, the record count of each of its stages (the last may be < 128)"""
    n_tasks = len(task_stage_counts)
    stages = {}                      # seq -> record count (-1 = end marker), filled by the producer
    released = {}                    # seq -> number of consumer warps that have read (and released) it
    bar = [0, 0, 0, 0]               # pending arrives of named barrier 1 + w
    sync_count = [0] * (n_tasks + 1) # __syncthreads at the end of task t
    folds = []                       # (seq, step) in the order the folds happen
    first_turn = [True, False, False, False]

    def barrier(t):
        sync_count[t] += 1
        while sync_count[t] < 5:
            yield

    def producer():
        k = 0
        for t, counts in enumerate(task_stage_counts):
            for c in list(counts) + [-1, -1]:                    # the task's stages, then one end marker per set
                while k >= RS and released.get(k - RS, 0) < 2:   # the ring slot is still being read
                    yield
                stages[k] = c
                k += 1
            yield from barrier(t)

    def consumer(w):
        st, half = w >> 1, w & 1
        k = st
        nxt_w = (w + 1) & 3

        def fetch(kk):
            while kk not in stages:
                yield
            released[kk] = released.get(kk, 0) + 1
            return stages[kk]

        for t in range(n_tasks):
            cur_k = k
            cur = yield from fetch(cur_k)
            k += 2
            while cur >= 0:
                nxt_k = k
                nxt = yield from fetch(nxt_k)
                k += 2
                if not first_turn[w]:
                    while bar[w] == 0:                           # bar.sync 1 + w, 64
                        yield
                    bar[w] -= 1
                first_turn[w] = False
                for h in (0, 1):
                    if 32 * (2 * half + h) < cur:
                        folds.append((cur_k, 2 * half + h))
                assert bar[nxt_w] == 0, "a second bar.arrive before the waiter passed"
                bar[nxt_w] += 1                                  # bar.arrive 1 + next
                cur_k, cur = nxt_k, nxt
            yield from barrier(t)

    procs = [producer()] + [consumer(w) for w in range(4)]
    alive = list(range(5))
    idle_rounds = 0
    while alive:
        i = alive[int(rng.integers(0, len(alive)))]
        before = (len(stages), sum(released.values()), len(folds), tuple(bar), tuple(sync_count))
        try:
            next(procs[i])
        except StopIteration:
            alive.remove(i)
        after = (len(stages), sum(released.values()), len(folds), tuple(bar), tuple(sync_count))
        idle_rounds = 0 if after != before or i not in alive else idle_rounds + 1
        assert idle_rounds < 20000, "deadlock"
    return folds, stages


@pytest.mark.parametrize("seed", range(12))
def test_turn_protocol_folds_in_stream_order_and_never_deadlocks(seed):
    rng = np.random.default_rng(seed)
    tasks = []
    for _ in range(int(rng.integers(1, 9))):
        n_stages = int(rng.choice([0, 0, 1, 1, 2, 3, 5, 8, 13]))
        counts = [STAGE] * n_stages
        if n_stages and rng.random() < 0.7:
            counts[-1] = int(rng.integers(1, STAGE + 1))         # a partly filled last stage
        tasks.append(counts)
    folds, stages = _run_protocol(tasks, rng)
    expect = [(k, s) for k in sorted(stages) if stages[k] >= 0 for s in range(4) if 32 * s < stages[k]]
    assert folds == expect
