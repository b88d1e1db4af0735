"""A small discrete-event model of the staging pipeline of k_dw_adam_fused (csrc/gemm_sm100.cu) with the mbarrier
PARITY-wait semantics of the hardware, used by tests/test_sync_protocol.py.

Why it exists: round 1 shipped a version of that kernel that stalled once in several thousand steps.  Two epilogue groups
alternated on a staging stage that had ONE "rows have landed" barrier; a warp could get two phases ahead of a barrier it
had not looked at for a while, and `mbarrier.try_wait.parity` cannot tell "phase k completed" from "phase k+2 completed".
Nothing on a GPU short of a long soak showed it; this model shows it in milliseconds, and pins the invariant the fix
restores: *every waiter observes consecutive phases of every barrier it waits on*.

Model (names as in the kernel):
  * row groups i = 0 .. n-1; stage s = i % NS; epilogue group g = i % 2 handles row group i with W warps;
  * I/O thread: loads row groups 0..NS-1, then for i = 0..: wait done[s] (all W warps of the group), store(i), and --
    once the store of i-1 has finished reading shared memory -- load(i - 1 + NS) into the stage of i-1;
  * warp: wait "rows landed" for i, process, arrive on done[s];
  * a load / store / processing step takes a random time (heavy tail: the interesting interleavings are the ones where an
    early load is slower than two later ones).
A barrier is (phase count, arrivals pending); `parity_wait(P)` passes iff (phase count & 1) != P -- exactly
mbarrier.try_wait.parity.  The model knows which phase a wait is MEANT to see, so it can flag
  early  -- the wait passed although that phase has not completed (the round-1 bug), and
  missed -- that phase has completed but the parity test no longer shows it (the waiter blocks until a later phase)."""
import heapq
import random


class Barrier:
    def __init__(self, count):
        self.count = count          # arrivals per phase
        self.pending = count
        self.phases = 0             # completed phases

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending = self.count
            self.phases += 1

    def parity_passes(self, parity):
        return (self.phases & 1) != parity


def simulate(n_groups_of_rows, NS, design, W=2, seed=0, tail=0.05):
    """design: "per_stage" (one rows-landed barrier per stage, parity (i // NS) & 1 -- the round-1 kernel) or
    "per_stage_and_group" (one per (stage, epilogue group), parity (i // (2 NS)) & 1 -- the fix; needs NS odd, or
    parity (i // NS) & 1 when NS is even and a stage always belongs to one group).
    -> dict(early=[...], missed=[...], deadlock=bool, finished=int)"""
    rng = random.Random(seed)
    n = n_groups_of_rows

    def dur(mean):
        # mostly ~mean, sometimes 20x (an HBM access that queues behind a burst, a warp that loses the scheduler)
        return rng.expovariate(1.0 / mean) * (20.0 if rng.random() < tail else 1.0)

    if design == "per_stage":
        ld_full = [Barrier(1) for _ in range(NS)]
        ld_bar = lambda i: ld_full[i % NS]
        ld_parity = lambda i: (i // NS) & 1
        ld_use = lambda i: i // NS                      # which phase of that barrier the wait is meant to see
    elif design == "per_stage_and_group":
        ld_full = [[Barrier(1), Barrier(1)] for _ in range(NS)]
        ld_bar = lambda i: ld_full[i % NS][i % 2]
        if NS % 2:
            ld_parity = lambda i: (i // (2 * NS)) & 1
            ld_use = lambda i: i // (2 * NS)
        else:
            ld_parity = lambda i: (i // NS) & 1
            ld_use = lambda i: i // NS
    else:
        raise ValueError(design)
    done = [Barrier(W) for _ in range(NS)]

    events = []                                          # (time, seq, kind, payload)
    seq = [0]

    def at(t, kind, payload):
        seq[0] += 1
        heapq.heappush(events, (t, seq[0], kind, payload))

    early, missed = [], []
    warps = [{"g": g, "k": 0, "busy": False} for g in (0, 1) for _ in range(W)]
    rows_of = {0: list(range(0, n, 2)), 1: list(range(1, n, 2))}
    io = {"i": 0, "state": "wait_done", "read_done": set()}
    finished = [0]
    now = [0.0]

    for i in range(min(NS, n)):
        at(dur(1.0), "load_done", i)

    def poll():
        """Let every agent that can make progress at the current time do so (zero-time steps)."""
        progressed = True
        while progressed:
            progressed = False
            for w in warps:
                if w["busy"] or w["k"] >= len(rows_of[w["g"]]):
                    continue
                i = rows_of[w["g"]][w["k"]]
                bar, use = ld_bar(i), ld_use(i)
                if bar.parity_passes(ld_parity(i)):
                    if bar.phases < use + 1:
                        early.append((i, w["g"], bar.phases, use))
                    w["busy"] = True
                    at(now[0] + dur(0.5), "processed", (w, i))
                    progressed = True
                elif bar.phases >= use + 1 and (i, w["g"]) not in [(m[0], m[1]) for m in missed]:
                    missed.append((i, w["g"], bar.phases, use))
            i = io["i"]
            if i < n and io["state"] == "wait_done" and done[i % NS].parity_passes((i // NS) & 1):
                at(now[0] + dur(0.3), "store_read_done", i)            # bulk store of i issued; it reads the stage for a while
                io["state"] = "wait_store_read"                        # cp.async.bulk.wait_group.read 1 (lag by one)
                progressed = True
            if i < n and io["state"] == "wait_store_read" and (i == 0 or (i - 1) in io["read_done"]):
                if i >= 1 and i - 1 + NS < n:
                    at(now[0] + dur(1.0), "load_done", i - 1 + NS)     # reload the stage of i-1
                io["i"] += 1
                io["state"] = "wait_done"
                progressed = True

    poll()
    steps = 0
    while events and steps < 400000:
        steps += 1
        t, _, kind, payload = heapq.heappop(events)
        now[0] = t
        if kind == "load_done":
            ld_bar(payload).arrive()
        elif kind == "processed":
            w, i = payload
            done[i % NS].arrive()
            w["k"] += 1
            w["busy"] = False
            finished[0] += 1
        elif kind == "store_read_done":
            io["read_done"].add(payload)
        poll()
    deadlock = finished[0] < n * W
    return {"early": early, "missed": missed, "deadlock": deadlock, "finished": finished[0]}
