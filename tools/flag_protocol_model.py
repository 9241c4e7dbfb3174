"""Protocol model of the cross-GPU flag protocol inside the collectives (csrc/comm.cu: sync_begin / sync_end), CPU only.

compute-sanitizer cannot watch several GPUs at once, so the multi-GPU hand-offs are checked the way the attention
pipelines are (tools/pipeline_model.py): the protocol is transcribed and replayed under randomised schedules.

Modelled system: W ranks.  Each rank has a *compute stream* that produces gradient-buffer contents (two rotating
buffers, like the engine's `_grad_bufs`) and a *communication stream* that runs one reduce-scatter kernel per unit, in
the same order on every rank.  A kernel is C CTAs:

  sync_begin : CTA 0 publishes ready[peer][rank] = seq to every peer; every CTA waits until ready[rank][p] >= seq for all p
  body       : the CTA reads its chunks of EVERY peer's buffer (the model records which version it saw)
  sync_end   : the CTA bumps a counter; the LAST CTA publishes done[peer][rank] = seq, waits until done[rank][p] >= seq for
               all p, resets the counter and stores seq_dev = seq.  Only then does the kernel -- and the event the compute
               stream waits on before it overwrites the buffer -- complete.

Checked for every schedule: liveness (no deadlock), every read sees exactly the version written for that call (never
the previous use of the buffer, never the next), and no buffer is overwritten while any rank still has reads of it
outstanding.  Fault injection (`bug=`): "no_ready_wait" and "no_done_wait" must be caught (the checker is sensitive);
"early_seq_store" (sequence counter stored before the done-wait) is shown to be harmless: the next kernel of the stream
only starts after this one has completed.
"""
from __future__ import annotations

import random
from typing import Dict, List


class FlagProtocolError(AssertionError):
    pass


class Rank:
    def __init__(self, r: int, world: int):
        self.r = r
        self.ready = [0] * world   # this rank's flag row: ready[p] written by peer p
        self.done = [0] * world
        self.seq_dev = 0
        self.cta_ctr = 0
        self.buf_version = [0, 0]          # contents of the two gradient buffers: call index + 1 they belong to
        self.buf_readers = [0, 0]          # reads in flight (any rank) per buffer
        self.kernel_done = 0               # number of completed reduce-scatter kernels (== events recorded)


def simulate(world: int, calls: int, ctas: int, seed: int, bug: str = "", max_steps: int = 400_000) -> None:
    rng = random.Random(seed)
    ranks = [Rank(r, world) for r in range(world)]

    # ---------------- actors: one generator per (rank, compute stream) and per (rank, call, CTA) ----------------
    def compute(rk: Rank):
        for k in range(calls):
            b = k & 1
            # the engine waits for the event recorded after the kernel that last read this buffer (call k - 2)
            while k >= 2 and rk.kernel_done < k - 1:
                yield "blocked"
            if rk.buf_readers[b]:
                raise FlagProtocolError(f"rank {rk.r}: gradient buffer {b} overwritten for call {k} while "
                                        f"{rk.buf_readers[b]} read(s) of call {k - 2} are still in flight")
            rk.buf_version[b] = k + 1
            rk.produced = k + 1
            yield "step"

    def cta(rk: Rank, k: int, c: int):
        # stream order: the kernel starts after the previous kernel of this rank has completed and after this rank's
        # compute stream has produced call k's gradients (event wait on the communication stream)
        while rk.kernel_done < k or getattr(rk, "produced", 0) < k + 1:
            yield "blocked"
        seq = rk.seq_dev + 1
        if seq != k + 1:
            raise FlagProtocolError(f"rank {rk.r} call {k}: sequence number {seq}, expected {k + 1}")
        if c == 0:
            for p in ranks:
                p.ready[rk.r] = seq
                yield "step"
        if bug != "no_ready_wait":
            for p in range(world):
                while rk.ready[p] < seq:
                    yield "blocked"
        # body: read every peer's buffer of this call
        b = k & 1
        for p in ranks:
            if p.buf_version[b] != k + 1:
                raise FlagProtocolError(f"rank {rk.r} call {k} CTA {c}: read buffer {b} of rank {p.r} holding version "
                                        f"{p.buf_version[b]}, expected {k + 1}")
            p.buf_readers[b] += 1
            yield "step"
            p.buf_readers[b] -= 1
        # sync_end
        rk.cta_ctr += 1
        last = rk.cta_ctr == ctas
        yield "step"
        if last:
            if bug == "early_seq_store":
                rk.seq_dev = seq
            for p in ranks:
                p.done[rk.r] = seq
                yield "step"
            if bug != "no_done_wait":
                for p in range(world):
                    while rk.done[p] < seq:
                        yield "blocked"
            rk.cta_ctr = 0
            rk.seq_dev = seq
            rk.kernel_done = k + 1
        yield "step"

    live: Dict[str, object] = {}
    for rk in ranks:
        live[f"compute{rk.r}"] = compute(rk)
        for k in range(calls):
            for c in range(ctas):
                live[f"r{rk.r}k{k}c{c}"] = cta(rk, k, c)
    names: List[str] = list(live)
    for _ in range(max_steps):
        if not live:
            return
        progressed = False
        order = names[:]
        rng.shuffle(order)
        for name in order:
            g = live.get(name)
            if g is None:
                continue
            try:
                res = next(g)
            except StopIteration:
                del live[name]
                progressed = True
                continue
            if res == "step":
                progressed = True
                break  # one action at a time: maximally interleaved schedules
        names = [n for n in names if n in live]
        if not progressed:
            raise FlagProtocolError(f"deadlock with {len(live)} actors alive, e.g. {sorted(live)[:4]}")
    raise FlagProtocolError("step limit reached")


if __name__ == "__main__":
    for seed in range(50):
        simulate(4, 6, 3, seed)
        simulate(8, 4, 2, seed)
    print("flag protocol held for 100 schedules")
