"""How much does a communication-shaped kernel running NEXT TO a GEMM cost the GEMM?  (1 GPU, "peers" = this GPU.)

Stream A runs a loop of ViT-10B GEMMs (qkv forward NT 32768x15360x5120 and fc1 wgrad TN); stream B runs, for the whole
duration, back-to-back copies of one candidate: the light all-gather kernel (with / without the L2 evict-first hint, 16
/ 64 / 148 CTAs), a copy-engine D2D memcpy, the reduce-scatter kernel.  Reported: GEMM loop time alone, with the
neighbour, and the slowdown per millisecond of neighbour activity -- the number that, multiplied by the collectives'
duty cycle, predicts the multi-GPU tax (profiles/r2_comm_v2.md).

    python tools/exp_overlap.py [--json gpurun_out/overlap.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co  # noqa: E402

C = co._C


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    ap.add_argument("--iters", type=int, default=16)
    args = ap.parse_args()
    dev = torch.device("cuda")
    T, D = 32768, 5120
    x = (torch.randn(T, D, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(3 * D, D, device=dev) * 0.02).to(torch.bfloat16)
    dy = (torch.randn(T, 4 * D, device=dev) * 0.5).to(torch.bfloat16)
    gw = torch.empty(4 * D, D, device=dev, dtype=torch.bfloat16)

    def gemm_loop():
        for _ in range(args.iters):
            co.linear_fwd(x, w)
            co.linear_wgrad(dy, x, out=gw)

    # neighbour payload: one ViT-10B block worth of bf16 parameters (600 MiB) "gathered" from two self-peers
    n = 300 * 2 ** 20 // 2  # elements per rank shard
    shard = torch.randn(n, device=dev).to(torch.bfloat16)
    full = torch.empty(2 * n, device=dev, dtype=torch.bfloat16)
    grad = torch.randn(2 * n, device=dev).to(torch.bfloat16)
    out32 = torch.empty(n, device=dev, dtype=torch.float32)
    chunk = C.ag_chunk_bytes()
    rows, prefix = [], 0
    for r in range(2):
        rows.append([r, 0, r * n * 2, n * 2, prefix])
        prefix += -(-n * 2 // chunk)
    ag_table, ag_chunks = torch.tensor(rows, dtype=torch.int64, device=dev), prefix
    rs_chunk = C.rs_chunk_vecs() * 8
    rs_table = torch.tensor([[0, 0, n, 0]], dtype=torch.int64, device=dev)
    rs_chunks = -(-n // rs_chunk)
    peers_shard = [shard.data_ptr(), shard.data_ptr()]
    peers_grad = [grad.data_ptr(), grad.data_ptr()]

    def ag(ctas):
        return lambda: C.p2p_all_gather(peers_shard, 0, full, ag_table, ag_chunks, ctas)

    def rs(ctas):
        return lambda: C.reduce_scatter(peers_grad, 0, 0, 2, out32, rs_table, rs_chunks, True, 0.5, None, ctas, [], [],
                                        None, None, None, None, None, None, [])

    def memcpy():
        full[:n].copy_(shard, non_blocking=True)
        full[n:].copy_(shard, non_blocking=True)

    cands = {"none": None, "ag_64": ag(64), "ag_16": ag(16), "ag_148": ag(148), "memcpy_d2d": memcpy,
             "rs_64": rs(64), "rs_148": rs(148)}
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    res = {"l2_hint": os.environ.get("B200_COMM_L2_HINT", "1"), "iters": args.iters}
    for name, fn in cands.items():
        # neighbour alone: how long does one call take, to size the queue
        one_ms = 0.0
        if fn is not None:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            one_ms = e0.elapsed_time(e1) / 10
        with torch.cuda.stream(sa):
            gemm_loop()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nb = 0
        if fn is not None:
            nb = int(args.iters * 8.5 / max(one_ms, 0.05)) + 8  # enough calls to cover the whole GEMM loop
            with torch.cuda.stream(sb):
                b0.record()
                for _ in range(nb):
                    fn()
                b1.record()
        with torch.cuda.stream(sa):
            a0.record()
            gemm_loop()
            a1.record()
        torch.cuda.synchronize()
        r = {"gemm_loop_ms": round(a0.elapsed_time(a1), 2), "neighbour_alone_ms_per_call": round(one_ms, 3),
             "neighbour_calls": nb}
        if fn is not None:
            r["neighbour_ms_per_call_overlapped"] = round(b0.elapsed_time(b1) / nb, 3)
        res[name] = r
        print(name, r, flush=True)
    base = res["none"]["gemm_loop_ms"]
    for name in cands:
        if name != "none":
            res[name]["gemm_slowdown_pct"] = round(100.0 * (res[name]["gemm_loop_ms"] / base - 1.0), 1)
    print(json.dumps(res), flush=True)
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
