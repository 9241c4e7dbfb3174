"""Summarise an .ncu-rep into profiles/<name>.md: key roofline metrics + top stall sites (run on the CPU box).

    python tools/summarize_ncu.py gpurun_out/gemm_prof.ncu-rep profiles/gemm_qkv_fwd "note"
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "gpc__cycles_elapsed.max",
    "sm__cycles_active.avg", "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_uniform.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            d[h] = (v, u)
        res.append(d)
    return res


def source_top(rep, n=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    data = []
    for r in rows:
        if "Address" in r and "Source" in r:
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            s = int(r[hdr.index("Warp Stall Sampling (All Samples)")])
        except ValueError:
            continue
        st = {h: int(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v.isdigit() and int(v)}
        data.append((s, r[hdr.index("Source")], st))
    tot = sum(s for s, _, _ in data) or 1
    data.sort(key=lambda x: -x[0])
    return [(s, 100.0 * s / tot, src, sorted(st.items(), key=lambda kv: -kv[1])[:2]) for s, src, st in data[:n]], tot


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    kernels = raw(rep)
    lines = [f"# ncu summary: {rep}", "", note, ""]
    js = []
    for k in kernels:
        name = k.get("Kernel Name", ("?", ""))[0]
        lines.append(f"## {name[:160]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        rec = {"kernel": name}
        for key in KEYS:
            for h, (v, u) in k.items():
                if h == key or h.endswith("." + key) or h.split(".", 2)[-1] == key:
                    lines.append(f"| {key} | {v} | {u} |")
                    rec[key] = v
                    break
        js.append(rec)
        lines.append("")
    try:
        top, tot = source_top(rep)
        lines += ["## top stall sites (warp-state samples, all warps)", "", f"total samples: {tot}", "",
                  "| samples | % | SASS | top stall reasons |", "|---|---|---|---|"]
        for s, pct, src, st in top:
            lines.append(f"| {s} | {pct:.1f} | `{src[:90]}` | {st} |")
    except Exception as e:  # pragma: no cover
        lines.append(f"(source page unavailable: {e})")
    open(out + ".md", "w").write("\n".join(lines) + "\n")
    json.dump(js, open(out + ".json", "w"), indent=1)
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
