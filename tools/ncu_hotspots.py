"""Source-level hot spots of one kernel in an .ncu-rep (CPU box; needs `ncu` on PATH and a report captured with
--import-source on from a -lineinfo build).

    python tools/ncu_hotspots.py REPORT.ncu-rep KERNEL_ID [TOP_N]

Prints a markdown table: the source lines that collected the most warp-stall samples, their share of all samples of the
kernel and the dominant stall reasons (ncu's --page source view, CUDA-C + SASS correlation, aggregated per source line).
"""
import csv
import subprocess
import sys


def hotspots(rep: str, kernel_id: int, top: int = 12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                          f":::{kernel_id}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name, path, hdr, lines = "", "", None, []
    for r in rows:
        if len(r) == 2 and r[0] == "Function Name":
            name = r[1]
        elif len(r) == 2 and r[0] == "File Path":
            path = r[1].rsplit("/", 1)[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].strip().isdigit() and len(r) >= len(hdr) - 3:
            d = dict(zip(hdr, r))
            try:
                samples = int(d.get("# Samples", "0") or 0)
            except ValueError:
                continue
            if samples <= 0:
                continue
            reasons = []
            for k, v in d.items():
                if k.startswith("stall_") and not k.endswith("(Not Issued)"):
                    try:
                        n = int(v)
                    except ValueError:
                        continue
                    if n > 0:
                        reasons.append((n, k[len("stall_"):]))
            reasons.sort(reverse=True)
            lines.append((samples, path, int(r[0]), r[1].strip(), reasons[:3]))
    total = sum(x[0] for x in lines)
    lines.sort(reverse=True)
    return name, total, lines[:top]


def main():
    rep, kid = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    name, total, lines = hotspots(rep, kid, top)
    print(f"`{name[:140]}` -- {total} warp-stall samples\n")
    print("| share | file:line | source | dominant stall reasons (samples) |")
    print("|---|---|---|---|")
    for samples, path, line, src, reasons in lines:
        why = ", ".join(f"{k} {n}" for n, k in reasons)
        src = src.replace("|", "\\|")
        print(f"| {100.0 * samples / max(1, total):.1f} % | {path}:{line} | `{src[:110]}` | {why} |")


if __name__ == "__main__":
    main()
