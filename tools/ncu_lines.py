#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report captured with --import-source on.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep focal_fwd [top]

Runs `ncu -i REP --page source --csv --print-source cuda,sass -k regex:KERNEL` and prints, for the `top` source lines by
warp-level instructions executed: share of instructions, share of stall samples, and the dominant stall reasons.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows, header, fpath, seen_kernel = [], None, None, None
    for rec in csv.reader(io.StringIO(txt)):
        if not rec:
            continue
        if rec[0] == "File Path":
            fpath = rec[1].split("/")[-1]
        elif rec[0] == "Function Name":
            if seen_kernel is None:
                seen_kernel = rec[1]
            cur_kernel = rec[1]
        elif rec[0] == "Line No":
            header = rec
        elif header and rec[2] == "-" and cur_kernel == seen_kernel:   # a source-line aggregate row
            d = dict(zip(header[4:], rec[4:]))
            try:
                inst = int(d["Instructions Executed"])
                samp = int(d["# Samples"])
            except (KeyError, ValueError):
                continue
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "(" not in k and v.isdigit() and int(v)}
            rows.append((fpath, int(rec[0]), rec[1].strip(), inst, samp, stalls))
    ti = sum(r[3] for r in rows) or 1
    ts = sum(r[4] for r in rows) or 1
    print(f"kernel: {seen_kernel}\n total warp-instructions {ti:,}   samples {ts:,}")
    rows.sort(key=lambda r: -(r[4] if '--by-samples' in sys.argv else r[3]))
    print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s}  top stalls | source")
    for f, ln, src, inst, samp, st in rows[:top]:
        s3 = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{f + ':' + str(ln):28s} {100 * inst / ti:6.2f} {100 * samp / ts:6.2f}  {s3:38s} | {src[:90]}")


if __name__ == "__main__":
    main()
