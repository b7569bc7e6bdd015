"""Per-source-line totals (warp instructions executed, stall samples) of one kernel in an ncu report.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n] [kernel-regex]

Uses `ncu -i ... --page source --print-source cuda,sass --csv`; needs a capture taken with --import-source on of a
library built with -lineinfo.  No GPU needed.
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if len(sys.argv) > 3:
        cmd += ["-k", "regex:" + sys.argv[3]]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    rows = csv.reader(out.splitlines())
    path, hdr, cur = None, None, None
    tot = {}
    src = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            i_long = hdr.index("stall_long_sb")
            continue
        if hdr is None:
            continue
        if r[0] not in ("", "-"):
            cur = (path, int(r[0]))
            src[cur] = r[1].strip()[:110]
        try:
            inst, samp, lsb = int(r[i_inst]), int(r[i_samp]), int(r[i_long])
        except (ValueError, IndexError):
            continue
        if r[2] in ("", "-"):
            continue  # the source line's own summary row: the SASS rows below it carry the numbers
        t = tot.setdefault(cur, [0, 0, 0])
        t[0] += inst
        t[1] += samp
        t[2] += lsb
    all_inst = sum(t[0] for t in tot.values()) or 1
    all_samp = sum(t[1] for t in tot.values()) or 1
    print(f"total warp instructions {all_inst}, samples {all_samp}")
    print("by instructions:")
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*t[0]/all_inst:5.1f}% inst {100*t[1]/all_samp:5.1f}% samp  {k[0]}:{k[1]}  {src.get(k,'')}")
    print("by samples:")
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100*t[1]/all_samp:5.1f}% samp ({100*t[2]/max(t[1],1):3.0f}% long_sb) {100*t[0]/all_inst:5.1f}% inst  {k[0]}:{k[1]}  {src.get(k,'')}")


if __name__ == "__main__":
    main()
