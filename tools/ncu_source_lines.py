"""Per-source-line warp-stall sample counts of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
    python tools/ncu_source_lines.py prof.ncu-rep regex:kernel_name [top_n]
"""
import collections
import csv
import subprocess
import sys


def main(path, kernel, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", kernel, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr = None, None
    agg, txt = collections.Counter(), {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():
            key = (cur_file, int(r[0]))
            try:
                agg[key] += int(r[4] or 0)
            except ValueError:
                pass
            txt[key] = r[1]
    tot = sum(agg.values())
    print("total samples", tot)
    for k, v in agg.most_common(top):
        print(f"{k[0]}:{k[1]:<5d} {v:6d} {100.0 * v / max(tot, 1):5.1f}%  {txt[k][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
