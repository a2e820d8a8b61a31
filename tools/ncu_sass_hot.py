"""Warp-stall samples of one kernel in an .ncu-rep at SASS level (needs `ncu --set full`; no -lineinfo required):

    python tools/ncu_sass_hot.py prof.ncu-rep regex:kernel_name [top_n]

Prints the most-sampled SASS instructions with the three instructions before each, and the samples that sit on mbarrier
waits grouped by the barrier's shared-memory offset -- in the warp-specialised kernels every role waits on its own barriers, so
that table says which role starves which (how the stem's loader / epilogue / pool bottlenecks of DESIGN.md section 10 were found;
source-line views fold all inlined `mbar_wait` calls into one line)."""
import collections
import csv
import re
import subprocess
import sys


def load(path, kernel):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kernel,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    hdr, data = None, []
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) == len(hdr):
            data.append(r)
    c = hdr.index("# Samples")
    return [(d[1], int(d[c] or 0)) for d in data]


def main(path, kernel, top=12):
    ins = load(path, kernel)
    tot = sum(n for _, n in ins)
    print(f"{tot} samples over {len(ins)} instructions")
    waits, last = collections.Counter(), None
    for i, (txt, n) in enumerate(ins):
        m = re.search(r"SYNCS\.PHASECHK.*\+(0x[0-9a-f]+)\]", txt)
        if m:
            last = (m.group(1), i)
        if last and i - last[1] <= 1:                 # the try-wait and the branch that spins on it
            waits[last[0]] += n
    print("samples on mbarrier waits, by barrier offset:")
    for off, n in sorted(waits.items()):
        print(f"  {off}: {n:6d}  {100.0 * n / max(tot, 1):5.1f} %")
    hot = sorted(range(len(ins)), key=lambda i: -ins[i][1])[:top]
    for i in sorted(hot):
        print(f"----- instruction {i}: {ins[i][1]} samples ({100.0 * ins[i][1] / max(tot, 1):.1f} %)")
        for j in range(max(0, i - 3), i + 1):
            print(f"    {j:5d} {ins[j][1]:6d}  {ins[j][0][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 12)
