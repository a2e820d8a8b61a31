"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library (run where cuobjdump is installed):

    python tools/sass_summary.py [tdnet_b200/lib/libtdnet_b200.so] > profiles/r02_sass_summary.txt

tcgen05.mma -> UTCHMMA (.2CTA for cta_group::2), tcgen05.ld / st -> LDTM / STTM, tcgen05.commit -> UTCBAR,
TMA loads -> UTMALDG, TMA stores -> UTMASTG, legacy mma.sync would show as HMMA (B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCCP", "HMMA", "SYNCS", "MUFU.EX2"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["instructions"] += 1
        for p in PATTERNS:
            if op == p or op.startswith(p + "."):
                if p == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                    continue
                cur[p] += 1
                break
    print(f"# {os.path.relpath(path, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
    print("kernel".ljust(58) + "".join(p.rjust(13) for p in ["instructions"] + PATTERNS))
    tot = collections.Counter()
    for name, c in per.items():
        if not any(c[p] for p in PATTERNS if p not in ("SYNCS", "MUFU.EX2")):
            continue
        print(name[:57].ljust(58) + "".join(str(c[p]).rjust(13) for p in ["instructions"] + PATTERNS))
        tot.update(c)
    print("TOTAL (kernels listed)".ljust(58) + "".join(str(tot[p]).rjust(13) for p in ["instructions"] + PATTERNS))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tdnet_b200", "lib", "libtdnet_b200.so"))
