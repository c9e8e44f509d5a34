"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses.

    python tools/sass_summary.py > profiles/sass_summary.txt

UTCHMMA / UTCQMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = 1-D bulk copy, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, USETMAXREG = setmaxnreg, FFMA2 = packed fp32 FMA, SYNCS = mbarrier operations.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "spoofsv_b200" / "libspoofsv_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "USETMAXREG", "FFMA2", "FFMA", "SYNCS", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], stdout=subprocess.PIPE, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), stdout=subprocess.PIPE, text=True).stdout.split("\n")
    counts, order, cur, i = {}, [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[i]; i += 1
            cur = re.sub(r"\(anonymous namespace\)::|ssv::|void ", "", cur)
            cur = re.sub(r"\(.*\)$", "", cur)
            counts[cur] = collections.Counter(); order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for k in KEYS:
                if op == k or (k in ("SYNCS", "BAR", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "USETMAXREG", "UTCHMMA", "UTCQMMA") and op.startswith(k)):
                    counts[cur][k] += 1
                    break
    print(f"# {LIB.name}: SASS mnemonic counts per kernel (tools/sass_summary.py; sm_100a)")
    print(f"{'kernel':72s} " + " ".join(f"{k:>9s}" for k in KEYS))
    for name in order:
        c = counts[name]
        if not any(c[k] for k in KEYS):
            continue
        print(f"{name[:72]:72s} " + " ".join(f"{c[k]:9d}" for k in KEYS))


if __name__ == "__main__":
    main()
