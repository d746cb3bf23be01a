"""Per-kernel SASS instruction histogram of libsignal_b200.so (run here, no GPU needed):

    python tools/sass_histogram.py > profiles/r2_sass_histogram.md

`cuobjdump -sass` of the in-tree library; for every kernel the total instruction count and the counts of the mnemonics
that prove which hardware path it uses (tcgen05: UTCHMMA / UTCBAR / LDTM; TMA: UTMALDG / UBLKCP; mbarrier: SYNCS;
NVLS multimem.ld_reduce: LDGMC (the switch adds the replicas); plain math: FFMA, HFMA2, MUFU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "signal_b200", "libsignal_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGMC", "FFMA", "HFMA2", "MUFU",
        "LDG", "STG", "LDS", "STS", "ATOM", "RED", "SHFL", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(n):
    for s in ("void ", "sig::", "(anonymous namespace)::", "<unnamed>::"):
        n = n.replace(s, "")
    depth, cut = 0, len(n)
    for i, ch in enumerate(n):      # drop the argument list: the first "(" outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return n[:cut][:110]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            h = hist[kern]
            h["_total"] += 1
            base = op.split(".")[0]
            if base in KEYS:
                h[base] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                h["UTCHMMA.2CTA"] += 1
    names = demangle(list(hist))
    cols = [k for k in KEYS if any(h[k] for h in hist.values())]
    tot = collections.Counter()
    for h in hist.values():
        tot.update(h)
    print("# SASS instruction histogram per kernel (round 2)\n")
    print("`cuobjdump -sass signal_b200/libsignal_b200.so` (sm_100a), %d kernels, %d instructions.  Whole library: " % (len(hist), tot["_total"])
          + ", ".join(f"{k} {tot[k]}" for k in cols) + ".\n")
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for k, h in sorted(hist.items(), key=lambda kv: -kv[1]["_total"]):
        print(f"| `{short(names[k])}` | {h['_total']} | " + " | ".join(str(h[c]) if h[c] else "" for c in cols) + " |")


if __name__ == "__main__":
    main()
