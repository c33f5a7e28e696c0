"""SASS opcode histogram of the in-tree library per kernel (developer tool; output committed under profiles/).

    python tools/sass_histogram.py > profiles/sass_histogram_r2.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "softgnss_python_b200", "libsoftgnss_b200.so")
MARK = ("UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "HMMA", "FFMA2", "DFMA", "IDP",
        "BAR", "WARPSYNC", "SHFL", "FFMA", "FADD", "FMUL", "LDS", "STS", "LDG", "STG", "LDL", "STL")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    kernels, cur = [], None
    it = iter(names)
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = [next(it), collections.Counter()]
            kernels.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[1][m.group(1)] += 1
    print("# SASS opcode histogram of `libsoftgnss_b200.so` (sm_100a), round 2\n")
    print("`cuobjdump -sass` of the in-tree library, static instruction counts per kernel (one row per kernel family: the "
          "largest instantiation).  Blackwell-specific rows: `UBLKCP` = `cp.async.bulk` (TMA 1-D bulk copy), `SYNCS` = "
          "mbarrier, `LDGSTS` = `cp.async`; there is no `UTC*MMA` / `LDTM` (tcgen05) and no `HMMA`: the tensor-core "
          "options were measured and lost (profiles/tensor_core_experiment_r2.md).\n")
    fam = {}
    for name, cnt in kernels:
        base = re.sub(r"<.*", "", name.replace("void ", "")).strip()
        tot = sum(cnt.values())
        if base not in fam or tot > fam[base][2]:
            fam[base] = (name, cnt, tot)
    total = collections.Counter()
    for _, cnt in kernels:
        total.update(cnt)
    print("| kernel | instructions | " + " | ".join(MARK[:14]) + " | top opcodes |")
    print("|---|---|" + "---|" * 15)
    for base, (name, cnt, tot) in sorted(fam.items(), key=lambda x: -x[1][2]):
        top = ", ".join("%s %d" % kv for kv in cnt.most_common(6))
        print("| `%s` | %d | " % (name[:90].replace("|", "/"), tot) + " | ".join(str(sum(v for k, v in cnt.items() if k.startswith(m))) for m in MARK[:14]) + " | %s |" % top)
    print("\nWhole library (%d kernels): " % len(kernels) + ", ".join("%s %d" % (m, sum(v for k, v in total.items() if k.startswith(m))) for m in MARK))


if __name__ == "__main__":
    main()
