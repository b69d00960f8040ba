"""Static evidence from the built library (no GPU): per kernel, the SASS instructions that prove which hardware
paths it uses -- UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP / UTMALDG (bulk / tensor TMA copies),
SYNCS (mbarrier), CREDUX / REDUX (redux.sync), FFMA2 (packed fp32), plus registers / shared memory / spills from
`cuobjdump -res-usage`.   python tools/sass_evidence.py > profiles/<tag>_sass_evidence.csv
Mnemonics as listed in /opt/skills/guides/B200_PROFILING.md (SASS names, the PTX names never appear in SASS)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3d_adapt_auto_driving_b200", "libpn2_b200.so")
GROUPS = [("tcgen05_mma", r"\bUTC[A-Z]*MMA"), ("tmem_ld", r"\bLDTM"), ("tmem_st", r"\bSTTM"), ("tma_bulk", r"\bUBLKCP"),
          ("tma_tensor", r"\bUTMA(LDG|STG)"), ("mbarrier", r"\bSYNCS"), ("redux", r"\bC?REDUX"), ("ffma2", r"\bFFMA2"),
          ("ffma", r"\bFFMA\b"), ("ldgsts", r"\bLDGSTS"), ("st_async", r"\bST\.ASYNC|\bSTAS\b"), ("local_mem", r"\b(LDL|STL)\b")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        counts[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4,}\*/", line) else 0
        for key, pat in GROUPS:
            if re.search(pat, line):
                counts[cur][key] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k: v for k, v in re.findall(r"(REG|SHARED|LOCAL|STACK):(\d+)", line)}
            cur = None
    names = demangle(list(counts))
    cols = ["instructions"] + [g[0] for g in GROUPS]
    print("# static SASS evidence of %s (cuobjdump -sass / -res-usage); sm_100a" % os.path.relpath(LIB, ROOT))
    print("kernel,regs,static_smem,stack_bytes," + ",".join(cols))
    for k, c in counts.items():
        u = usage.get(k, {})
        short = re.sub(r"\(anonymous namespace\)::", "", names.get(k, k))
        short = re.sub(r"\(.*", "", short).replace(",", ";")
        print("%s,%s,%s,%s,%s" % (short, u.get("REG", ""), u.get("SHARED", ""), u.get("STACK", ""),
                                  ",".join(str(c.get(col, 0)) for col in cols)))


if __name__ == "__main__":
    sys.exit(main())
