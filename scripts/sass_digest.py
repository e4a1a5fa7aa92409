"""Developer tool: per-kernel SASS mnemonic counts of the built library -> profiles/<tag>_sass_digest.md, so that the
Blackwell-native claims (tcgen05 MMA = UTCHMMA, TMEM loads / stores = LDTM / STTM, TMA = UTMALDG / UTMASTG, bulk copies =
UBLKCP) can be checked from the tree.  `python scripts/sass_digest.py r02`"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "i2v_adapter_unofficial_b200", "libi2v_attn_b200.so")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU.EX2", "FFMA2", "SYNCS", "BAR"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
counts, order, cur = {}, [], None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        c = counts[cur]
        c["total"] += 1
        if op.startswith("UTCHMMA"):
            c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        for k in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "FFMA2", "SYNCS", "BAR"):
            if op.startswith(k):
                c[k] += 1
        if op.startswith("MUFU.EX2"):
            c["MUFU.EX2"] += 1


def short(n):
    n = n.replace("(int)", "").replace("(bool)", "").replace("(unsigned int)", "").replace("(anonymous namespace)::", "")
    n = re.sub(r"\((?:i2v::|const |int|float|long|unsigned|void|__nv).*$", "", n)
    return n.replace("void ", "").replace("i2v::", "")


lines = [f"# SASS digest of libi2v_attn_b200.so ({TAG})", "",
         "`cuobjdump -sass i2v_adapter_unofficial_b200/libi2v_attn_b200.so`, mnemonic counts per kernel (static instructions).",
         "UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,",
         "UBLKCP = cp.async.bulk, HMMA = mma.sync, SYNCS = mbarrier operations.", "",
         "| kernel | instr | " + " | ".join(COLS) + " |", "|---|---:|" + "---:|" * len(COLS)]
tot = collections.Counter()
for n in order:
    c = counts[n]
    tot.update(c)
    lines.append(f"| `{short(n)[:110]}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |")
lines.append(f"| **all {len(order)} kernels** | {tot['total']} | " + " | ".join(str(tot[k]) for k in COLS) + " |")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{TAG}_sass_digest.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-3:]))
