"""Opcode histogram of the built library (run here, no GPU): which kernels use tcgen05 / TMEM / TMA, and that none falls
back to the legacy mma.sync path.   python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "protnote_b200", "lib", "libprotnote_b200.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG.2D", "UTMALDG.3D", "UTMALDG.4D", "UTMASTG", "SYNCS", "ELECT",
         "HMMA", "HGMMA", "LDGSTS", "USETMAXREG", "UCGABAR_ARV", "BAR", "LDG", "STG", "LDS", "STS", "ATOMG", "RED"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        funcs[cur][op] += 1
        funcs[cur]["_total"] += 1
print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({os.path.getsize(LIB)} bytes), sm_100a; instruction counts per kernel")
print("tcgen05.mma = UTCHMMA (.2CTA: cta_group::2), tcgen05.commit = UTCBAR, tcgen05.ld = LDTM, TMA loads = UTMALDG, mbarrier = SYNCS,")
print("setmaxnreg = USETMAXREG; HMMA / HGMMA (legacy mma.sync / wgmma) must be absent everywhere.\n")
for name, c in funcs.items():
    def fam(prefix):
        return sum(v for k, v in c.items() if k == prefix or k.startswith(prefix + "."))
    short = re.sub(r"\(.*", "", name)
    cols = {"UTCHMMA": fam("UTCHMMA") - fam("UTCHMMA.2CTA"), "UTCHMMA.2CTA": fam("UTCHMMA.2CTA"), "UTCBAR": fam("UTCBAR"),
            "LDTM": fam("LDTM"), "UTMALDG": fam("UTMALDG"), "SYNCS": fam("SYNCS"), "ELECT": fam("ELECT"),
            "USETMAXREG": fam("USETMAXREG"), "HMMA": fam("HMMA") + fam("HGMMA"), "LDG": fam("LDG"), "STG": fam("STG"),
            "LDS": fam("LDS"), "STS": fam("STS"), "ATOM/RED": fam("ATOMG") + fam("RED") + fam("ATOMS")}
    keep = {k: v for k, v in cols.items() if v or k in ("HMMA",)}
    print(f"{short[:70]:70s} {c['_total']:6d} instr  " + "  ".join(f"{k} {v}" for k, v in keep.items()))
tot_h = sum(sum(v for k, v in c.items() if k.startswith("HMMA") or k.startswith("HGMMA")) for c in funcs.values())
print(f"\nlegacy tensor instructions (HMMA/HGMMA) in the whole library: {tot_h}")
sys.exit(1 if tot_h else 0)
