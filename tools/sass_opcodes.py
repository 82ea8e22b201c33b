"""Opcode histogram of the built library per kernel (cuobjdump -sass): the Blackwell-native evidence of
B200_PROFILING.md (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA and bulk
copies, HMMA = the legacy mma.sync path).      python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "h-tsp_b200", "libhtsp_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "MUFU", "FFMA", "FFMA2",
       "FADD2", "LDGSTS", "LDG", "STG", "LDS", "STS", "BAR", "R2P"]
tot = collections.Counter()
print(f"# {os.path.basename(lib)}: static SASS opcode counts per kernel (cuobjdump -sass)")
print("# kernel".ljust(86) + " instr  " + " ".join(k.rjust(7) for k in KEY))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0].strip()
    try:
        name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        pass
    ops = collections.Counter()
    n = 0
    for ln in f.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            ops[m.group(1)] += 1
            n += 1
    tot.update(ops)
    short = re.sub(r"\(.*", "", name).replace("htsp::", "").replace("void ", "")[:84]
    print(short.ljust(86) + f"{n:6d}  " + " ".join(str(ops.get(k, 0)).rjust(7) for k in KEY))
print("TOTAL".ljust(86) + f"{sum(tot.values()):6d}  " + " ".join(str(tot.get(k, 0)).rjust(7) for k in KEY))
