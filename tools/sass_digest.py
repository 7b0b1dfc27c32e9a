"""SASS digest of the shipped library: per kernel, the counts of the Blackwell-native instructions
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, UTMALDG/UTMASTG =
cp.async.bulk.tensor, LDGSTS = cp.async), registers, and local-memory (spill) instructions.

    python tools/sass_digest.py [posegen_b200/lib/libposegen_b200.so] > profiles/r2_sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "posegen_b200", "lib", "libposegen_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "HMMA", "LDL", "STL", "RED", "ATOM"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        for k in KEYS:
            if op == k or (k == "UTCHMMA" and op.startswith("UTC") and op.endswith("MMA")):
                counts[cur][k] += 1


def demangle(n):
    try:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        return d.split("(")[0].replace("void ", "")[-70:]
    except Exception:  # noqa: BLE001
        return n[-70:]


print(f"# SASS digest of {os.path.basename(lib)} (cuobjdump -sass / -res-usage; sm_100a)")
print(f"{'kernel':72s} {'instr':>7s} {'regs':>5s} {'local':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
for fn, c in counts.items():
    reg, sh, loc = usage.get(fn, (0, 0, 0))
    print(f"{demangle(fn):72s} {c['_total']:7d} {reg:5d} {loc:6d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
