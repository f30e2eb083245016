"""tools/sass_census.py [lib.so]: per kernel, the SASS mnemonics that prove what the source claims — TMA bulk copies (UBLKCP),
mbarrier traffic (SYNCS.*), warp reductions (REDUX), votes (VOTE / VOTEU), 3-input logic (LOP3), shared loads (LDS*), and that no
tensor-core instruction exists on this integer path."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "bobrapet_b200/lib/libbobrafrontier.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print("# %s  (embedded cubins: %s)" % (lib, ", ".join(arch)))
keys = ["UBLKCP", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "REDUX", "VOTE", "LOP3", "LDS", "STS", "SHFL", "BMSK", "POPC", "FLO", "ATOMS", "RED", "BAR.SYNC", "WARPSYNC", "NANOSLEEP",
        "HMMA", "UTC", "LDTM", "IMMA"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
print("%-92s %6s " % ("kernel", "instr") + " ".join("%7s" % k[:7] for k in keys))
for k, c in counts.items():
    if "frontier" in k or "compact" in k or "sched" in k or "exp_" in k or "apply" in k or "validate" in k or "closure" in k or "move_records" in k:
        print("%-92s %6d " % (k[:92], c["_total"]) + " ".join("%7d" % c[x] for x in keys))
