"""Developer tool: per-kernel SASS opcode histogram of perception_b200/libcuboid_cuda.so (cuobjdump -sass), written as markdown.
Shows which instruction classes each kernel is made of and whether the TMA bulk-copy (UBLKCP), mbarrier (SYNCS), async-copy (LDGSTS),
cluster barrier (UCGABAR / BAR with cluster scope) and tensor-core (UTCxMMA / HMMA) opcodes are present."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "perception_b200", "libcuboid_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("cuboid::", "").replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(2).split(".")[0]] += 1
special = ["UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "UTCHMMA", "UTCQMMA", "HMMA", "UCGABAR_ARV", "UCGABAR_WAIT", "MATCH", "REDUX", "ATOMS", "ATOMG", "RED", "BAR", "MUFU", "FFMA", "FMUL", "FADD", "DFMA", "DADD", "DMUL"]
print("| kernel | SASS instructions | top opcodes | of note |")
print("|---|---|---|---|")
for k, c in hist.items():
    tot = sum(c.values())
    if tot < 50:
        continue
    top = ", ".join("%s %d" % (o, n) for o, n in c.most_common(8))
    note = ", ".join("%s %d" % (o, c[o]) for o in special if c.get(o))
    print("| `%s` | %d | %s | %s |" % (k, tot, top, note))
