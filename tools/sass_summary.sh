#!/usr/bin/env bash
# SASS evidence for the hot kernels of the built library: which tensor-core / TMA / bulk-copy instructions they contain.
#   bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
so=chrono-mind_b200/libchronomind_b200.so
echo "# cuobjdump -sass $so (sm_100a), instruction mnemonics per kernel; built from $(git rev-parse --short HEAD)"
cuobjdump -sass "$so" > /tmp/cm_sass.txt
python3 - <<'PY'
import re, collections
cur, per = None, collections.OrderedDict()
for line in open("/tmp/cm_sass.txt"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        per[cur][m.group(1)] += 1
want = ("UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "UTCATOMSWS", "SYNCS", "FFMA", "HMMA", "LDG", "LDS", "ATOM", "RED", "BAR", "ELECT", "UTMAPF")
for fn, c in per.items():
    short = re.sub(r"^_ZN2cm\d+", "", fn)
    total = sum(c.values())
    hits = collections.OrderedDict()
    for k in want:
        n = sum(v for op, v in c.items() if op.startswith(k))
        if n: hits[k] = n
    detail = sorted((op for op in c if op.startswith(("UTC", "UTMA", "UBLKCP", "LDTM", "SYNCS", "UTMAPF"))))
    print("%-70s %6d instr  %s" % (short[:70], total, " ".join("%s=%d" % kv for kv in hits.items())))
    if detail:
        print("    " + " ".join(detail))
PY
