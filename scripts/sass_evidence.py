#!/usr/bin/env python
"""cuobjdump -sass of libpsld_b200.so -> profiles/<tag>_sass_tc_kernels.txt: per tensor-core kernel, the
counts of the tcgen05 / TMEM / TMA mnemonics (UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit,
LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UTCATOMSWS / UTCMOV... = TMEM alloc) and their first
occurrences, as committed evidence that the contractions run on the 5th-generation tensor cores.

usage: python scripts/sass_evidence.py <tag>
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "psld_b200", "libpsld_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
PAT = re.compile(r"\b(UTCHMMA[\w.]*|UTCBAR[\w.]*|LDTM[\w.]*|STTM[\w.]*|UTMALDG[\w.]*|UTMAPF[\w.]*|UTCATOMSWS[\w.]*|"
                 r"UTCMOV[\w.]*|UTCCP[\w.]*|SYNCS[\w.]*|ELECT[\w.]*|UCGABAR[\w.]*|FFMA2|HMMA[\w.]*)\b")
out = [f"# SASS evidence ({tag}): tensor-core / TMEM / TMA instructions per kernel of libpsld_b200.so",
       "# produced by scripts/sass_evidence.py from `cuobjdump -sass` (sm_100a)", ""]
cur, body = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        body[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]+\*/", line):
        body[cur].append(line.strip())
for fn, lines in body.items():
    name = demangle(fn)
    if not re.search(r"conv_tc_kernel|conv_gn_tc_kernel|conv_gn_x3_kernel|attn_tc_kernel", name):
        continue
    cnt = collections.Counter()
    first = {}
    for l in lines:
        for m in PAT.finditer(l):
            k = m.group(1)
            cnt[k] += 1
            first.setdefault(k, l)
    out.append(f"## {name.split('(')[0]}   ({len(lines)} SASS instructions)")
    for k, v in sorted(cnt.items()):
        out.append(f"  {v:5d}  {k}")
    out.append("  first occurrences:")
    for k in sorted(first):
        if k.startswith(("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR")):
            out.append("    " + first[k][:150])
    out.append("")
path = os.path.join(ROOT, "profiles", f"{tag}_sass_tc_kernels.txt")
open(path, "w").write("\n".join(out) + "\n")
print("wrote", path, len(out), "lines")
