"""SASS evidence per kernel family (profiles/r02_sass_<family>.txt): `cuobjdump -sass` of the shipped library, per kernel the instruction
count and the counts of the mnemonics that prove the data path — UTMALDG (cp.async.bulk.tensor = tensor-tile TMA), UBLKCP (bulk TMA),
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), FFMA2 (fma.rn.f32x2), LDGSTS (cp.async), SYNCS (mbarrier) — plus the first lines of each.
    python tools/sass_excerpt.py [lib.so] [outdir]      (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "yolo_fastest_b200", "libyf_b200.so")
outdir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")
KEYS = ["UTMALDG", "UBLKCP", "UTCHMMA", "LDTM", "STTM", "FFMA2", "FFMA", "LDGSTS", "SYNCS", "LDS", "STS", "BAR"]
FAMILIES = {"wirb": ["wirb_kernel", "wstem_kernel"], "tcgen05": ["irbt_kernel", "irbtc_kernel", "irbtc2_kernel", "dwpw_tc_kernel", "upcat_tc_kernel", "dense_tc_kernel", "dense_ta_kernel"],
            "ffma": ["irb_kernel", "pw_kernel", "pwpw_kernel", "stem_kernel", "upcat_kernel"], "post": ["post_kernel", "compact_dets_kernel", "prep_bgr"]}
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\*[0-9a-fx]+\*/\s*$", "", line).strip())
for fam, pats in FAMILIES.items():
    out = ["# %s: SASS of %s (sm_100a), kernels matching %s" % (fam, os.path.basename(lib), ", ".join(pats)), ""]
    for name, ins in funcs.items():
        if not any(("yf::" + p) in name or (" " + p) in name for p in pats):
            continue
        ops = collections.Counter()
        for l in ins:
            t = l.split()
            op = t[1] if len(t) > 1 and not t[1].startswith("@") else (t[2] if len(t) > 2 else "?")
            ops[op.split(".")[0]] += 1
        out.append("%s" % name.replace("yf::", "")[:200])
        out.append("  %d instructions: " % len(ins) + " ".join("%s %d" % (k, ops[k]) for k in KEYS if ops[k]))
        for k in ("UTMALDG", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "FFMA2"):
            ex = [l for l in ins if re.search(r"\b%s\b" % k, l.split(";")[0])][:2]
            out += ["    " + e[:150] for e in ex]
        out.append("")
    open(os.path.join(outdir, "r02_sass_%s.txt" % fam), "w").write("\n".join(out))
    print(fam, sum(1 for n in funcs if any(("yf::" + p) in n for p in pats)), "kernels")
