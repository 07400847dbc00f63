"""Development aid: attribute the per-instruction samples / executed counts of an `ncu --set full --import-source on`
capture to source lines and inlined call chains, using nvdisasm's line info of the same cubin.

    cuobjdump -xelf all lib.so; nvdisasm --print-line-info-inline X.cubin > X.dis
    ncu -i rep.ncu-rep --page source --csv > X.csv
    python tools/sass_profile.py X.dis X.csv <source dir> [n_envs]
"""
import collections
import csv
import re
import sys

dis, src_csv, srcdir = sys.argv[1:4]
n_envs = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
ins = []
chain = []
fresh = True
for line in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        if fresh:
            chain = []
            fresh = False
        if not chain:
            chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        if m.group(3):
            chain.append((m.group(3).split('/')[-1], int(m.group(4))))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        ins.append((list(chain), m.group(2)))
        fresh = True
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] != "Address"][:len(ins)]
assert len(data) == len(ins), (len(data), len(ins))

# enclosing function of a source line
funcs = {}
for f in set(c[0] for ch, _ in ins for c in ch):
    try:
        lines = open(f"{srcdir}/{f}").read().split("\n")
    except OSError:
        continue
    cur = "?"
    tab = []
    for i, l in enumerate(lines, 1):
        m = re.match(r'\s*(?:MUAV_HD|__device__|__global__|template|static|inline).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$', l)
        if m and not l.strip().startswith("//") and "return" not in l.split("(")[0]:
            cur = m.group(1)
        tab.append(cur)
    funcs[f] = tab


def fn(c):
    f, l = c
    t = funcs.get(f)
    return (t[l - 1] if t and l - 1 < len(t) else f) if t else f


S = ix["# Samples"]
E = ix["Instructions Executed"]
TE = ix["Thread Instructions Executed"]
tot_s = sum(int(r[S]) for r in data)
tot_e = sum(int(r[E]) for r in data)
print(f"instructions {len(ins)}  executed/env {tot_e / n_envs:.0f}  samples {tot_s}")
for depth_name, keyf in (("outer frames (kernel line > next frame function)",
                          lambda ch: (ch[-1], fn(ch[-2]) if len(ch) > 1 else "-")),
                         ("innermost function", lambda ch: fn(ch[0])),
                         ("second frame from outside (function:line)",
                          lambda ch: (fn(ch[-2]), ch[-2][1], fn(ch[-3]) if len(ch) > 2 else "-") if len(ch) > 1 else ("kernel", ch[-1][1], "-"))):
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for (ch, txt), r in zip(ins, data):
        if not ch:
            continue
        k = keyf(ch)
        a = agg[k]
        a[0] += int(r[S])
        a[1] += int(r[E])
        a[2] += int(r[TE])
        a[3] += 1
    print(f"\n== by {depth_name}: samples%  exec/env  lanes  static")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"  {100 * a[0] / tot_s:5.1f}%  {a[1] / n_envs:7.1f}  {a[2] / max(a[1], 1):5.1f}  {a[3]:5d}  {k}")
