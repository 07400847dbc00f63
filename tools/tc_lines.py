"""Development aid: samples of an ncu source-page CSV per outermost source line of att_pair_tc_kernel.
  cuobjdump -xelf all build/muav_scorer_tc.o; nvdisasm --print-line-info-inline X.cubin > tc.dis
  ncu -i rep --page source --csv > src.csv;  python tools/tc_lines.py tc.dis src.csv"""
import collections, csv, re, sys
txt = open(sys.argv[1]).read()
secs = re.split(r'(?m)^//-+ \.text\.', txt)
sec = [s for s in secs if s.startswith('_ZN7muav_tc18att_pair')][0]
ins, chain, fresh = [], [], True
for line in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        if fresh:
            chain, fresh = [], False
        if not chain:
            chain.append(int(m.group(2)))
        if m.group(3):
            chain.append(int(m.group(4)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        ins.append((list(chain), m.group(2)))
        fresh = True
rows = list(csv.reader(open(sys.argv[2])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(ins) == len(data), (len(ins), len(data))
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
by, byi = collections.Counter(), collections.Counter()
for (ch, op), r in zip(ins, data):
    outer = ch[-1] if ch else 0
    by[outer] += int(r[ix['# Samples']] or 0)
    byi[outer] += int(r[ix['Instructions Executed']] or 0)
src = open('multi_uav_ta_gym_env_b200/csrc/muav_scorer_tc.cu').read().splitlines()
print("total samples", tot)
for ln, c in by.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    print(ln, "%.1f%%" % (100 * c / tot), byi[ln], src[ln - 1].strip()[:100] if ln else "")
