"""Attribute the executed warp instructions of one kernel in an .ncu-rep to source lines / regions:
   ncu_by_line.py report.ncu-rep object.o kernel-mangled-substring source.cu [top]"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, obj, ksub, src = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h = next(r for r in rows if "Instructions Executed" in r)
ie, sc, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
cnt = []
for r in rows[rows.index(h) + 1:]:
    try: cnt.append((int(r[ie]), r[sc].strip(), int(r[ss])))
    except Exception: pass
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside, cur, lines = False, None, []
for l in dis:
    if l.startswith("\t.section\t.text."):
        inside = ksub in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"^\s*/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
print("static instructions: report %d, object %d" % (len(cnt), len(lines)))
assert len(cnt) == len(lines)
tot = sum(c for c, _, _ in cnt); ts = sum(s for _, _, s in cnt)
by, bs = collections.Counter(), collections.Counter()
for (c, _, s), ln in zip(cnt, lines): by[ln] += c; bs[ln] += s
text = open(src).read().split("\n")
print("total warp instructions %d" % tot)
for (f, ln), c in by.most_common(top):
    t = text[ln - 1].strip()[:95] if f == os.path.basename(src) else f
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100 * c / tot, 100 * bs[(f, ln)] / ts, f, ln, t))
