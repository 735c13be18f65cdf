"""Join an ncu SASS source page (csv) with nvdisasm -g line info -> warp-stall samples / executed instructions per CUDA
source line (and per named line range).  Usage:
  python tools/ncu_lines.py <report.ncu-rep> <libgtf_b200.so> <mangled-kernel-substring> [top_n] [name:file:lo-hi,...]"""
import csv, os, re, sys, subprocess, collections
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
subprocess.run("cd /tmp && rm -rf cubx && mkdir cubx && cd cubx && cuobjdump -xelf all %s >/dev/null 2>&1 && nvdisasm -g -c *.cubin > all.sass 2>/dev/null" % so, shell=True, check=True)
lines = open('/tmp/cubx/all.sass').read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('//---') and ('.text.' in l) and kern in l][0]
off2line = {}
cur = None
inl = None
for l in lines[start + 1:]:
    if l.startswith('//---'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m: off2line[int(m.group(1), 16)] = cur
flt = ["-k", os.environ["NCU_K"]] if os.environ.get("NCU_K") else []
if os.environ.get("NCU_SKIP"): flt += ["--launch-skip", os.environ["NCU_SKIP"], "--launch-count", "1"]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(out.split('\n')))
hdr = rows[1]
ia, isamp, iinst = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows[2:]:
    if len(r) <= iinst or r[ia] == 'Address': continue
    a = int(r[ia], 16)
    if base is None: base = a
    key = off2line.get(a - base, ('?', 0))
    agg[key][0] += int(r[isamp]); agg[key][1] += int(r[iinst])
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v: agg[key][2][h[6:]] += v
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print("total samples %d, warp instructions %d" % (tot, toti))
src = {}
def srcline(f, n):
    try:
        if f not in src: src[f] = open('' + os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gnn-track-finding_b200', 'csrc') + '/' + f).read().split('\n')
        return src[f][n - 1].strip()[:90]
    except Exception: return ''
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  [%s]  %s" % (100 * v[0] / tot, 100 * v[1] / toti, key[0], key[1],
          ' '.join('%s:%d' % kv for kv in v[2].most_common(3)), srcline(*key)))

if len(sys.argv) > 5:
    # phase summary: "name:file:lo-hi,..."
    import collections as C
    ph = C.OrderedDict()
    for spec in sys.argv[5].split(','):
        nm, fl, rng = spec.split(':'); lo, hi = map(int, rng.split('-')); ph[nm] = (fl, lo, hi)
    res = C.defaultdict(lambda: [0, 0, C.Counter()])
    for key, v in agg.items():
        nm = 'other'
        for k2, (fl, lo, hi) in ph.items():
            if key[0] == fl and lo <= key[1] <= hi: nm = k2; break
        res[nm][0] += v[0]; res[nm][1] += v[1]; res[nm][2].update(v[2])
    print("---- phases")
    for nm, v in sorted(res.items(), key=lambda kv: -kv[1][0]):
        print("%-10s %5.1f%% smp %5.1f%% inst  %s" % (nm, 100 * v[0] / tot, 100 * v[1] / toti, ' '.join('%s:%d' % kv for kv in v[2].most_common(5))))
