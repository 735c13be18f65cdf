"""Refresh profiles/ from gpurun_out/: python tools/make_profiles.py <bench.json> <launches.csv> <full.ncu-rep> <lib.so> [round prefix]"""
import collections, csv, json, os, subprocess, sys
bench_json, launches, rep, so = sys.argv[1:5]
RND = sys.argv[5] if len(sys.argv) > 5 else "r02"
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
P = os.path.join(R, "profiles")
j = json.loads(open(bench_json).read().strip().splitlines()[-1])
n_active = (j.get('per_gpu') or j['config']['per_gpu'])['active_edges']
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; un = rows[1]
col = hdr.index
f = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
tot, per = 0, collections.OrderedDict()
for r in rows[2:]:
    name = r[col('Kernel Name')].split('(')[0].replace('void ', '')
    b = float(r[col('dram__bytes_read.sum')]) * f[un[col('dram__bytes_read.sum')]] + float(r[col('dram__bytes_write.sum')]) * f[un[col('dram__bytes_write.sum')]]
    tot += b
    per[name] = {'dram_bytes': b, 'duration_us': float(r[col('gpu__time_duration.sum')])}
json.dump({"source": "ncu --set full --clock-control none, one steady-state iteration of `python tools/prof_iter.py 128` (the bench workload: 128 cfg2 events, gtf_iterate_dry)",
           "active_edges": n_active, "dram_bytes_per_iteration": tot, "dram_bytes_per_active_edge": tot / n_active, "kernels": per},
          open(os.path.join(P, RND + '_pipeline_traffic.json'), 'w'), indent=1)
open(os.path.join(P, RND + '_pipeline_ncu_summary.txt'), 'w').write(
    subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
d = collections.OrderedDict()
for r in rows[1:]:
    try: d.setdefault(r[ki].split('(')[0].replace('void ', ''), []).append(float(r[vi].replace(',', '')))
    except ValueError: pass
allt = sum(sum(v) for v in d.values())
L = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-loop (set-up + iterations)",
     "kernel                     launches   mean us    share of all launch time"]
for k, v in d.items(): L.append("%-26s %5d %10.1f %8.1f %%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / allt))
pipe = [k for k in d if k.split('<')[0] in ('k_begin', 'k_send', 'k_exec', 'k_node2', 'k_hv', 'k_big')]
med = lambda v: sorted(v)[len(v) // 2]      # noqa: E731  (the set-up's one seed-dict pass of k_node2 / k_hv is not an iteration)
s = sum(med(d[k]) for k in pipe)
L.append("one iteration (sum of the pipeline kernels' MEDIAN durations, serialised under ncu): %.1f us" % (s / 1e3))
for k in pipe: L.append("  share of the iteration  %-12s %5.1f %%   (median %.1f us)" % (k, 100 * med(d[k]) / s, med(d[k]) / 1e3))
open(os.path.join(P, RND + '_launches_bench_summary.txt'), 'w').write("\n".join(L) + "\n")
import shutil
shutil.copy(launches, os.path.join(P, RND + '_launches_bench.csv'))
shutil.copy(bench_json, os.path.join(P, RND + '_bench_line.json'))
hs = []
for skip, mangled, title, n in ((1, '_Z6k_send', 'k_send', 16), (2, '_Z6k_exec', 'k_exec', 16), (3, '_Z7k_node2', 'k_node2', 12), (5, '_Z4k_hvILi16E', 'k_hv<16>', 16)):
    env = dict(os.environ, NCU_SKIP=str(skip))
    o = subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_lines.py"), rep, so, mangled, str(n)], capture_output=True, text=True, env=env).stdout
    hs.append("## %s (ncu source page joined with -lineinfo: %% of warp-stall samples, %% of executed warp instructions, top stall reasons)\n%s" % (title, o))
open(os.path.join(P, RND + '_pipeline_source_hotspots.txt'), 'w').write("\n".join(hs))
print("iteration (ncu, serialised) %.1f us; dram %.1f MB = %.1f B per active edge" % (s / 1e3, tot / 1e6, tot / n_active))
print("\n".join(L[-10:]))
