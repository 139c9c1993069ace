"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, bench_TAG.json} into the tracked summaries under profiles/."""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
src, dst = 'gpurun_out', 'profiles'
os.makedirs(dst, exist_ok=True)
out = []
with open(os.path.join(src, 'launches_%s.csv' % tag)) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
agg = collections.OrderedDict(); n = 0
for row in r:
    name = row[ki].split('(')[0].replace('void ', '')
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(row[vi].replace(',', '')); n += 1
tot = sum(a[1] for a in agg.values())
out.append('ncu launch list of ONE denoising step (bench.py --steps 1 --warmup 3 --eager, timed step bracketed by '
           'cudaProfilerStart/Stop; --metrics gpu__time_duration.sum --clock-control none)')
out.append('launches: %d   sum of device time: %.1f us   (cold-cache, serialised: compare SHARES)' % (n, tot / 1e3))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append('  %-48s n=%4d  %10.1f us  %5.1f%%' % (name[:48], a[0], a[1] / 1e3, 100 * a[1] / tot))
open(os.path.join(dst, '%s_launches_summary.txt' % tag), 'w').write('\n'.join(out) + '\n')
shutil.copy(os.path.join(src, 'launches_%s.csv' % tag), os.path.join(dst, '%s_launches.csv' % tag))
# full capture of the dominant kernel
raw = subprocess.run(['ncu', '-i', os.path.join(src, 'prof_%s.ncu-rep' % tag), '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_dynamic']
txt = ['ncu --set full --clock-control none --import-source on (same command), units row: ' +
       ', '.join('%s[%s]' % (h, rows[1][hdr.index(h)]) for h in want if h in hdr and rows[1][hdr.index(h)])]
for row in rows[2:]:
    txt.append('---')
    for h in want:
        if h in hdr:
            txt.append('  %-80s %s' % (h, row[hdr.index(h)][:90]))
open(os.path.join(dst, '%s_top_kernel_ncu.txt' % tag), 'w').write('\n'.join(txt) + '\n')
for f in ('bench_%s.json' % tag, 'bench_ref_%s.json' % tag):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
print('\n'.join(out)); print('\n'.join(txt))
