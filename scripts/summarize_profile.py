"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, bench_TAG.json} into the tracked summaries under profiles/."""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32      # per-GPU batch of the profiled run (recorded with the traffic)
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
for f in ('bench_%s.json' % tag, 'bench_%s_k20.json' % tag, 'bench_ref_%s.json' % tag):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
print('\n'.join(out)); print('\n'.join(txt))

# ---- extras (when present): per-launch sections of every conv GEMM, gn_apply / attention captures, other bench lines
def raw_rows(rep):
    r = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(r.splitlines()))
    return (rr[0], rr[1], rr[2:]) if len(rr) > 2 else (None, None, [])


sec = os.path.join(src, 'sections_convs_%s.ncu-rep' % tag)
if os.path.exists(sec):
    h, u, rows_ = raw_rows(sec)
    cols = ['launch__grid_size', 'launch__cluster_size', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'launch__shared_mem_per_block_dynamic']
    cols = [c for c in cols if c in h]
    with open(os.path.join(dst, '%s_conv_sections.csv' % tag), 'w') as f:
        w = csv.writer(f)
        w.writerow(['launch'] + ['%s [%s]' % (c, u[h.index(c)]) for c in cols])
        tot_t = tot_b = 0.0
        for i, r in enumerate(rows_):
            w.writerow([i] + [r[h.index(c)] for c in cols])
            tot_t += float(r[h.index('gpu__time_duration.sum')]) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(
                u[h.index('gpu__time_duration.sum')].replace('second', 's').replace('usecond', 'us'), 1.0)   # -> us
            if 'dram__bytes_read.sum' in h:
                def tobytes(v, unit):
                    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
                tot_b += tobytes(r[h.index('dram__bytes_read.sum')], u[h.index('dram__bytes_read.sum')]) + \
                    tobytes(r[h.index('dram__bytes_write.sum')], u[h.index('dram__bytes_write.sum')])
    n = len(rows_)
    json.dump({'kernel': 'conv_gemm_kernel', 'per_gpu_batch': batch, 'launches': n, 'sum_gpu_time_us': tot_t, 'dram_bytes_per_launch_avg': tot_b / max(n, 1),
               'dram_bytes_per_step': tot_b, 'source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d launches of one step' % n},
              open(os.path.join(dst, '%s_conv_traffic.json' % tag), 'w'), indent=1)
    print('conv sections: %d launches, %.1f us, %.1f MB DRAM traffic per launch' % (n, tot_t, tot_b / max(n, 1) / 1e6))

for kname in ('gn', 'attn'):
    rep = os.path.join(src, 'prof_%s_%s.ncu-rep' % (kname, tag))
    if not os.path.exists(rep):
        continue
    h, u, rows_ = raw_rows(rep)
    want2 = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
             'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
             'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
             'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
             'smsp__inst_executed.sum', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
             'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
    t2 = ['ncu --set full --clock-control none --import-source on; units: ' +
          ', '.join('%s[%s]' % (x, u[h.index(x)]) for x in want2 if x in h and u[h.index(x)])]
    for r in rows_:
        t2.append('---')
        for x in want2:
            if x in h:
                t2.append('  %-80s %s' % (x, r[h.index(x)][:90]))
    open(os.path.join(dst, '%s_%s_kernel_ncu.txt' % (tag, kname)), 'w').write('\n'.join(t2) + '\n')
for f in ('bench_emnist_%s.json' % tag, 'bench_in128_%s.json' % tag, 'layers_%s.txt' % tag, 'bench_n2_%s.json' % tag,
          'bench_ref_n2_%s.json' % tag):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
