"""Summarise an `ncu --metrics ... --csv` log: one line per (kernel, grid) with averaged metrics."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = collections.OrderedDict()
for r in csv.DictReader(lines):
    key = r['ID']
    d = rows.setdefault(key, {'name': r['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:34], 'grid': r['Grid Size']})
    try:
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
        d[r['Metric Name'] + '#u'] = r['Metric Unit']
    except ValueError:
        pass
agg = collections.OrderedDict()
for d in rows.values():
    k = (d['name'], d['grid'])
    a = agg.setdefault(k, collections.defaultdict(float))
    a['n'] += 1
    for m, v in d.items():
        if isinstance(v, float):
            a[m] += v
def us(a):
    return a['gpu__time_duration.sum'] / a['n'] / 1e3
top = sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum'])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
print(f'{"kernel":34s} {"grid":>14s} {"n":>4s} {"us":>8s} {"totms":>7s} {"rdMB":>7s} {"wrMB":>7s} {"GB/s":>6s} {"dram%":>6s} {"sm%":>5s} {"tens%":>6s} {"warps%":>6s} {"stsec/req":>9s} {"ldsec/req":>9s} {"bankconf":>9s}')
for (name, grid), a in top:
    n = a['n']
    t = us(a)
    rd, wr = a['dram__bytes_read.sum'] / n / 1e6, a['dram__bytes_write.sum'] / n / 1e6
    def ratio(x, y):
        return a[x] / a[y] if a[y] else 0.0
    print(f'{name:34s} {grid:>14s} {int(n):4d} {t:8.1f} {t * n / 1e3:7.2f} {rd:7.1f} {wr:7.1f} {(rd + wr) / t * 1e3 / 1e3:6.0f} '
          f'{a["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"] / n:6.1f} {a["sm__throughput.avg.pct_of_peak_sustained_elapsed"] / n:5.1f} '
          f'{a["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] / n:6.1f} {a["sm__warps_active.avg.pct_of_peak_sustained_active"] / n:6.1f} '
          f'{ratio("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"):9.1f} '
          f'{ratio("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"):9.1f} '
          f'{a["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"] / n:9.0f}')
