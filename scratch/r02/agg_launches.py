"""ncu launch-list CSV (gpu__time_duration + dram bytes) -> per-kernel summary JSON + compact per-launch CSV for profiles/"""
import csv, re, collections, sys, json
src, out_json, out_csv = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src, errors='ignore')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
ix = {k: h.index(k) for k in ('ID', 'Kernel Name', 'Grid Size', 'Metric Name', 'Metric Unit', 'Metric Value')}
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= ix['Metric Value']:
        continue
    d = launch.setdefault(r[ix['ID']], {'kernel': r[ix['Kernel Name']], 'grid': r[ix['Grid Size']]})
    v = float(r[ix['Metric Value']].replace(',', ''))
    u = r[ix['Metric Unit']]
    name = r[ix['Metric Name']]
    if name.startswith('gpu__time'):
        d['us'] = v / 1000 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1000)
    else:
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        d['rd' if 'read' in name else 'wr'] = v * mult
def short(k):
    k = re.sub(r'^void ', '', k)
    m = re.match(r'(edadm::(?:g2::)?\w+)', k)
    if m:
        return m.group(1)
    if k.startswith('at::') or 'at::native' in k:
        return 'at::'
    return re.sub(r'[<(].*', '', k)[:60]
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(short(d['kernel']), {'launches': 0, 'us': 0.0, 'dram_read_MB': 0.0, 'dram_write_MB': 0.0})
    a['launches'] += 1; a['us'] += d.get('us', 0.0); a['dram_read_MB'] += d.get('rd', 0) / 1e6; a['dram_write_MB'] += d.get('wr', 0) / 1e6
tot = sum(a['us'] for a in agg.values())
for a in agg.values():
    a['share'] = round(a['us'] / tot, 4); a['us'] = round(a['us'], 1)
    a['dram_bytes_per_launch'] = int((a['dram_read_MB'] + a['dram_write_MB']) * 1e6 / a['launches'])
    a['dram_read_MB'] = round(a['dram_read_MB'], 1); a['dram_write_MB'] = round(a['dram_write_MB'], 1)
json.dump({'total_us': round(tot, 1), 'kernels': dict(sorted(agg.items(), key=lambda kv: -kv[1]['us']))}, open(out_json, 'w'), indent=1)
with open(out_csv, 'w') as f:
    f.write('id,kernel,grid,us,dram_read_bytes,dram_write_bytes\n')
    for i, d in enumerate(launch.values()):
        f.write(f"{i},\"{d['kernel'][:90]}\",\"{d['grid']}\",{d.get('us', 0):.2f},{int(d.get('rd', 0))},{int(d.get('wr', 0))}\n")
print('total us', round(tot, 1))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us'])[:18]:
    print(f"{a['us']:9.1f} us {a['launches']:5d} {100*a['share']:5.1f}%  rd {a['dram_read_MB']:9.1f} MB wr {a['dram_write_MB']:9.1f} MB  {k}")
