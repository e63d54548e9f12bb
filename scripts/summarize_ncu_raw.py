"""Per-launch table from `ncu --page raw --csv`: duration, DRAM bytes, DRAM/L2/SM throughput %, tensor pipe %."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=''):
    i = idx.get(name)
    return r[i] if i is not None and i < len(r) else default


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return float('nan')


units = rows[1]
print('# columns: ms | dram_read_MB | dram_write_MB | dram% | l2% | sm% | tensor% | grid | kernel')
tot = 0.0
out = []
for r in rows[2:]:
    t = num(get(r, 'gpu__time_duration.sum'))
    tu = units[idx['gpu__time_duration.sum']] if 'gpu__time_duration.sum' in idx else 'ns'
    ms = t * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(tu, 1e-6)
    tot += ms

    def mb(name):
        v = num(get(r, name))
        u = units[idx[name]] if name in idx else 'byte'
        return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)

    name = re.sub(r'\(.*', '', get(r, 'Kernel Name'))[:60]
    out.append((ms, mb('dram__bytes_read.sum'), mb('dram__bytes_write.sum'),
                num(get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')),
                num(get(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')),
                num(get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed')),
                num(get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')),
                get(r, 'Grid Size'), name))
print(f'# {len(out)} launches, {tot:.3f} ms total (under ncu: serialised, cold cache)')
for o in sorted(out, key=lambda o: -o[0]):
    print(f'{o[0]:8.3f} | {o[1]:9.1f} | {o[2]:9.1f} | {o[3]:5.1f} | {o[4]:5.1f} | {o[5]:5.1f} | {o[6]:5.1f} | {o[7]:>14s} | {o[8]}')
