#!/bin/bash
# DRAM bytes of every tc_conv_kernel launch of one training step (for roofline.traffic) + the default bench line
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    -k "regex:tc_conv_(pair_)?kernel" --csv --log-file gpurun_out/tc_conv_traffic.csv python scripts/profile_step.py > gpurun_out/traffic.log 2>&1
python - <<'PY'
import csv, json, collections
lines = [l for l in open('gpurun_out/tc_conv_traffic.csv') if not l.startswith('==')]
per = collections.defaultdict(dict)
for r in csv.DictReader(lines):
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3}.get(u, 1)
    per[r['ID']][r['Metric Name']] = v * scale
n = len(per)
rd = sum(p.get('dram__bytes_read.sum', 0) for p in per.values())
wr = sum(p.get('dram__bytes_write.sum', 0) for p in per.values())
t = sum(p.get('gpu__time_duration.sum', 0) for p in per.values())
out = {'tc_conv_kernel': {'launches': n, 'dram_bytes_per_launch': (rd + wr) / max(n, 1), 'dram_read_bytes_total': rd,
                          'dram_write_bytes_total': wr, 'duration_s_total_under_ncu': t},
       'how': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tc_conv_(pair_)?kernel over one eager training step (mnist DCGAN, batch 128, bf16)'}
json.dump(out, open('gpurun_out/roofline_traffic.json', 'w'), indent=1)
print(out)
PY
cp gpurun_out/roofline_traffic.json profiles/roofline_traffic.json
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_full.json; cut -c1-250 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json; cut -c1-200 gpurun_out/bench_reference.json
