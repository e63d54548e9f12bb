#!/bin/bash
# End-of-round validation of HEAD, most important first: whole GPU suite, smoke, default bench line, reference arm,
# full-size parity report, launch list, ncu --set full per-launch table of the tensor-core convolution family.
O=gpurun_out/r02; mkdir -p $O
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6 | tee $O/test_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $O/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/bench_mnist_bf16_1gpu.json; cut -c1-400 $O/bench_mnist_bf16_1gpu.json
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $O/bench_reference_cpu.json; cut -c1-200 $O/bench_reference_cpu.json
echo "== full-size parity"; timeout 900 python scripts/fullsize_parity_report.py > $O/fullsize_parity.log 2>&1; tail -4 $O/fullsize_parity.log; cp gpurun_out/fullsize_parity.json $O/ 2>/dev/null
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_mnist_bf16.csv python scripts/profile_step.py > $O/prof.log 2>&1
python scripts/summarize_launches.py $O/launches_mnist_bf16.csv > $O/launch_summary_mnist_bf16.txt; head -14 $O/launch_summary_mnist_bf16.txt
echo "== ncu full: tensor-core convolution launches"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:tc_conv_(pair_|shift_)?kernel" -c 40 -o /tmp/r02_tc_conv -f python scripts/profile_step.py > $O/ncu_tc_conv.log 2>&1; tail -1 $O/ncu_tc_conv.log
ncu -i /tmp/r02_tc_conv.ncu-rep --page raw --csv > /tmp/r02_tc_conv_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py /tmp/r02_tc_conv_raw.csv > $O/ncu_per_launch_tc_conv.txt; head -30 $O/ncu_per_launch_tc_conv.txt
python - <<'PY'
import csv, json
rows = list(csv.reader(open('/tmp/r02_tc_conv_raw.csv')))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = float(r[ix[name]].replace(',', '')); u = units[ix[name]]
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
rd = sum(val(r, 'dram__bytes_read.sum') for r in rows[2:]); wr = sum(val(r, 'dram__bytes_write.sum') for r in rows[2:])
n = len(rows) - 2
out = {'tc_conv_kernel': {'launches': n, 'dram_bytes_per_launch': (rd + wr) / max(n, 1), 'dram_read_bytes_total': rd, 'dram_write_bytes_total': wr},
       'how': 'ncu --set full -k regex:tc_conv_(pair_|shift_)?kernel over one eager training step (mnist DCGAN, batch 128, bf16), round 2'}
json.dump(out, open('gpurun_out/r02/roofline_traffic.json', 'w'), indent=1); print(out)
PY
ncu -i /tmp/r02_tc_conv.ncu-rep --page details --kernel-name regex:tc_conv_shift --launch-skip 0 --launch-count 2 > $O/ncu_tc_conv_shift_details.txt 2>/dev/null
