"""Top SASS instructions of an `ncu --page source --csv` export by stall samples and by executed count, with opcode totals."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
def num(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
tot_s = sum(num(r, '# Samples') for r in data); tot_i = sum(num(r, 'Instructions Executed') for r in data)
print('instructions', len(data), 'samples', tot_s, 'warp-instr executed', tot_i)
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
agg = {n: sum(num(r, n) for r in data) for n in stalls}
print('stall totals:', ', '.join(f'{k[6:]} {v / max(tot_s, 1) * 100:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
ops = collections.Counter()
for r in data:
    ops[r[ix['Source']].split()[0] if not r[ix['Source']].strip().startswith('@') else r[ix['Source']].split()[1]] += num(r, 'Instructions Executed')
print('opcodes:', ', '.join(f'{k} {v / tot_i * 100:.1f}%' for k, v in ops.most_common(14)))
print('--- top by samples')
for i, r in sorted(enumerate(data), key=lambda ir: -num(ir[1], '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = max(stalls, key=lambda n: num(r, n))
    print(f'{i:5d} {num(r, "# Samples") / max(tot_s, 1) * 100:5.1f}%  exec {num(r, "Instructions Executed"):9.0f}  {top[6:]:12s} {r[ix["Source"]].strip()[:90]}')
