"""Scratch: summarise an `ncu --page source --csv` dump (hot loop vs rest)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(r[ix[k]] or 0)
tot = sum(I(r, '# Samples') for r in data)
totinst = sum(I(r, 'Instructions Executed') for r in data)
print("total samples", tot, "total warp-inst", totinst)
mx = max(I(r, 'Instructions Executed') for r in data)
hot = [r for r in data if I(r, 'Instructions Executed') > 0.5 * mx]
cold = [r for r in data if I(r, 'Instructions Executed') <= 0.5 * mx]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for name, grp in (("hot", hot), ("cold", cold)):
    print(name, "instrs", len(grp), "samples", sum(I(r, '# Samples') for r in grp), "warp-inst", sum(I(r, 'Instructions Executed') for r in grp))
    agg = {s: sum(I(r, s) for r in grp) for s in stalls}
    print("   ", sorted(agg.items(), key=lambda x: -x[1])[:7])
top = sorted(data, key=lambda r: -I(r, '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]
for r in top:
    print(r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']][:80])
