import csv, sys, collections, re
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
I = {h:i for i,h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
def num(x):
    try: return float(x.replace(',',''))
    except: return 0.0
tot_s = sum(num(r[I['# Samples']]) for r in data)
tot_i = sum(num(r[I['Instructions Executed']]) for r in data)
print(f"total samples {tot_s:.0f}, warp-instructions {tot_i:.3e}, SASS lines {len(data)}")
by_op = collections.defaultdict(lambda:[0,0,0])
for r in data:
    op = r[I['Source']].strip().split()
    name = op[1] if op and op[0].startswith('@') and len(op)>1 else (op[0] if op else '?')
    name = name.split('.')[0]
    by_op[name][0] += num(r[I['# Samples']]); by_op[name][1] += num(r[I['Instructions Executed']]); by_op[name][2]+=1
print("\nby opcode: samples%  instr%  (static count)")
for k,(s,i,c) in sorted(by_op.items(), key=lambda kv:-kv[1][0])[:28]:
    print(f"  {k:12s} {100*s/tot_s:6.2f}% {100*i/tot_i:6.2f}%  ({c})")
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(num(r[I[h]]) for r in data) for h in stalls}
print("\nstall reasons (all samples):")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:10]: print(f"  {k:24s} {100*v/tot_s:6.2f}%")
print("\ntop 25 SASS lines by samples:")
for r in sorted(data, key=lambda r:-num(r[I['# Samples']]))[:25]:
    top = sorted(stalls, key=lambda h:-num(r[I[h]]))[:2]
    print(f"  {100*num(r[I['# Samples']])/tot_s:5.2f}%  {r[I['Source']][:70]:70s} {top[0]}={r[I[top[0]]]} {top[1]}={r[I[top[1]]]}")
