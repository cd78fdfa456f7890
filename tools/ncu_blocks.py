"""Basic-block view of an `ncu --page source --csv` export: consecutive SASS instructions with the same
execution count are merged; prints share of warp-instructions and of stall samples per block."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
hdr = rows[1]; data = rows[2:]
isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
ith = hdr.index("Thread Instructions Executed")
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[isamp]) for r in data)
print('total warp-instr %.3fG samples %d' % (tot / 1e9, tots))
blocks = []
for k, r in enumerate(data):
    ex = int(r[iex])
    if blocks and blocks[-1]['ex'] == ex:
        b = blocks[-1]
    else:
        b = dict(start=k, ex=ex, n=0, samples=0, ops=collections.Counter(), thr=0)
        blocks.append(b)
    b['n'] += 1; b['samples'] += int(r[isamp]); b['thr'] += int(r[ith])
    op = r[isrc].strip().split()
    op = op[1] if op[0].startswith('@') else op[0]
    b['ops'][op.split('.')[0]] += 1
for b in blocks:
    share = 100.0 * b['ex'] * b['n'] / tot
    if share < thr: continue
    ops = ' '.join(f"{o}:{c}" for o, c in b['ops'].most_common(12))
    print(f"@{b['start']:4d} n={b['n']:3d} x{b['ex']/1e6:8.2f}M  {share:5.2f}% instr  {100.0*b['samples']/tots:5.2f}% samples  thr/inst {b['thr']/max(1,b['ex']*b['n']):.1f} | {ops}")
