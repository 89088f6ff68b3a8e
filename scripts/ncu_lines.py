"""Aggregate an ncu source-page export (cuda,sass CSV) by region/line (dev tool)."""
import csv, collections, sys, subprocess, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0]); src = {}
reasons = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == 'File Path': cur_file = r[1].split('/')[-1]
    elif r and r[0] == 'Line No': hdr = r
    elif r and r[0].isdigit() and hdr:
        off = len(r) - len(hdr)
        li = int(r[0])
        k = hdr.index('Warp Stall Sampling (All Samples)'); ie = hdr.index('Instructions Executed')
        sv = r[k + off]; iv = r[ie + off]
        if sv in ('-', ''): continue
        key = (cur_file, li)
        agg[key][0] += int(sv); agg[key][1] += int(iv) if iv not in ('-', '') else 0
        src[key] = ','.join(r[1:2 + off])[:90]
        for j, h in enumerate(hdr):
            if h.startswith('stall_') and 'Not Issued' not in h:
                v = r[j + off]
                if v not in ('-', ''): reasons[key][h[6:]] += int(v)
tot = sum(v[0] for v in agg.values())
print('total samples', tot)
# function map from source file: find enclosing function by scanning the file
import os
fn_of = {}
for f in set(k[0] for k in agg):
    path = None
    for cand in ('qpth_b200/csrc/' + f,):
        if os.path.exists(cand): path = cand
    if not path: continue
    cur = '?'
    for i, line in enumerate(open(path), 1):
        m = re.match(r'^(?:template.*\n)?\s*(?:__device__|__global__|static|inline|__forceinline__|\w+\s)+.*?\b(\w+)\s*\(', line)
        if (line.startswith('__device__') or line.startswith('__global__') or line.startswith('template') or re.match(r'^k_\w+\(', line)) :
            m2 = re.search(r'\b(\w+)\s*\(', line)
            if m2 and m2.group(1) not in ('__launch_bounds__', 'template'): cur = m2.group(1)
        fn_of[(f, i)] = cur
byfn = collections.Counter(); fr = collections.defaultdict(collections.Counter)
for key, v in agg.items():
    fn = fn_of.get(key, key[0])
    byfn[fn] += v[0]
    fr[fn].update(reasons[key])
for fn, v in byfn.most_common(16):
    print('%5.1f%%  %-24s %s' % (100 * v / tot, fn, ', '.join('%s %.0f%%' % (h, 100 * c / max(v, 1)) for h, c in fr[fn].most_common(4))))
print()
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print('%5.1f%% %9d  %s:%d [%s] %s' % (100 * v[0] / tot, v[1], key[0], key[1], fn_of.get(key, '?'), src[key]))
