"""Splits an `ncu --page source --print-source sass --csv` dump at BAR.SYNC and prints per-phase instruction / stall shares."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data, seen = [], set()
for r in rows[2:]:
    if len(r) > 20 and r[0].startswith('0x'):
        if r[0] in seen: break
        seen.add(r[0]); data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
def I(x):
    try: return int(x)
    except: return 0
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
segs, cur = [], None
def new(k): return dict(inst=0, wf=0, wfx=0, samples=0, start=k, ops={}, st={})
cur = new(0)
for k, r in enumerate(data):
    s = r[ci['Source']].strip()
    inst = I(r[ci['Instructions Executed']])
    cur['inst'] += inst; cur['wf'] += I(r[ci['L1 Wavefronts Shared']]); cur['wfx'] += I(r[ci['L1 Wavefronts Shared Excessive']]); cur['samples'] += I(r[ci['# Samples']])
    t = s.split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    cur['ops'][op] = cur['ops'].get(op, 0) + inst
    for h in stalls: cur['st'][h] = cur['st'].get(h, 0) + I(r[ci[h]])
    if 'BAR.SYNC' in s or 'EXIT' in s:
        cur['end'] = k; segs.append(cur); cur = new(k + 1)
tot = sum(s['inst'] for s in segs); tots = sum(s['samples'] for s in segs)
print('total warp inst', tot, 'samples', tots, 'sass lines', len(data))
for i, s in enumerate(segs):
    top = sorted(s['ops'].items(), key=lambda x: -x[1])[:8]
    st = sorted(s['st'].items(), key=lambda x: -x[1])[:4]
    print('seg%d [%d-%d] inst=%.1f%% samples=%.1f%% wf=%.1fM wfx=%.1fM | %s | %s' % (i, s['start'], s['end'], 100 * s['inst'] / tot, 100 * s['samples'] / max(tots, 1), s['wf'] / 1e6, s['wfx'] / 1e6,
          ' '.join('%s:%.1f' % (o, 100 * c / tot) for o, c in top), ' '.join('%s:%.1f' % (o[6:], 100 * c / max(tots, 1)) for o, c in st)))
