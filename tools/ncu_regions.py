"""Group the SASS lines of an `ncu --page source --csv` export into regions of equal execution count (loop bodies)
and print each region's share of the issued warp instructions and its average active lanes.
    python tools/ncu_regions.py source.csv [min_share_percent]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows[:5]):
    if 'Source' in r:
        hdr, start = r, i + 1
        break
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start:] if len(r) > ix['Source']]
def f(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return 0.0
ti = sum(f(r, 'Instructions Executed') for r in data); tt = sum(f(r, 'Thread Instructions Executed') for r in data)
print('SASS lines', len(data), 'warp inst %.3e thread inst %.3e lanes %.2f' % (ti, tt, tt / ti))
reg, cur = [], None
for n, r in enumerate(data):
    ie, te = f(r, 'Instructions Executed'), f(r, 'Thread Instructions Executed')
    sh, ln = ie / ti * 100, te / max(ie, 1)
    if cur and abs(cur['ie'] - sh) <= 0.15 * max(cur['ie'], sh, 1e-9):
        cur['n2'] = n; cur['sum'] += sh; cur['tl'] += te; cur['il'] += ie
    else:
        if cur: reg.append(cur)
        cur = {'n1': n, 'n2': n, 'ie': sh, 'sum': sh, 'tl': te, 'il': ie, 'src': r[ix['Source']][:70]}
reg.append(cur)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
for g in reg:
    if g['sum'] > thr:
        print('lines %4d-%4d (%3d)  per-line %.3f%%  share %5.2f%%  lanes %4.1f  first: %s' % (g['n1'], g['n2'], g['n2'] - g['n1'] + 1, g['ie'], g['sum'], g['tl'] / max(g['il'], 1), g['src']))
