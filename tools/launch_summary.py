"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv [steps]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = re.sub(r'\(.*', '', r[kn]); name = re.sub(r'^void ', '', name)
    agg[name[:100]][0] += 1; agg[name[:100]][1] += v
tot = sum(v[1] for v in agg.values())
print('%d launches, %.1f us total, per step (/%g): %.1f us' % (sum(v[0] for v in agg.values()), tot / 1e3, steps, tot / 1e3 / steps))
ours = sum(v[1] for k, v in agg.items() if k.startswith(('gin::', 'tc::', 'tcp::', 'tcw', 'cv2::', 'wg2::', 'narrow::', 'bn::', 'head::', 'dist::')) or '::gin::' in k or 'gin::' in k)
print('kernels of this repo: %.1f%% of the device time' % (100 * ours / tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print('%9.1f us/step %6.1f n/step %5.1f%%  %s' % (v[1] / 1e3 / steps, v[0] / steps, 100 * v[1] / tot, k))
