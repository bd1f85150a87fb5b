"""Render tests/diag/precision_study.py JSON files as one markdown report:

    python tools/precision_md.py fp16fwd=gpurun_out/r02b_precision_study.json bf16fwd=gpurun_out/precision_study.json > profiles/r02_precision_study.md
"""
import json
import sys

runs = [a.split('=', 1) for a in sys.argv[1:]]
data = {k: json.load(open(v)) for k, v in runs}
first = next(iter(data.values()))
cfg = first['config']
print('# Precision study: 16-bit-operand tcgen05 path vs exact fp32 (`tests/diag/precision_study.py`)\n')
print('I%d, batch %d, %d Adam steps (lr %g) over %d distinct batches, %s.  `fp32` = module-wise exact-fp32 CUDA-core path (`impl=\'simt\'`), '
      'which tracks the CPU oracle to cosine 1.00000.  Runs: %s.\n' % (cfg['level'], cfg['batch'], cfg['steps'], cfg['lr'], cfg['pool_batches'], cfg['gpu'],
                                                                    ', '.join('**%s**' % k for k in data)))
for name in ('ico2ico', 'ico2ico_vae'):
    print('## %s\n' % name)
    print('### Total training loss per step (same init, data order and reparameterisation noise)\n')
    hdr = ['step'] + ['%s fused' % k for k in data] + ['fp32']
    print('| ' + ' | '.join(hdr) + ' |')
    print('|' + '---|' * len(hdr))
    steps = len(first[name]['curves']['fp32_simt'])
    for i in list(range(0, steps, 20)) + [steps - 1]:
        row = ['%d' % i] + ['%.5f' % d[name]['curves']['bf16_fused'][i][-1] for d in data.values()] + ['%.5f' % first[name]['curves']['fp32_simt'][i][-1]]
        print('| ' + ' | '.join(row) + ' |')
    print()
    for k, d in data.items():
        c = d[name]['curves']
        print('* %s: mean of the last 10 steps %.5f (fp32 %.5f), largest relative difference over the run %.3f, %.1f ms/step eager (fp32 path %.1f ms/step)' % (
            k, c['mean_last10']['bf16_fused'], c['mean_last10']['fp32_simt'], c['max_rel_diff_total'], c['sec_per_step']['bf16_fused'] * 1e3,
            c['sec_per_step']['fp32_simt'] * 1e3))
    print()
    print('### Per-parameter gradient cosine against the CPU oracle (one held-out batch)\n')
    print('| run | state | path | min cosine (parameter) | mean cosine | loss | oracle loss |')
    print('|---|---|---|---|---|---|---|')
    for k, d in data.items():
        for tag in ('random_init', 'conditioned'):
            e = d[name][tag]
            for path, c in e.get('vs_oracle', {}).items():
                print('| %s | %s | %s | %.5f (%s) | %.5f | %.6f | %.6f |' % (k, tag, path, c['min_cos'], c['worst'], c['mean_cos'], e['loss'][path], e['loss']['oracle_cpu_fp32']))
    print()
