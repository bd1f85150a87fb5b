"""Summarise .ncu-rep files (read on the CPU box): python tools/ncu_summary.py report.ncu-rep [--src N]"""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum']


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name'), 'grid', d.get('launch__grid_size'))
        for k in KEYS:
            if k in d:
                print('   %-72s %s %s' % (k, d[k], units[hdr.index(k)]))
        st = {k: float(v.replace(',', '')) for k, v in d.items() if k.startswith('smsp__pcsamp_warps_issue_stalled') and not k.endswith('not_issued') and v}
        tot = sum(st.values()) or 1
        print('   stalls: ' + ', '.join('%s %.0f%%' % (k.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / tot)
                                         for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))


def source(path, n):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] in ('Address', '#')]
    if not hi:
        print(out[:2000]); return
    hdr = rows[hi[0]]
    try:
        ci = hdr.index('# Samples') if '# Samples' in hdr else [i for i, h in enumerate(hdr) if 'Sampling' in h and 'All' in h][0]
    except Exception:
        print(hdr); return
    si = hdr.index('Source')
    body = [r for r in rows[hi[0] + 1:] if len(r) > ci]
    def num(v):
        try: return float(v.replace(',', ''))
        except Exception: return 0.0
    tot = sum(num(r[ci]) for r in body) or 1
    top = sorted(range(len(body)), key=lambda i: -num(body[i][ci]))[:n]
    print('   total samples', tot)
    for i in sorted(top):
        print('   %5d %5.1f%%  %s' % (i, 100 * num(body[i][ci]) / tot, body[i][si][:150]))


if __name__ == '__main__':
    n = 0
    args = sys.argv[1:]
    if '--src' in args:
        n = int(args[args.index('--src') + 1]); args = [a for a in args if a not in ('--src', str(n))]
    for p in args:
        raw(p)
        if n:
            source(p, n)
