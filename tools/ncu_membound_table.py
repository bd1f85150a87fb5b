"""Text table from an ncu report of the memory-bound kernels (tools/gpu/r02_final_n1.sh: SpeedOfLight + MemoryWorkloadAnalysis +
Occupancy + LaunchStats sections and the DRAM byte counters):

    python tools/ncu_membound_table.py gpurun_out/r02_membound.ncu-rep > profiles/r02_membound_ncu.txt

GB/s = (dram read + write bytes) / duration under ncu (serialised, cold cache): a lower bound on the rate inside the replayed step.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'second': 1.0}
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
peak = 6546.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass


def val(r, name):
    if name not in col:
        return float('nan')
    try:
        return float(r[col[name]].replace(',', '')) * UNIT.get(units[col[name]], 1.0)
    except ValueError:
        return float('nan')


print('# %s: one launch per row, in launch order; HBM peak %.0f GB/s (MEASURED_PEAKS.json)' % (os.path.basename(sys.argv[1]), peak))
print('# %-58s %8s %8s %8s %7s %6s %5s %6s %6s %6s %6s' % ('kernel', 'us', 'rd MB', 'wr MB', 'GB/s', 'of pk', 'regs', 'grid', 'waves', 'occ %', 'L2hit'))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r'^void ', '', re.sub(r'\(.*', '', r[col['Kernel Name']]))[:58]
    t, rd, wr = val(r, WANT[0]), val(r, WANT[1]), val(r, WANT[2])
    gbs = (rd + wr) / t / 1e9 if t > 0 else float('nan')
    print('%-60s %8.1f %8.1f %8.1f %7.0f %5.0f%% %5.0f %6.0f %6.2f %6.1f %6.1f' % (name, t * 1e6, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak, val(r, WANT[3]), val(r, WANT[4]),
                                                                                  val(r, WANT[5]), val(r, WANT[6]), val(r, WANT[7])))
