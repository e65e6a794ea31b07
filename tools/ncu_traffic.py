#!/usr/bin/env python
"""profiles/r02_conv_traffic.json from an `ncu --page raw --csv` export of the tower kernel: DRAM bytes per launch of the plain and
the residual layer (dram__bytes_read.sum + dram__bytes_write.sum), which bench.py scales to the tick's leaf count for
`roofline.traffic`.

    ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > profiles/r02_ncu_k_conv_tc_x_raw.csv
    python tools/ncu_traffic.py profiles/r02_ncu_k_conv_tc_x_raw.csv go9_c2/bf16/mode5 <leaves>
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    return float(v.replace(',', '')) * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main():
    path, key, leaves = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    ir, iw, it, ik = (h.index(k) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'Kernel Name'))
    per = []
    for r in rows[2:]:
        if 'k_conv_tc' not in r[ik]:
            continue
        per.append((to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]), to_bytes(r[ir], units[ir]), float(r[it].replace(',', ''))))
    assert len(per) >= 2, 'need one plain and one residual launch'
    per.sort()
    plain, res = per[0], per[-1]
    out_path = os.path.join(ROOT, 'profiles', 'r02_conv_traffic.json')
    try:
        out = json.load(open(out_path))
    except Exception:
        out = {}
    out[key] = {'leaves': leaves, 'bytes_plain': plain[0], 'bytes_residual': res[0], 'us_plain': plain[2], 'us_residual': res[2],
                'source': f'{os.path.relpath(path, ROOT)} (ncu --set full, {leaves} leaves per launch)'}
    json.dump(out, open(out_path, 'w'), indent=1)
    print(out[key])


if __name__ == '__main__':
    main()
