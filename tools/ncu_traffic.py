"""Extract per-launch DRAM traffic of the dominant kernels from ncu reports into profiles/<round>_traffic.json
(bench.py reports it as roofline.traffic).  Run here, no GPU needed:
    python tools/ncu_traffic.py r01 gpurun_out/prof_spmm_r01d.ncu-rep gpurun_out/prof_tc_r01e.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    rnd, reps = sys.argv[1], sys.argv[2:]
    out = {'how': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, yelp-lightgcn (tools/prof_step.py)',
           'kernels': {}}
    for rep in reps:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, check=True).stdout.decode()
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ni, ri, wi, ti = (hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                                                 'gpu__time_duration.sum'))
        for r in rows[2:]:
            name = r[ni]
            key = ('igcn_spmm' if 'prop_kernel' in name and name.rstrip(')').split(',')[-1].strip().startswith('0') and ', 0, 0>' in name
                   else 'igcn_spmm_rows' if ', 0, 1>' in name else 'igcn_spmm_cols' if ', 0, 2>' in name
                   else 'igcn_tc_candidates' if 'score_tc' in name else name.split('(')[0])
            b = float(r[ri].replace(',', '')) * UNIT[units[ri]] + float(r[wi].replace(',', '')) * UNIT[units[wi]]
            e = out['kernels'].setdefault(key, {'launches': 0, 'dram_bytes': 0.0, 'report': os.path.basename(rep)})
            e['launches'] += 1
            e['dram_bytes'] += b
    for e in out['kernels'].values():
        e['dram_bytes_per_launch'] = e.pop('dram_bytes') / e['launches']
    path = os.path.join(ROOT, 'profiles', rnd + '_traffic.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print(path, json.dumps(out['kernels']))


if __name__ == '__main__':
    main()
