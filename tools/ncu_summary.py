"""Condense an ncu capture of one bench step into profiles/<name>.json and profiles/conv_traffic.json.

On the GPU box (one timed step only, cudaProfilerStart/Stop inside bench.py); either the per-metric log
    VFS_BENCH_CUPROFILE=1 ncu --clock-control none --profile-from-start off --metrics <METRICS> --csv \
        --log-file gpurun_out/step_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
(METRICS = the keys of WANT below, comma separated) or a full report converted with
    ncu -i gpurun_out/step.ncu-rep --page raw --csv > gpurun_out/step_raw.csv
Here (no GPU needed):
    python tools/ncu_summary.py gpurun_out/step_metrics.csv profiles/r01_ncu_step_v10.json "<source note>"
"""
import csv
import json
import os
import sys

WANT = {
    'gpu__time_duration.sum': 'time_us',
    'dram__bytes_read.sum': 'dram_read_bytes',
    'dram__bytes_write.sum': 'dram_write_bytes',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct_alt',
    'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active': 'tensor_pct_alt2',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_pct',
    'sm__inst_issued.avg.pct_of_peak_sustained_active': 'issue_pct',
    'launch__grid_size': 'grid',
    'launch__registers_per_thread': 'regs',
    'launch__cluster_size': 'cluster',
}
UNIT_SCALE = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3, 'second': 1e6,
              'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return None


def main():
    src, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    with open(src, newline='') as fh:
        rows = [r for r in csv.reader(fh) if r]
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    header = rows[start]
    col = {name: i for i, name in enumerate(header)}
    name_col = col['Kernel Name']
    kernels = []
    if 'Metric Name' in col:   # long format of `--metrics ... --csv --log-file`: one row per (launch, metric)
        by_id = {}
        for r in rows[start + 1:]:
            if len(r) != len(header):
                continue
            k = by_id.setdefault(r[col['ID']], {'kernel': r[name_col].replace('void vfs::', '').replace('vfs::', '')
                                                .split('(')[0]})
            key = WANT.get(r[col['Metric Name']])
            v = num(r[col['Metric Value']])
            if key is None or v is None:
                continue
            if key == 'time_us' or key.endswith('_bytes'):
                v *= UNIT_SCALE.get(r[col['Metric Unit']], 1.0)
            k[key] = v
        kernels = [by_id[i] for i in sorted(by_id, key=int)]
        rows = []
    units, data = (rows[start + 1], rows[start + 2:]) if rows else ([], [])
    tensor_cols = [n for n in header if 'pipe_tensor' in n and 'pct' in n]
    for r in data:
        if len(r) != len(header):
            continue
        k = {'kernel': r[name_col].replace('void vfs::', '').replace('vfs::', '').split('(')[0]}
        for metric, key in WANT.items():
            if metric in col:
                v = num(r[col[metric]])
                if v is None:
                    continue
                u = units[col[metric]]
                if key == 'time_us' or key.endswith('_bytes'):
                    v *= UNIT_SCALE.get(u, 1.0)
                k[key] = v
        if 'tensor_pct' not in k:
            for n in tensor_cols:
                v = num(r[col[n]])
                if v is not None:
                    k['tensor_pct'] = v
                    k['tensor_metric'] = n
                    break
        kernels.append(k)
    conv = [k for k in kernels if k['kernel'].startswith('conv_tc_kernel')]
    conv_bytes = sum(k.get('dram_read_bytes', 0) + k.get('dram_write_bytes', 0) for k in conv)
    conv_time = sum(k.get('time_us', 0) for k in conv)
    out = {
        'source': note,
        'kernels': kernels,
        'conv_launches': len(conv),
        'conv_dram_bytes_per_step': conv_bytes,
        'conv_time_weighted_tensor_pipe_active_pct':
            (sum(k.get('tensor_pct', 0) * k.get('time_us', 0) for k in conv) / conv_time) if conv_time else None,
        'sum_time_us': sum(k.get('time_us', 0) for k in kernels),
    }
    with open(dst, 'w') as fh:
        json.dump(out, fh, indent=1)
    with open(os.path.join(os.path.dirname(dst), 'conv_traffic.json'), 'w') as fh:
        json.dump({'dram_bytes_per_step': conv_bytes, 'launches': len(conv), 'source': dst}, fh, indent=1)
    print(f'{len(kernels)} launches, {len(conv)} conv launches, conv DRAM bytes/step {conv_bytes / 1e9:.3f} GB, '
          f'time-weighted tensor-pipe active {out["conv_time_weighted_tensor_pipe_active_pct"]}')


if __name__ == '__main__':
    main()
