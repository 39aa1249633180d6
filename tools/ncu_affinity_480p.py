"""Run the 480p restricted-attention bench (bench.bench_affinity_480p, T = 1 and T = 21) for an ncu capture of its
kernels' DRAM traffic:

    ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv \
        -k regex:attn --log-file gpurun_out/affinity_480p_ncu.csv python tools/ncu_affinity_480p.py

and, without a GPU, condense that log into profiles/affinity_traffic.json (read by bench.py: `roofline.traffic`):

    python tools/ncu_affinity_480p.py --summarise gpurun_out/affinity_480p_ncu.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summarise(path):
    with open(path, newline='') as fh:
        rows = [r for r in csv.reader(fh) if r]
    start = next(i for i, r in enumerate(rows) if r[0] == 'ID')
    col = {n: i for i, n in enumerate(rows[start])}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3,
             'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}
    launches = {}
    for r in rows[start + 1:]:
        if len(r) != len(rows[start]):
            continue
        k = launches.setdefault(int(r[col['ID']]), {'kernel': r[col['Kernel Name']].split('(')[0]})
        v = float(r[col['Metric Value']].replace(',', '')) * scale.get(r[col['Metric Unit']], 1.0)
        k[r[col['Metric Name']]] = v
    ls = [launches[i] for i in sorted(launches)]
    # launch order: [T = 1: (scores, merge) x n1] then [T = 21: (scores, merge) x n21]; the score kernel's time tells
    # the two apart (T = 21 is ~10x longer)
    scores = [k for k in ls if 'scores_topk' in k['kernel']]
    merges = [k for k in ls if 'merge' in k['kernel']]
    tmax = max(k['gpu__time_duration.sum'] for k in scores)
    out = {'source': os.path.basename(path), 'note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, '
           'cold cache: every launch is replayed after a cache flush), scores+top-k kernel plus merge+propagate kernel'}
    for tag, sel in (('T1', lambda k: k['gpu__time_duration.sum'] < 0.3 * tmax),
                     ('T21', lambda k: k['gpu__time_duration.sum'] >= 0.3 * tmax)):
        idx = [i for i, k in enumerate(scores) if sel(k)]
        if not idx:
            continue
        sb = [scores[i]['dram__bytes_read.sum'] + scores[i]['dram__bytes_write.sum'] for i in idx]
        mb = [merges[i]['dram__bytes_read.sum'] + merges[i]['dram__bytes_write.sum'] for i in idx if i < len(merges)]
        out[f'{tag}_dram_bytes'] = sum(sb) / len(sb) + (sum(mb) / len(mb) if mb else 0.0)
        out[f'{tag}_scores_kernel_dram_bytes'] = sum(sb) / len(sb)
        out[f'{tag}_scores_kernel_us_under_ncu'] = sum(scores[i]['gpu__time_duration.sum'] for i in idx) / len(idx)
        out[f'{tag}_launches_captured'] = len(idx)
    with open(os.path.join(ROOT, 'profiles', 'affinity_traffic.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == '--summarise':
        return summarise(sys.argv[2])
    import torch
    import bench
    dev = torch.device('cuda', 0)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for T in (1, 21):
        r = bench.bench_affinity_480p(torch, dev, T, {}, 1678.8, flush)
        print(T, r['us_per_frame'])


if __name__ == '__main__':
    main()
