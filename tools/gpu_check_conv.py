"""On-box bring-up check of the tcgen05 conv path against (a) the fp32 SIMT instrument on the device and
(b) torch CPU fp64 convolution of the reconstructed (hi+lo) operands.  Every case runs in its own
subprocess with a timeout so a device trap in one case cannot poison the others.

    python tools/gpu_check_conv.py            # run all cases, write gpurun_out/conv_check.log
    python tools/gpu_check_conv.py --case 3   # run one case in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, N, H, W, Cin, Cout, k, stride, dil, relu, residual)
CASES = [
    ('layout_stem', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0),
    ('1x1_c64_bn64', 2, 16, 16, 64, 64, 1, 1, 1, 0, 0),
    ('1x1_c256_bn128', 2, 16, 16, 256, 128, 1, 1, 1, 1, 0),
    ('1x1_ragged_M', 1, 15, 13, 128, 256, 1, 1, 1, 1, 1),
    ('3x3_s1', 2, 16, 16, 64, 64, 3, 1, 1, 1, 0),
    ('3x3_s1_odd', 1, 15, 13, 64, 128, 3, 1, 1, 1, 1),
    ('3x3_s2', 2, 16, 16, 128, 128, 3, 2, 1, 1, 0),
    ('3x3_s2_odd', 1, 15, 13, 64, 64, 3, 2, 1, 0, 0),
    ('3x3_d2', 1, 20, 20, 64, 64, 3, 1, 2, 1, 0),
    ('3x3_d4', 1, 20, 20, 64, 64, 3, 1, 4, 1, 0),
    ('1x1_s2', 2, 16, 16, 256, 512, 1, 2, 1, 0, 0),
    ('3x3_davis', 1, 60, 107, 128, 128, 3, 1, 1, 1, 0),
    ('1x1_persistent', 8, 64, 64, 64, 256, 1, 1, 1, 1, 1),
    ('3x3_persistent', 8, 32, 32, 128, 128, 3, 1, 1, 1, 0),
    ('3x3_small7', 8, 7, 7, 512, 512, 3, 1, 1, 1, 0),
]


def run_case(idx):
    import torch
    import torch.nn.functional as F
    from vfs_b200 import ops

    name, N, H, W, Cin, Cout, k, stride, dil, relu, use_res = CASES[idx]
    torch.manual_seed(idx)
    dev = 'cuda'
    res = {'case': name}
    if name == 'layout_stem':
        x = torch.randn(2, 72, 9, 11)
        xs = ops.to_split(x.to(dev))
        back = ops.from_split(xs).cpu()
        res['layout_roundtrip_maxrel'] = float(((back - x).abs() / x.abs().clamp_min(1e-6)).max())
        img = torch.randn(2, 3, 67, 93)
        w = torch.randn(64, 3, 7, 7) * 0.1
        scale = torch.rand(64) + 0.5
        shift = torch.randn(64) * 0.1
        out = ops.from_split(ops.stem_forward(img.to(dev), w.to(dev), scale.to(dev), shift.to(dev))).cpu()
        ref = F.conv2d(img.double(), w.double(), stride=2, padding=3)
        ref = torch.relu(ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1))
        ref = F.max_pool2d(ref, 3, 2, 1)
        res['stem_shape_ok'] = list(out.shape) == list(ref.shape)
        res['stem_max_abs_err'] = float((out.double() - ref).abs().max())
        res['stem_ref_absmax'] = float(ref.abs().max())
        res['ok'] = bool(res['layout_roundtrip_maxrel'] < 1e-4 and res['stem_shape_ok'] and
                         res['stem_max_abs_err'] < 1e-4 * max(1.0, res['stem_ref_absmax']))
        return res

    x = torch.randn(N, Cin, H, W)
    w = torch.randn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout) + 0.5
    shift = torch.randn(Cout) * 0.1
    xs = ops.to_split(x.to(dev))
    wp = ops.pack_conv_weight(w.to(dev))
    Ho, Wo = ops.conv_out_hw(H, W, k, stride, dil)
    rs = None
    r = None
    if use_res:
        r = torch.randn(N, Cout, Ho, Wo)
        rs = ops.to_split(r.to(dev))
    sc, sh = scale.to(dev), shift.to(dev)
    t0 = time.time()
    out_split, out32 = ops.conv_bn_act(xs, wp, sc, sh, k, stride, dil, relu, rs, want_split=True, want_f32=True)
    torch.cuda.synchronize()
    res['tc_seconds_first_call'] = time.time() - t0
    simt = ops.debug_conv_bn_act_simt(xs, wp, sc, sh, k, stride, dil, relu, rs)
    torch.cuda.synchronize()
    # CPU fp64 reference on the reconstructed operands
    xr = ops.from_split(xs).cpu().double()
    wr = (wp[0].float() + wp[1].float()).cpu().double().view(Cout, k, k, Cin).permute(0, 3, 1, 2)
    pad = 0 if k == 1 else dil
    ref = F.conv2d(xr, wr, stride=stride, padding=pad, dilation=dil if k == 3 else 1)
    ref = ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    if use_res:
        ref = ref + ops.from_split(rs).cpu().double()
    if relu:
        ref = torch.relu(ref)
    ref_nhwc = ref.permute(0, 2, 3, 1).contiguous()
    o32 = out32.cpu().double()
    osplit = ops.from_split(out_split).cpu().double().permute(0, 2, 3, 1)
    sim = simt.cpu().double()
    denom = float(ref_nhwc.abs().max())
    res['ref_absmax'] = denom
    res['tc_f32_vs_ref_maxabs'] = float((o32 - ref_nhwc).abs().max())
    res['tc_split_vs_ref_maxabs'] = float((osplit - ref_nhwc).abs().max())
    res['simt_vs_ref_maxabs'] = float((sim - ref_nhwc).abs().max())
    res['tc_vs_simt_maxabs'] = float((o32 - sim).abs().max())
    res['ok'] = bool(res['tc_f32_vs_ref_maxabs'] < 2e-5 * max(1.0, denom) and
                     res['tc_split_vs_ref_maxabs'] < 5e-5 * max(1.0, denom))
    if not res['ok']:
        bad = (o32 - ref_nhwc).abs() > 1e-3 * max(1.0, denom)
        res['num_bad'] = int(bad.sum())
        res['num_total'] = int(bad.numel())
        idxs = bad.nonzero()[:8].tolist()
        res['first_bad'] = [(i, float(o32[tuple(i)]), float(ref_nhwc[tuple(i)])) for i in idxs]
        # error structure: which pixels / channels are wrong
        res['bad_pixels_frac'] = float(bad.any(dim=-1).float().mean())
        res['bad_channels_frac'] = float(bad.flatten(0, 2).any(dim=0).float().mean())
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', type=int, default=None)
    ap.add_argument('--timeout', type=int, default=120)
    a = ap.parse_args()
    if a.case is not None:
        print('RESULT ' + json.dumps(run_case(a.case)))
        return
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    log = open(os.path.join(ROOT, 'gpurun_out', 'conv_check.log'), 'w')
    n_ok = 0
    for i, c in enumerate(CASES):
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', str(i)], capture_output=True,
                               text=True, timeout=a.timeout)
            lines = [l for l in p.stdout.splitlines() if l.startswith('RESULT ')]
            if lines:
                line = lines[-1]
                n_ok += int(json.loads(line[7:]).get('ok', False))
            else:
                line = f'CASE {c[0]} FAILED rc={p.returncode}\nstdout: {p.stdout[-1500:]}\nstderr: {p.stderr[-2500:]}'
        except subprocess.TimeoutExpired:
            line = f'CASE {c[0]} TIMEOUT'
        print(line, flush=True)
        log.write(line + '\n')
        log.flush()
    summary = f'SUMMARY {n_ok}/{len(CASES)} ok'
    print(summary)
    log.write(summary + '\n')


if __name__ == '__main__':
    main()
