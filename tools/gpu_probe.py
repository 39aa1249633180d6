"""Quick on-box timing probe (not the bench): backbone forward and attention at the headline shapes."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from vfs_b200 import ops  # noqa: E402
from vfs_b200.backbones import ResNet  # noqa: E402
from vfs_b200.common import spatial_neighbor  # noqa: E402


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters


def main():
    res = {}
    net = ResNet(50, norm_cfg=dict(type='SyncBN', requires_grad=True), strides=(1, 2, 1, 1), out_indices=(2, ))
    net.load_state_dict(oracle.seeded_state_dict(net, seed=0))
    net = net.cuda()
    net.train(False)
    eng = net.engine
    for shape in ((16, 3, 256, 256), (10, 3, 480, 854), (1, 3, 480, 854)):
        x = torch.randn(shape, device='cuda')
        res[f'r50_res4_fwd_ms_{shape}'] = timed(lambda: eng.forward_split(x, 2))
        # host-side issue time (no sync) to see launch-bound behaviour
        torch.cuda.synchronize()
        t = time.perf_counter()
        eng.forward_split(x, 2)
        res[f'r50_res4_host_issue_ms_{shape}'] = (time.perf_counter() - t) * 1e3
        torch.cuda.synchronize()
    # per-layer timing at the 256x256x16 shape
    x = torch.randn(16, 3, 256, 256, device='cuda')
    layers = []
    orig_conv = eng.conv

    def conv_timed(cm, xs, relu, residual=None, want_f32=False):
        ms = timed(lambda: orig_conv(cm, xs, relu, residual, want_f32), warm=1, iters=3)
        k = cm.conv.kernel_size[0]
        _, N, H, W, Cin = xs.shape
        Ho, Wo = ops.conv_out_hw(H, W, k, cm.conv.stride[0], cm.conv.dilation[0])
        flops = 2.0 * N * Ho * Wo * cm.conv.out_channels * Cin * k * k
        layers.append(dict(k=k, s=cm.conv.stride[0], Cin=Cin, Cout=cm.conv.out_channels, H=H, W=W, ms=ms,
                           tflops=flops / ms / 1e9))
        return orig_conv(cm, xs, relu, residual, want_f32)

    eng.conv = conv_timed
    res['stem_ms_16x256'] = timed(lambda: eng.stem(x))
    eng.forward_split(x, 2)
    eng.conv = orig_conv
    res['layers_16x256'] = layers
    res['layers_sum_ms'] = sum(l['ms'] for l in layers)

    # attention 480p
    H, W, C, Cv = 60, 107, 1024, 4
    mask = spatial_neighbor(1, H, W, 36)
    for T in (1, 5, 21):
        bank = ops.features_to_split(torch.relu(torch.randn(T + 1, C, H, W, device='cuda')), True)
        vals = torch.rand(T + 1, Cv, H * W, device='cuda')
        ids = list(range(T))
        res[f'attn_480p_T{T}_ms'] = timed(lambda: ops.attention_bank(bank[:, T:T + 1], bank, ids, vals, Cv * H * W,
                                                                      H * W, Cv, mask, 0.07, 10))
    q = torch.relu(torch.randn(1, C, H, W, device='cuda'))
    res['features_to_split_480p_ms'] = timed(lambda: ops.features_to_split(q, True))
    lay = res.pop('layers_16x256')
    print(json.dumps(res, indent=1))
    for l in lay:
        print(f"k{l['k']} s{l['s']} {l['Cin']:5d}->{l['Cout']:5d} {l['H']:3d}x{l['W']:3d} {l['ms']*1e3:7.1f} us {l['tflops']:6.1f} TF")
    res['layers_16x256'] = lay
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', os.environ.get('PROBE_OUT', 'probe.json')), 'w') as fh:
        json.dump(res, fh, indent=1)


if __name__ == '__main__':
    main()
