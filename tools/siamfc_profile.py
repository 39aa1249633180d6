import cProfile, pstats, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import vfs_b200
from vfs_b200.siamfc import TrackerSiamFC, build_cfg
from vfs_b200.synthetic import seeded_state_dict
cfg = build_cfg(dict(type='ResNet', depth=18, pretrained=None, norm_cfg=dict(type='BN', requires_grad=True)), exemplar_sz=127, out_scale=1e-3)
trk = TrackerSiamFC(cfg)
trk.net.backbone.load_state_dict(seeded_state_dict(trk.net.backbone, seed=3))
trk.net.head.load_state_dict(seeded_state_dict(trk.net.head, seed=4))
trk.net.to('cuda')
rng = np.random.RandomState(0)
frames = [rng.randint(0, 256, (480, 640, 3)).astype(np.uint8) for _ in range(4)]
trk.init(frames[0], [300, 200, 80, 60])
for i in range(5): trk.update(frames[i % 4])
pr = cProfile.Profile(); pr.enable()
for i in range(30): trk.update(frames[i % 4])
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
