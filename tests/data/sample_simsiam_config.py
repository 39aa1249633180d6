# Test fixture written for this repository: same structure as the reference's configs/*.py (model dict selected by
# registry type names, train_cfg / test_cfg dicts, optimizer block) at toy sizes.
model = dict(
    type='SimSiamBaseTracker',
    backbone=dict(
        type='ResNet', pretrained=None, depth=18, out_indices=(3, ),
        norm_cfg=dict(type='SyncBN', requires_grad=True), norm_eval=False, zero_init_residual=True),
    img_head=dict(
        type='SimSiamHead', in_channels=512, norm_cfg=dict(type='SyncBN'), num_projection_fcs=3,
        projection_mid_channels=512, projection_out_channels=512, num_predictor_fcs=2, predictor_mid_channels=128,
        predictor_out_channels=512, with_norm=True, loss_feat=dict(type='CosineSimLoss', negative=False),
        spatial_type='avg'))
train_cfg = dict(intra_video=True)
test_cfg = dict(
    precede_frames=20, topk=10, temperature=0.07, strides=(1, 2, 1, 1), out_indices=(2, ), neighbor_range=24,
    with_first=True, with_first_neighbor=True, output_dir='eval_results')
optimizer = dict(type='SGD', lr=0.05, momentum=0.9, weight_decay=0.0001)
data = dict(videos_per_gpu=32, workers_per_gpu=16)
dist_params = dict(backend='nccl')
