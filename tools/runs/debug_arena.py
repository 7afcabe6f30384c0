import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mvfnet_b200 import build_recognizer, ops
from mvfnet_b200.graph import GraphedTrainStep
from mvfnet_b200.tail import FlatSGD, preprocess_frames
from mvfnet_b200.utils import to_channels_last

def cfg(p):
    return dict(type="Recognizer2D",
                backbone=dict(type="ResNet", pretrained=None, depth=50, out_indices=(3,), norm_eval=False,
                              partial_norm=False, norm_cfg=dict(type="BN", requires_grad=True)),
                cls_head=dict(type="TSNClsHead", spatial_size=-1, spatial_type="avg", with_avg_pool=False,
                              temporal_feature_size=1, spatial_feature_size=1, dropout_ratio=p, in_channels=2048,
                              init_std=0.01, num_classes=400),
                module_cfg=dict(type="MVF", n_segment=4, alpha=0.125, mvf_freq=(0, 0, 1, 1), mode="THW"))
torch.manual_seed(0)
g = torch.Generator().manual_seed(1)
imgs = [torch.randint(0, 256, (2, 4, 64, 64, 3), generator=g, dtype=torch.uint8).cuda() for _ in range(4)]
lbls = [torch.randint(0, 400, (2, 1), generator=g).cuda() for _ in range(4)]
orig = ops._stat_sums
for mode in ("plain", "arena", "plain2"):
    if mode.startswith("plain"):
        ops._stat_sums = lambda n, device: torch.zeros((2, n), dtype=torch.float32, device=device)
    else:
        ops._stat_sums = orig
    torch.manual_seed(0)
    m3 = to_channels_last(build_recognizer(cfg(0.5), None, None).cuda()).train()
    o3 = FlatSGD(m3.parameters(), lr=0.0, momentum=0.0, weight_decay=0.0, nesterov=False, max_norm=40)
    out = []
    for i in range(3):
        o3.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = m3(preprocess_frames(imgs[0]), lbls[0])["loss_cls"]
        loss.backward()
        o3.step(1)
        out.append(loss.item())
    print(mode, "eager", out, "grad norm", float(o3.grad_norm))
