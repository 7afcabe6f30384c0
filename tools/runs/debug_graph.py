import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mvfnet_b200 import build_recognizer
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

for px, B in ((64, 2), (128, 4)):
    torch.manual_seed(0)
    base = build_recognizer(cfg(0.0), None, None)
    g = torch.Generator().manual_seed(1)
    imgs = [torch.randint(0, 256, (B, 4, px, px, 3), generator=g, dtype=torch.uint8).cuda() for _ in range(4)]
    lbls = [torch.randint(0, 400, (B, 1), generator=g).cuda() for _ in range(4)]
    def fresh():
        m = to_channels_last(copy.deepcopy(base).cuda()).train()
        return m, FlatSGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True, max_norm=40)
    def eager():
        m, o = fresh(); out = []
        for img, lbl in zip(imgs, lbls):
            o.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = m(preprocess_frames(img), lbl)["loss_cls"]
            loss.backward(); o.step(1); out.append(loss.item())
        return out
    e1, e2 = eager(), eager()
    m2, o2 = fresh()
    step = GraphedTrainStep(m2, o2, imgs[0], lbls[0], warmup=0)
    gr = [step(i, l).item() for i, l in zip(imgs, lbls)]
    print(px, B, "eager1", e1); print(px, B, "eager2", e2); print(px, B, "graph ", gr)
