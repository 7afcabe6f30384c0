"""Recognizer2D: (B, T, 3, H, W) clips -> backbone on (B*T, 3, H, W) frames -> head -> loss / scores.
Mirrors codes/models/recognizers/base.py:11-82 and recognizer2d.py:8-179 for the RGB + ResNet + MVF
configuration (configs/MVFNet/K400/*.py); the other module / backbone branches raise."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .builder import RECOGNIZERS, build_backbone, build_head


class BaseRecognizer(nn.Module):
    def __init__(self, backbone, cls_head):
        super().__init__()
        self.fp16_enabled = False
        self.backbone = build_backbone(backbone)
        if cls_head is not None:
            self.cls_head = build_head(cls_head)
        self.init_weights()

    @property
    def with_cls_head(self):
        return getattr(self, 'cls_head', None) is not None

    def init_weights(self):
        self.backbone.init_weights()
        if self.with_cls_head:
            self.cls_head.init_weights()

    def extract_feat(self, img_group):
        if img_group.is_cuda and img_group.dim() == 4:
            # the kernels' native activation layout is NHWC (torch channels_last); logical shape unchanged
            img_group = img_group.contiguous(memory_format=torch.channels_last)
        return self.backbone(img_group)

    def average_clip(self, cls_score):
        """'prob': softmax then mean over clips; 'score': mean; None: untouched (base.py:47-74)."""
        if self.test_cfg is None:
            self.test_cfg = {'average_clips': None}
        if 'average_clips' not in self.test_cfg.keys():
            raise KeyError('"average_clips" must defined in test_cfg\'s keys')
        mode = self.test_cfg['average_clips']
        if mode not in ['score', 'prob', None]:
            raise ValueError(f'{mode} is not supported. Currently supported ones are ["score", "prob", None]')
        if mode == 'prob':
            return F.softmax(cls_score, dim=1).mean(dim=0, keepdim=True)
        if mode == 'score':
            return cls_score.mean(dim=0, keepdim=True)
        return cls_score

    def forward(self, img_group, label, return_loss=True, return_numpy=True, **kwargs):
        if return_loss:
            return self.forward_train(img_group, label, **kwargs)
        return self.forward_test(img_group, return_numpy, **kwargs)


@RECOGNIZERS.register_module
class Recognizer2D(BaseRecognizer):
    def __init__(self, modality='RGB', backbone='BNInception', cls_head='TSNClsHead', fcn_testing=False,
                 module_cfg=None, nonlocal_cfg=None, train_cfg=None, test_cfg=None):
        super().__init__(backbone, cls_head)
        self.fcn_testing, self.modality = fcn_testing, modality
        self.train_cfg, self.test_cfg, self.module_cfg = train_cfg, test_cfg, module_cfg
        if self.module_cfg:
            self._prepare_base_model(backbone, self.module_cfg, nonlocal_cfg)
        if modality != 'RGB':
            raise NotImplementedError('only the RGB modality of the MVFNet configs is built')
        self.in_channels = 3

    def _prepare_base_model(self, backbone, module_cfg, nonlocal_cfg):
        """Splice the temporal module into the backbone (recognizer2d.py:45-100).  Like the reference,
        `type` is popped from module_cfg in place, leaving the make_multi_view_fusion kwargs."""
        module_name = module_cfg.pop('type')
        self.module_name = module_name
        if backbone['type'] != 'ResNet' or module_name != 'MVF' or nonlocal_cfg:
            raise NotImplementedError('mvfnet_b200 builds ResNet + MVF; got backbone=%s module=%s nonlocal=%s'
                                      % (backbone['type'], module_name, bool(nonlocal_cfg)))
        from .mvf import make_multi_view_fusion
        make_multi_view_fusion(self.backbone, **module_cfg)

    def forward_train(self, imgs, labels, **kwargs):
        num_batch = imgs.shape[0]
        imgs = imgs.reshape((-1, self.in_channels) + imgs.shape[3:])
        num_seg = imgs.shape[0] // num_batch
        x = self.extract_feat(imgs)
        losses = dict()
        if self.with_cls_head:
            temporal_pool = imgs.shape[0] // x.shape[0]
            from . import tail
            if tail.head_loss_eligible(x, self.cls_head, labels):
                # pool -> dropout -> Linear -> consensus -> cross-entropy on the library's kernels (tail.py)
                losses['loss_cls'] = tail.head_loss(x, self.cls_head, labels, num_seg // temporal_pool)
            else:
                cls_score = self.cls_head(x, num_seg // temporal_pool)
                losses.update(self.cls_head.loss(cls_score.float(), labels.squeeze()))
        return losses

    def forward_test(self, imgs, return_numpy, **kwargs):
        num_batch = imgs.shape[0]
        imgs = imgs.reshape((-1, self.in_channels) + imgs.shape[3:])
        num_frames = imgs.shape[0] // num_batch
        x = self.extract_feat(imgs)
        cls_score = x
        if self.with_cls_head:
            temporal_pool = imgs.shape[0] // x.shape[0]
            if self.module_cfg:
                seg = self.module_cfg['n_segment'] // temporal_pool
                if self.fcn_testing:
                    x = x.reshape((-1, seg) + x.shape[1:]).transpose(1, 2)      # (clips, C, T, h, w)
                cls_score = self.cls_head(x, seg)
            else:
                cls_score = self.cls_head(x, num_frames // temporal_pool)
            cls_score = self.average_clip(cls_score.float())
        return cls_score.cpu().numpy() if return_numpy else cls_score
