"""TSN wrapper around a (TSM-)ResNet focus network -- mirror of STH/models/tsn.py (TSN :11-127, forward :215-241).

`base_model` is a torchvision-style bottleneck ResNet (same child and parameter names as torchvision.models.resnet50 /
resnet101, STH/models/tsn.py:114) with TemporalShift wrappers on conv1 (:116-120).  Callers strip the fc after
construction -- `model.focuser.net.base_model = nn.Sequential(*children[:-1])`, STH/evaluate.py:83 -- and that surgery
is supported: the runner is rebuilt from whatever `base_model` currently is."""
import torch
from torch import nn

from ..engine import get_engine
from ..models.mobilenet import _param_key
from ..packcache import cached_runner
from ..models.resnet import ResNet, ResNetRunner
from .temporal_shift import make_temporal_shift

_LAYERS = {"resnet50": [3, 4, 6, 3], "resnet101": [3, 4, 23, 3]}


class TSN(nn.Module):
    def __init__(self, num_segments, modality="RGB", base_model="resnet50", new_length=None, crop_num=1,
                 partial_bn=True, print_spec=True, pretrain="imagenet", is_shift=False, shift_div=8,
                 shift_place="blockres", fc_lr5=False, temporal_pool=False, non_local=False):
        super().__init__()
        if modality != "RGB":
            raise NotImplementedError("only RGB (the shipped configuration) is implemented")
        if non_local:
            raise NotImplementedError
        if base_model not in _LAYERS:
            raise ValueError("Unknown base model: {}".format(base_model))
        self.modality, self.num_segments, self.reshape, self.crop_num = modality, num_segments, False, crop_num
        self.pretrain, self.is_shift, self.shift_div, self.shift_place = pretrain, is_shift, shift_div, shift_place
        self.base_model_name, self.fc_lr5, self.temporal_pool, self.non_local = base_model, fc_lr5, temporal_pool, False
        self.new_length = 1 if new_length is None else new_length
        self.base_model = ResNet(_LAYERS[base_model])
        if is_shift:
            make_temporal_shift(self.base_model, num_segments, n_div=shift_div, place=shift_place,
                                temporal_pool=temporal_pool)
        self.base_model.last_layer_name = "fc"
        self.base_model.avgpool = nn.AdaptiveAvgPool2d(1)
        self.input_size, self.input_mean, self.input_std = 224, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
        self._enable_pbn = partial_bn
        self._runner = None

    def runner(self):
        key = (id(self.base_model), _param_key(self.base_model))
        if self._runner is None or self._runner.key != key:
            self._runner = cached_runner(self.base_model, "ResNetRunner", lambda: ResNetRunner(self.base_model, key), key)
        return self._runner

    def forward(self, input, no_reshape=False):
        """(N,3,P,P) fp32 patches (no_reshape=True) or (B, 3T, P, P) -> base_model(input).squeeze(): (N, 2048) once
        the fc has been stripped (STH/models/tsn.py:215-241)."""
        if any(isinstance(m, nn.Linear) for m in self.base_model.children()):
            raise NotImplementedError("the fc head of fL is stripped by every inference caller (STH/evaluate.py:83); "
                                      "running it is stage-1 training territory")
        if not no_reshape:
            input = input.reshape((-1, 3 * self.new_length) + tuple(input.shape[-2:]))
        eng = get_engine(input.device)
        fmap = self.runner().run(eng, input.contiguous())
        n, h, w, c = fmap.shape
        out = torch.empty(n, c, dtype=torch.float32, device=input.device)
        eng.avgpool(fmap, out_f32=out, out_f32_stride=c)
        return out.view(n, c, 1, 1).squeeze()

    @property
    def feature_dim(self):
        return 2048

    @property
    def crop_size(self):
        return self.input_size

    @property
    def scale_size(self):
        return self.input_size * 256 // 224
