"""Image backbone (stays in PyTorch; excluded from the timed hot path, reported separately).

Reference: GMF_PointDSC/models/Img_Encoder.py:9-18 wraps a torchvision-style ResNet-34
(models/resnet.py:59-216) whose forward stops after `layer2` (resnet.py:195-216) and returns
[B,128,H/8,W/8].  The parameter tree keeps the full ResNet-34 (layer3/4/fc are constructed but
never executed) so that a reference checkpoint's `encoder.image_encoder.backbone.*` keys load
unchanged.  No pretrained download is attempted (no network; BASELINE.json uses random init).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _Block(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + (x if self.downsample is None else self.downsample(x)))


def _stage(cin: int, cout: int, n: int, stride: int) -> nn.Sequential:
    return nn.Sequential(*[_Block(cin if i == 0 else cout, cout, stride if i == 0 else 1) for i in range(n)])


class ResNet34Trunk(nn.Module):
    """ResNet-34 parameter tree; forward = conv1/bn1/relu/maxpool/layer1/layer2 only."""

    def __init__(self, in_channels: int = 3):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = _stage(64, 64, 3, 1)
        self.layer2 = _stage(64, 128, 4, 2)
        self.layer3 = _stage(128, 256, 6, 2)     # constructed for state_dict parity, never run
        self.layer4 = _stage(256, 512, 3, 2)     # idem
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, 1000)           # idem
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer2(self.layer1(x))


class ImageEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = ResNet34Trunk(3)

    def forward(self, x):
        return self.backbone(x)

    @torch.no_grad()
    def tokens(self, image: torch.Tensor) -> torch.Tensor:
        """[B,3,H,W] -> [B,(H/8)*(W/8),128] row-major image tokens (PointDSC.py:129-135)."""
        f = self.backbone(image)
        b, c, h, w = f.shape
        return f.view(b, c, h * w).permute(0, 2, 1).contiguous()

    # ---- opt-in accelerated trunk: channels_last + bf16 autocast + CUDA graph ---------------------------------
    # Still PyTorch / cuDNN (the backbone is outside the hot path by BASELINE.json); eval-mode only.  The reference-equivalent fp32 eager trunk
    # above stays the default: `gmf_b200.PointDSC(..., ).backbone_mode = "bf16_graph"` switches.  The graph is keyed on the input shape and
    # replays with static input / output buffers; weights are read from the module's parameters at capture time, so `invalidate_fast()`
    # must be called (the module does it) when they change.
    def invalidate_fast(self):
        self._fast = {}

    @torch.no_grad()
    def tokens_fast(self, image: torch.Tensor, use_graph: bool = True) -> torch.Tensor:
        if self.training:
            raise RuntimeError("the accelerated trunk is eval-mode only (BatchNorm running statistics)")
        if image.device.type != "cuda":
            raise RuntimeError("the accelerated trunk needs a CUDA device")
        cache = self.__dict__.setdefault("_fast", {})
        key = (tuple(image.shape), image.device.index)
        ent = cache.get(key)

        def run(x):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                f = self.backbone(x)
            b, c, h, w = f.shape
            return f.float().permute(0, 2, 3, 1).reshape(b, h * w, c).contiguous()     # NHWC feature map == token-major layout

        if ent is None:
            if cache.get("_cl") is not True:                   # one-time: channels_last weights (no numerical change)
                self.backbone.to(memory_format=torch.channels_last)
                cache["_cl"] = True
            static_in = torch.empty(image.shape, device=image.device, dtype=torch.float32).contiguous(memory_format=torch.channels_last)
            static_in.copy_(image)
            if not use_graph:
                return run(static_in)
            side = torch.cuda.Stream(device=image.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up outside capture (cuDNN autotune, lazy init)
                for _ in range(2):
                    run(static_in)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = run(static_in)
            ent = cache[key] = (graph, static_in, static_out)
        graph, static_in, static_out = ent
        static_in.copy_(image)
        graph.replay()
        return static_out.clone()
