"""slow_r50 video backbone in stock PyTorch — OUTSIDE the hot path (north_star: "The slow_r50 video backbone stays on
stock PyTorch").

The reference obtains it through ``torch.hub.load('facebookresearch/pytorchvideo', 'slow_r50')`` and replaces the
classification head by an identity (video_encoder.py:30-51), so ``encode(clip[B,3,T,224,224])`` returns the res5
feature map ``[B,2048,T,7,7]``.  There is no network here, so the architecture (SlowFast paper's slow pathway on
ResNet-50: no temporal stride, temporal 3x1x1 kernels in res4/res5) is defined locally with the SAME module names as
pytorchvideo's (``blocks.0.conv``, ``blocks.N.res_blocks.M.branch2.conv_a`` ...), so a pretrained hub checkpoint —
and the ``vid_encoder.backbone.*`` keys of a reference ``.pth`` — load with ``strict=True`` (the head block is an
``nn.Identity`` exactly as in the reference).  Used by ``compat.AGQAHotPath(vid_encoder=VideoBackbone())`` and by
``bench.py``'s ``with_backbone`` line (SURVEY §8f-4 hygiene: bf16 autocast + channels_last_3d, frozen / eval)."""
import torch
from torch import nn


class _Stem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv3d(3, 64, (1, 7, 7), stride=(1, 2, 2), padding=(0, 3, 3), bias=False)
        self.norm = nn.BatchNorm3d(64)
        self.activation = nn.ReLU()
        self.pool = nn.MaxPool3d((1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))

    def forward(self, x):
        return self.pool(self.activation(self.norm(self.conv(x))))


class _Bottleneck(nn.Module):
    def __init__(self, c_in, c_inner, c_out, temporal, stride):
        super().__init__()
        self.conv_a = nn.Conv3d(c_in, c_inner, (temporal, 1, 1), padding=(temporal // 2, 0, 0), bias=False)
        self.norm_a = nn.BatchNorm3d(c_inner)
        self.act_a = nn.ReLU()
        self.conv_b = nn.Conv3d(c_inner, c_inner, (1, 3, 3), stride=(1, stride, stride), padding=(0, 1, 1), bias=False)
        self.norm_b = nn.BatchNorm3d(c_inner)
        self.act_b = nn.ReLU()
        self.conv_c = nn.Conv3d(c_inner, c_out, 1, bias=False)
        self.norm_c = nn.BatchNorm3d(c_out)

    def forward(self, x):
        x = self.act_a(self.norm_a(self.conv_a(x)))
        x = self.act_b(self.norm_b(self.conv_b(x)))
        return self.norm_c(self.conv_c(x))


class _ResBlock(nn.Module):
    def __init__(self, c_in, c_inner, c_out, temporal, stride):
        super().__init__()
        if c_in != c_out or stride != 1:
            self.branch1_conv = nn.Conv3d(c_in, c_out, 1, stride=(1, stride, stride), bias=False)
            self.branch1_norm = nn.BatchNorm3d(c_out)
        self.branch2 = _Bottleneck(c_in, c_inner, c_out, temporal, stride)
        self.activation = nn.ReLU()

    def forward(self, x):
        short = self.branch1_norm(self.branch1_conv(x)) if hasattr(self, "branch1_conv") else x
        return self.activation(self.branch2(x) + short)


class _ResStage(nn.Module):
    def __init__(self, depth, c_in, c_inner, c_out, temporal, stride):
        super().__init__()
        self.res_blocks = nn.ModuleList([_ResBlock(c_in if i == 0 else c_out, c_inner, c_out, temporal,
                                                   stride if i == 0 else 1) for i in range(depth)])

    def forward(self, x):
        for b in self.res_blocks:
            x = b(x)
        return x


class SlowR50(nn.Module):
    """blocks.0 stem | blocks.1-4 res2..res5 | blocks.5 Identity (the reference drops the head)."""

    def __init__(self):
        super().__init__()
        self.blocks = nn.ModuleList([
            _Stem(),
            _ResStage(3, 64, 64, 256, temporal=1, stride=1),
            _ResStage(4, 256, 128, 512, temporal=1, stride=2),
            _ResStage(6, 512, 256, 1024, temporal=3, stride=2),
            _ResStage(3, 1024, 512, 2048, temporal=3, stride=2),
            nn.Identity()])

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x

    def encode(self, x):
        return self.forward(x)


class VideoBackbone(nn.Module):
    """Same surface as video_encoder.py's VideoBackbone for backbone_name='slow_r50': ``.backbone`` holds the network
    (state_dict keys ``backbone.blocks...``), ``encode(clip)`` -> ``[B,2048,T,7,7]``.  autocast=True runs it under
    bf16 autocast with channels_last_3d activations; freeze=True puts it in eval mode without
    gradients, as the STAR script does (star.py:87-88) — the AGQA script only calls ``.eval()`` (agqaHGQA.py:263)."""

    def __init__(self, backbone_name="slow_r50", autocast=True, freeze=False):
        super().__init__()
        if backbone_name != "slow_r50":
            raise NotImplementedError("only slow_r50 (the README's --backbone) is defined locally")
        self.backbone_name = backbone_name
        self.backbone = SlowR50().eval()
        self.autocast, self.freeze = autocast, freeze
        if freeze:
            for p in self.backbone.parameters():
                p.requires_grad_(False)

    def encode(self, x):
        x = x.contiguous(memory_format=torch.channels_last_3d)
        on = self.autocast and x.is_cuda
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=on):
            if self.freeze:
                with torch.no_grad():
                    y = self.backbone(x)
            else:
                y = self.backbone(x)
        return y.float()

    def forward(self, x):
        return self.encode(x)
