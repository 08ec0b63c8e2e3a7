"""LPIPS (AlexNet, v0.1): mirror of src/criteria/lpips/{lpips.py:8-35, networks.py:23-95, utils.py:6-8}.

Same module tree and state-dict keys (``net.layers.{0,3,6,8,10}.{weight,bias}``, ``net.mean``, ``net.std``,
``lin.{0..4}.1.weight``), so a checkpoint assembled from torchvision's AlexNet and the LPIPS linear layers loads
unchanged.  The reference constructor downloads both (networks.py:77 ``models.alexnet(True)``, utils.py:11-19); nothing
is fetched here - load weights with ``load_state_dict`` (benchmarks use seeded stand-ins, e4s_b200/synthetic.py).

The convolutions are plain library convolutions (cuDNN): the loss networks are standard CNNs and SURVEY.md section 2 #14
keeps them out of the hand-written hot path; what this package adds is the caching of the target image's features.
"""
from typing import List, Sequence

import torch
import torch.nn as nn


def normalize_activation(x: torch.Tensor, eps: float = 1e-10) -> torch.Tensor:
    """x / (sqrt(sum_c x^2 + 1e-16) + eps), utils.py:6-8."""
    norm_factor = torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True) + 1e-16)
    return x / (norm_factor + eps)


class LinLayers(nn.ModuleList):
    def __init__(self, n_channels_list: Sequence[int]):
        super().__init__([nn.Sequential(nn.Identity(), nn.Conv2d(nc, 1, 1, 1, 0, bias=False)) for nc in n_channels_list])
        for param in self.parameters():
            param.requires_grad = False


class AlexNet(nn.Module):
    """torchvision ``alexnet().features`` (same indices) with the activations after ReLU 1-5 as outputs (networks.py:74-83)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("mean", torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("std", torch.Tensor([.458, .448, .450])[None, :, None, None])
        self.layers = nn.Sequential(
            nn.Conv2d(3, 64, kernel_size=11, stride=4, padding=2), nn.ReLU(inplace=True), nn.MaxPool2d(kernel_size=3, stride=2),
            nn.Conv2d(64, 192, kernel_size=5, padding=2), nn.ReLU(inplace=True), nn.MaxPool2d(kernel_size=3, stride=2),
            nn.Conv2d(192, 384, kernel_size=3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(384, 256, kernel_size=3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(256, 256, kernel_size=3, padding=1), nn.ReLU(inplace=True), nn.MaxPool2d(kernel_size=3, stride=2))
        self.target_layers = [2, 5, 8, 10, 12]
        self.n_channels_list = [64, 192, 384, 256, 256]
        for param in self.parameters():
            param.requires_grad = False

    def z_score(self, x: torch.Tensor) -> torch.Tensor:
        return (x - self.mean) / self.std

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        x = self.z_score(x)
        output = []
        for i, layer in enumerate(self.layers, 1):
            x = layer(x)
            if i in self.target_layers:
                output.append(normalize_activation(x))
            if len(output) == len(self.target_layers):
                break
        return output


class LPIPS(nn.Module):
    def __init__(self, net_type: str = "alex", version: str = "0.1"):
        assert version in ["0.1"], "v0.1 is only supported now"
        if net_type != "alex":
            raise NotImplementedError("the inversion loop uses LPIPS(net_type='alex') (scripts/optimization.py:79)")
        super().__init__()
        self.net = AlexNet()
        self.lin = LinLayers(self.net.n_channels_list)

    def features(self, x: torch.Tensor) -> List[torch.Tensor]:
        """Unit-normalised activations of the five target layers (cache these for a fixed target image)."""
        return self.net(x)

    def distance(self, feat_x: List[torch.Tensor], feat_y: List[torch.Tensor], batch: int) -> torch.Tensor:
        diff = [(fx - fy) ** 2 for fx, fy in zip(feat_x, feat_y)]
        res = [l(d).mean((2, 3), True) for d, l in zip(diff, self.lin)]
        return torch.sum(torch.cat(res, 0)) / batch

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self.distance(self.net(x), self.net(y), x.shape[0])
