"""Weights holder for the engine's policy/value network.

PyTorch is used only to hold the parameters: `AlphaZeroNet` reproduces the parameter names and shapes
of the reference module (/root/reference/alpha_zero/core/network.py:85-156) so that reference
checkpoints (`torch.load(...)['network']`) load unchanged and `state_dict()` can be handed to
Engine.set_weights(), which folds BatchNorm and copies everything into the kernels' own buffers.
`forward` exists for the learner side (training is out of scope of the engine) and as a convenience;
the self-play path never calls it.
"""
import torch
from torch import nn


def _conv_bn(cin, cout, k, pad, relu):
    layers = [nn.Conv2d(cin, cout, k, 1, pad, bias=False), nn.BatchNorm2d(cout)]
    if relu:
        layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class _Residual(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.conv_block1 = _conv_bn(width, width, 3, 1, True)
        self.conv_block2 = _conv_bn(width, width, 3, 1, False)

    def forward(self, x):
        return torch.relu(self.conv_block2(self.conv_block1(x)) + x)


class AlphaZeroNet(nn.Module):
    def __init__(self, input_shape, num_actions, num_res_block=19, num_filters=256, num_fc_units=256, gomoku=False):
        super().__init__()
        planes, h, w = input_shape
        pad = 3 if gomoku else 1  # network.py:101
        side_h, side_w = h + 2 * pad - 2, w + 2 * pad - 2
        cells = side_h * side_w
        self.geometry = dict(planes=planes, board=(h, w), canvas=(side_h, side_w), actions=num_actions, blocks=num_res_block,
                             filters=num_filters, fc=num_fc_units, gomoku=gomoku)
        self.conv_block = _conv_bn(planes, num_filters, 3, pad, True)
        self.res_blocks = nn.Sequential(*[_Residual(num_filters) for _ in range(num_res_block)])
        self.policy_head = nn.Sequential(*_conv_bn(num_filters, 2, 1, 0, True), nn.Flatten(), nn.Linear(2 * cells, num_actions))
        self.value_head = nn.Sequential(*_conv_bn(num_filters, 1, 1, 0, True), nn.Flatten(), nn.Linear(cells, num_fc_units), nn.ReLU(),
                                        nn.Linear(num_fc_units, 1), nn.Tanh())
        for m in self.modules():  # network.py:31-39
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_uniform_(m.weight, nonlinearity='relu')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        feat = self.res_blocks(self.conv_block(x))
        return self.policy_head(feat), self.value_head(feat)


def randomize_batchnorm(net, seed=321):
    """Give BatchNorm non-trivial statistics so that BN folding is exercised (tests / synthetic benchmarks)."""
    gen = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) * 1.5 + 0.25)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=gen) * 1.0 + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=gen) * 0.2)
    return net
