"""Drop-in for the reference's module/pcgrad.py (`PCGrad(optimizer, reduction)`, :13-141) plus the two toy
networks of its self-check demo (:144-162, :165-195) so the demo's known-answer output can be reproduced."""
import torch
import torch.nn as nn
import torch.optim as optim

from mtdgan_b200.pcgrad import PCGrad  # noqa: F401


class TestNet(nn.Module):
    __test__ = False

    def __init__(self):
        super().__init__()
        self._linear = nn.Linear(3, 4)

    def forward(self, x):
        return self._linear(x)


class MultiHeadTestNet(nn.Module):
    def __init__(self):
        super().__init__()
        self._linear = nn.Linear(3, 2)
        self._head1 = nn.Linear(2, 4)
        self._head2 = nn.Linear(2, 4)

    def forward(self, x):
        feat = self._linear(x)
        return self._head1(feat), self._head2(feat)


def run_demo(device="cuda"):
    """The reference's `__main__` self-check (:165-195); returns the gradients instead of printing them.
    Inputs and initial weights are drawn on the CPU with seed 4 exactly as the reference does, then moved."""
    out = []
    for net_cls, heads in ((TestNet, False), (MultiHeadTestNet, True)):
        torch.manual_seed(4)
        x, y = torch.randn(2, 3), torch.randn(2, 4)
        net = net_cls()
        net, x, y = net.to(device), x.to(device), y.to(device)
        pc_adam = PCGrad(optim.Adam(net.parameters()))
        pc_adam.zero_grad()
        if heads:
            y1, y2 = net(x)
            l1, l2 = nn.MSELoss()(y1, y), nn.MSELoss()(y2, y)
        else:
            yp = net(x)
            l1, l2 = nn.L1Loss()(yp, y), nn.MSELoss()(yp, y)
        pc_adam.pc_backward([l1, l2])
        out.append([p.grad.detach().cpu() for p in net.parameters()])
    return out


if __name__ == '__main__':
    for grads in run_demo():
        for g in grads:
            print(g)
        print('-' * 80)
