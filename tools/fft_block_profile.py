"""One Res-FFT-Conv block forward + backward at the train shape (B = 20, 64 x 64, C = 32), three times: a target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtdgan_b200 import ops
torch.manual_seed(0)
C = 32
x = torch.randn(20, 64, 64, C, device="cuda", requires_grad=True)
img_w = (torch.randn(C, C, 3, 3, device="cuda") * 0.05).requires_grad_(True)
img_b = torch.zeros(C, device="cuda", requires_grad=True)
fft_w = (torch.randn(2 * C, 2 * C, 1, 1, device="cuda") * 0.05).requires_grad_(True)
fft_b = torch.zeros(2 * C, device="cuda", requires_grad=True)
for _ in range(3):
    y = ops.fft_conv_block(x, img_w, img_b, fft_w, fft_b)
    y.sum().backward()
torch.cuda.synchronize()
