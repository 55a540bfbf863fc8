"""`gaussian_splatting.utils.image_utils` (imported by utils/eval_utils_0806.py:27-29): mse / psnr [UPSTREAM-RECALL]."""
import torch


def mse(img1, img2):
    return (((img1 - img2)) ** 2).view(img1.shape[0], -1).mean(1, keepdim=True)


def psnr(img1, img2):
    mse_ = (((img1 - img2)) ** 2).view(img1.shape[0], -1).mean(1, keepdim=True)
    return 20 * torch.log10(1.0 / torch.sqrt(mse_))
