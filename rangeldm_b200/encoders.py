"""Condition encoders (`ldm/encoders.py:86-95`)."""
import torch


class SparseRangeImageEncoder2(torch.nn.Module):
    """(B, C, W, H) -> (B, 4C, W/4, H): four neighbouring azimuth columns become channels
    (channel = (w % 4) * C + c).  Pure index shuffle, no arithmetic."""

    def encode(self, x):
        return self(x)

    def forward(self, x):
        B, C, W, H = x.shape
        return x.reshape(B, C, W // 4, 4, H).permute(0, 3, 1, 2, 4).reshape(B, 4 * C, W // 4, H)
