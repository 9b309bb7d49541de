"""Host-side view of the 128 x 64 bf16 swizzled tile images the MLP kernels exchange through HBM (debug / tests and
the Python side of the backward): image byte offset of (row r, 16-byte unit u) is r*128 + ((u ^ (r & 7)) << 4)."""
import torch


def _perm(device):
    r = torch.arange(128, device=device)[:, None]
    c = torch.arange(64, device=device)[None, :]
    unit = c // 8
    return (r * 64 + ((unit ^ (r & 7)) * 8) + c % 8).reshape(-1)          # logical (r, c) -> element index in the image


def decode(acts, slot, nblocks):
    """acts uint8 [tiles, slots, 16384] -> float32 [tiles*128, nblocks*64]."""
    tiles = acts.shape[0]
    perm = _perm(acts.device)
    out = []
    for b in range(nblocks):
        img = acts[:, slot + b].contiguous().view(torch.bfloat16).view(tiles, 128 * 64)
        out.append(img[:, perm].view(tiles, 128, 64))
    return torch.cat(out, dim=2).reshape(tiles * 128, nblocks * 64).float()


def encode(x, tiles):
    """float [tiles*128, 64*nb] -> uint8 images [tiles, nb, 16384] (inverse of decode)."""
    nb = x.shape[1] // 64
    perm = _perm(x.device)
    xb = x.to(torch.bfloat16).view(tiles, 128, nb, 64).permute(0, 2, 1, 3).reshape(tiles, nb, 128 * 64)
    img = torch.empty_like(xb)
    img[:, :, perm] = xb
    return img.view(torch.uint8).view(tiles, nb, 16384)


def add_masks(acts):
    """Append the mask images the forward kernel writes after the data images (csrc/common.cuh act_mask_offset): one 32-bit
    word per (image, 32-column group, row), bit i = value of column 32*group + i is non-zero."""
    from .nerf_program import act_tile_images
    tiles, slots, _ = acts.shape
    out = torch.zeros((tiles, act_tile_images(slots), 16384), dtype=torch.uint8, device=acts.device)
    out[:, :slots] = acts
    words = out[:, slots:].reshape(tiles, -1).view(torch.int32)                    # [tiles, mask images * 4096]
    weights = (1 << torch.arange(32, device=acts.device, dtype=torch.int64))
    for s in range(slots):
        nz = decode(acts, s, 1).view(tiles, 128, 2, 32) != 0                         # [tile, row, group, bit]
        w = (nz.to(torch.int64) * weights).sum(-1)                                   # [tile, row, group]
        w = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)
        base = (s // 16) * 4096 + (s % 16) * 256
        for g in range(2):
            words[:, base + g * 128: base + g * 128 + 128] = w[:, :, g]
    return out
