"""TEST INFRASTRUCTURE - CPU oracle for the front / back of the decode (rows f3 of SURVEY.md section 8).

Plain-PyTorch restatement of reference matcha/inference.py:146-172 and matcha/utils/model.py:7-68; only tests/ import it.
Pinned by tests/golden/front_back.npz, which tests/golden/make_golden.py generates with the reference's OWN functions
(matcha/utils/model.py imports nothing but torch, so it runs unmodified here).
"""
import torch
import torch.nn.functional as F


def sequence_mask(length, max_length):
    """ref: utils/model.py:7-9"""
    return torch.arange(max_length, dtype=length.dtype, device=length.device).unsqueeze(0) < length.unsqueeze(1)


def fix_len_compatibility(length: int, num_downsamplings_in_unet: int = 1) -> int:
    """ref: utils/model.py:15-21"""
    factor = 2 ** num_downsamplings_in_unet
    return -(-int(length) // factor) * factor


def generate_path(duration, mask):
    """ref: utils/model.py:24-40"""
    b, t_x, t_y = mask.shape
    cum = torch.cumsum(duration.long(), 1)
    path = sequence_mask(cum.view(b * t_x), t_y).to(mask.dtype).view(b, t_x, t_y)
    path = path - F.pad(path, [0, 0, 1, 0, 0, 0])[:, :-1]
    return path * mask


def front(mu_x, phoneme_durations, x_mask):
    """ref: inference.py:146-167.  phoneme_durations: (B, Tx) already rounded / clamped / masked (inference.py:143).
    Returns mu_y (B, F, T), y_mask (B, 1, T), y_lengths (B,), y_max_length."""
    y_fine_lengths = torch.clamp_min(phoneme_durations.sum(dim=1).long(), 1)
    y_fine_max_length_ = fix_len_compatibility(int(y_fine_lengths.max())) * 2
    y_fine_mask = sequence_mask(y_fine_lengths, y_fine_max_length_).unsqueeze(1).to(x_mask.dtype)
    attn_mask_fine = x_mask.unsqueeze(-1) * y_fine_mask.unsqueeze(2)
    attn_fine = generate_path(phoneme_durations, attn_mask_fine.squeeze(1)).unsqueeze(1)
    mu_y_fine = torch.matmul(mu_x.float(), attn_fine.float().squeeze(1))
    mu_y = F.avg_pool1d(mu_y_fine, kernel_size=3, stride=2, padding=1)  # utils/model.py:57-68
    y_max_length_ = y_fine_max_length_ // 2
    y_lengths = torch.clamp_min((y_fine_lengths + 1) // 2, 1)
    y_mask = sequence_mask(y_lengths, y_max_length_).unsqueeze(1).to(x_mask.dtype)
    return mu_y, y_mask, y_lengths, int(y_lengths.max())


def back(decoder_outputs, y_max_length, mel_mean, mel_std):
    """ref: inference.py:170-172, utils/model.py:52-54"""
    return decoder_outputs[:, :, :y_max_length] * mel_std + mel_mean
