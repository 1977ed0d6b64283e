"""Small host-side helpers of the global aligner (mirrors reference dust3r/cloud_opt_flow/commons.py)."""
from __future__ import annotations

import numpy as np
import torch


def edge_str(i, j):
    """Key of a directed pair, reference commons.py:12-13."""
    return f'{i}_{j}'


def i_j_ij(ij):
    return edge_str(*ij), ij


def get_conf_trf(mode):
    """Confidence -> pixel weight maps, reference commons.py:53-64."""
    if mode == 'log':
        return lambda x: x.log()
    if mode == 'sqrt':
        return lambda x: x.sqrt()
    if mode == 'm1':
        return lambda x: x - 1
    if mode in ('id', 'none'):
        return lambda x: x
    raise ValueError(f'bad mode for {mode=}')


def get_conf_untrf(mode):
    """Inverse of get_conf_trf (the packed planes store the transformed weight)."""
    if mode == 'log':
        return lambda x: x.exp()
    if mode == 'sqrt':
        return lambda x: x.square()
    if mode == 'm1':
        return lambda x: x + 1
    if mode in ('id', 'none'):
        return lambda x: x
    raise ValueError(f'bad mode for {mode=}')


def signed_log1p(x):
    """reference commons.py:78-80."""
    return torch.sign(x) * torch.log1p(torch.abs(x))


def signed_expm1(x):
    """reference commons.py:83-85."""
    return torch.sign(x) * torch.expm1(torch.abs(x))


def cosine_schedule(t, lr_start, lr_end):
    """reference commons.py:88-90."""
    assert 0 <= t <= 1
    return lr_end + (lr_start - lr_end) * (1 + np.cos(t * np.pi)) / 2


def linear_schedule(t, lr_start, lr_end):
    """reference commons.py:93-95."""
    assert 0 <= t <= 1
    return lr_start + (lr_end - lr_start) * t


def cycled_linear_schedule(t, lr_start, lr_end, num_cycles=2):
    """reference commons.py:97-103."""
    assert 0 <= t <= 1
    cycle_t = t * num_cycles
    cycle_t = cycle_t - int(cycle_t)
    if t == 1:
        cycle_t = 1
    return linear_schedule(cycle_t, lr_start, lr_end)


def schedule_lr(schedule, cur_iter, niter, lr_base, lr_min):
    """Learning rate of one iteration, reference base_opt.py:551-563."""
    t = cur_iter / niter
    if schedule == 'cosine':
        return cosine_schedule(t, lr_base, lr_min)
    if schedule == 'linear':
        return linear_schedule(t, lr_base, lr_min)
    if schedule.startswith('cycle'):
        try:
            num_cycles = int(schedule[5:])
        except ValueError:
            num_cycles = 2
        return cycled_linear_schedule(t, lr_base, lr_min, num_cycles=num_cycles)
    raise ValueError(f'bad lr {schedule=}')


def unitquat_to_rotmat(q):
    """XYZW unit quaternion -> rotation matrix (roma convention used by reference base_opt.py:206)."""
    x, y, z, w = q.unbind(-1)
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, dim=-1).reshape(q.shape[:-1] + (3, 3))


def rotmat_to_unitquat(R):
    """Rotation matrix -> XYZW unit quaternion (roma.rotmat_to_unitquat semantics, base_opt.py:221)."""
    from scipy.spatial.transform import Rotation
    q = Rotation.from_matrix(R.detach().cpu().double().numpy()).as_quat()
    return torch.as_tensor(q, dtype=R.dtype, device=R.device)
