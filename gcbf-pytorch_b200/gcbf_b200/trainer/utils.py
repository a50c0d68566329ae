"""Host-side helpers with the names the reference's scripts import from gcbf/trainer/utils.py."""
import os
import random
from typing import Optional

import numpy as np
import torch

# gcbf/trainer/hyperparams.yaml (gcbf rows)
_HYPERPARAMS = {
    'SimpleCar': dict(alpha=1.0, eps=0.02, inner_iter=10, loss_action_coef=0.05, loss_unsafe_coef=1.0,
                      loss_safe_coef=1.0, loss_h_dot_coef=0.5),
    'SimpleDrone': dict(alpha=1.0, eps=0.02, inner_iter=10, loss_action_coef=0.05, loss_unsafe_coef=1.0,
                        loss_safe_coef=1.0, loss_h_dot_coef=0.5),
    'DubinsCar': dict(alpha=1.0, eps=0.02, inner_iter=10, loss_action_coef=0.0001, loss_unsafe_coef=1.0,
                      loss_safe_coef=1.0, loss_h_dot_coef=0.2),
}


def set_seed(seed: int):
    """reference gcbf/trainer/utils.py:20-25"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


# gcbf/trainer/hyperparams.yaml (macbf rows): only the action and h_dot coefficients differ from the gcbf rows
_MACBF_COEFS = {'SimpleCar': (0.0001, 1.0), 'SimpleDrone': (0.01, 1.0), 'DubinsCar': (0.0005, 1.0)}


def read_params(env: str, algo: str) -> Optional[dict]:
    """reference gcbf/trainer/utils.py:317-340 (per-env hyper-parameter table)."""
    if algo not in ('gcbf', 'macbf') or env not in _HYPERPARAMS:
        return None
    hp = dict(_HYPERPARAMS[env])
    if algo == 'macbf':
        hp['loss_action_coef'], hp['loss_h_dot_coef'] = _MACBF_COEFS[env]
    return hp


def init_logger(log_path: str, env: str, algo: str, seed: int, args: dict = None, hyper_params: dict = None) -> str:
    """Creates <log_path>/<env>/<algo>/seed<seed>_<k>/ and writes settings.yaml (reference utils.py:28-105)."""
    import datetime
    import yaml
    base = os.path.join(log_path, env, algo)
    os.makedirs(base, exist_ok=True)
    stamp = datetime.datetime.now().strftime('%Y%m%d%H%M%S')
    run = os.path.join(base, f'seed{seed}_{stamp}')
    os.makedirs(run, exist_ok=True)
    with open(os.path.join(run, 'settings.yaml'), 'w') as f:
        yaml.safe_dump({**(args or {}), 'hyper_params': hyper_params or {}}, f)
    return run
