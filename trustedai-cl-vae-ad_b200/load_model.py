"""Model factory + config IO: the drop-in seam (src/load_model.py:9-83)."""
import os
from copy import deepcopy

import yaml


def import_vae_based_on_type(vae_type: str):
    """src/load_model.py:9-31 - same type strings, same errors."""
    AVAILABLE_TYPES = ['KLGaussian', 'KurtosisGlobal', 'KurtosisSingle']
    from .model import KurtosisGlobalCVAE, KurtosisSingleCVAE
    if vae_type is None:
        return KurtosisGlobalCVAE
    if vae_type not in AVAILABLE_TYPES:
        raise Exception(f'Error, type {vae_type} not found in available types: {AVAILABLE_TYPES}')
    kind = vae_type.lower()
    if kind == 'klgaussian':
        raise NotImplementedError('KLGaussian not yet implemented')
    return KurtosisGlobalCVAE if kind == 'kurtosisglobal' else KurtosisSingleCVAE


def load_config(config_filename: str):
    assert os.path.exists(config_filename)
    assert os.path.isfile(config_filename)
    with open(config_filename, 'r') as ifile:
        return yaml.safe_load(ifile)


def save_config(config: dict, config_filename: str):
    with open(config_filename, 'w') as ofile:
        yaml.safe_dump(dict(config), ofile)


def load_model_from_config_path(config_path: str):
    assert os.path.exists(config_path)
    config = load_config(config_path)
    return load_model_from_config(config), config


def load_model_from_config(config: dict, **kwargs):
    # deep copy like the reference (:72) so the caller's dict is never mutated
    return import_vae_based_on_type(config['model'].get('type'))(deepcopy(config), **kwargs)


def load_model_from_directory(log_dir: str, **kwargs):
    assert os.path.exists(log_dir)
    assert os.path.isdir(log_dir)
    config = load_config(os.path.join(log_dir, 'config.yml'))
    model = load_model_from_config(config, **kwargs)
    model.load_model(log_dir)
    return model, config


def save_model_to_directory(model, log_dir: str, with_optimizer: bool = True):
    """What train.py:75-89,127-128 leaves behind: config.yml + encoder/ + decoder/ (+ Adam state)."""
    os.makedirs(log_dir, exist_ok=True)
    save_config(model.config, os.path.join(log_dir, 'config.yml'))
    model.encoder.save(os.path.join(log_dir, 'encoder'))
    model.decoder.save(os.path.join(log_dir, 'decoder'))
    if with_optimizer and model.optimizer is not None:
        model.save_optimizer(log_dir)
