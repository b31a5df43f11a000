"""Run files of the reference: `<run>_params.pth` = the pickled argparse namespace train_dual.py:127
writes and test_dual.py:93 reads, `<run>_model.pth` = `net.state_dict()` (train_dual.py:276, test_dual.py:130).  Both are
read with torch's restricted unpickler (argparse.Namespace allow-listed) instead of a free pickle load.
"""
from __future__ import annotations

import argparse
import os

import torch


def save_params(opt, path):
    """train_dual.py:127."""
    torch.save(opt if isinstance(opt, argparse.Namespace) else argparse.Namespace(**dict(opt)), path)


def load_params(path):
    """test_dual.py:93 -> argparse.Namespace."""
    with torch.serialization.safe_globals([argparse.Namespace]):
        return torch.load(path, map_location="cpu", weights_only=True)


def save_model(net, path):
    """train_dual.py:276."""
    torch.save(net.state_dict(), path)


def load_model(net, path):
    """test_dual.py:130; PyG 1.x FeaStConv keys are converted by nn.FeaStConv._load_from_state_dict."""
    net.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))
    return net


def load_run(params_path, device="cuda", sub_size=None):
    """test_dual.py:93-96,127-132: (opt, net in eval mode on `device`) from a params file and the model file it names
    (looked up next to it)."""
    from . import network
    opt = load_params(params_path)
    opt.sub_size = opt.sub_size if sub_size is None else sub_size
    net = network.DualGNN(force_depth=opt.force_depth, pool_type=opt.pool_type, wei_param=opt.wei_param)
    load_model(net, os.path.join(os.path.dirname(params_path), opt.model_name))
    return opt, net.to(device).eval()
