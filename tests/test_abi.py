"""CPU-only checks of the C-ABI boundary: the library builds, loads, and exports every symbol
that include/gen6d_b200.h declares; the ctypes binding covers all of them; the product path
fails loudly without a GPU (no fallback)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from gen6d_b200 import _lib


@pytest.fixture(scope='module')
def built():
    from gen6d_b200.build import build
    return build()


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    names = _lib.header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/gen6d_b200.h but not exported'


def test_binding_covers_header(built):
    bound = set(_lib._SIGNATURES) | {'g6d_last_error', 'g6d_version', 'g6d_launch_count'}
    assert set(_lib.header_symbols()) == bound


def test_error_reporting_without_gpu(built):
    l = _lib.lib()
    assert l.g6d_version() >= 100
    d = _lib.ConvDesc()  # all zeros -> invalid
    assert l.g6d_conv_workspace_bytes(ctypes.byref(d)) == -1
    assert b'bad dims' in l.g6d_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_compute_fails_loudly_without_cuda():
    from gen6d_b200.network import name2network
    det = name2network['detector']({})
    with pytest.raises(Exception):
        det.load_ref_imgs(np.zeros((2, 128, 128, 3), np.uint8))
