"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol include/pinn_b200.h
declares, the ctypes prototypes cover the header one to one, and compute entry points fail LOUDLY without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pinn_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pinn_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ["pinn_create", "pinn_destroy", "pinn_set_params", "pinn_get_params", "pinn_set_data", "pinn_set_collocation",
                 "pinn_sample_collocation", "pinn_loss_grad", "pinn_loss_grad_device", "pinn_adam_steps", "pinn_adam_apply",
                 "pinn_predict", "pinn_admm_update", "pinn_admm_init", "pinn_last_error", "pinn_packed_ptr"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from pinns_b200 import _capi
    lib = C.CDLL(_capi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "libpinn_b200.so does not export %s" % name


def test_ctypes_prototypes_match_header():
    from pinns_b200 import _capi
    assert sorted(_capi.PROTOTYPES) == declared_symbols()
    assert C.sizeof(_capi.PinnConfig) == 4 * 2 + 4 * 16 + 4 * 2 + 8 * 4 + 4 * 3 + 4 * 3 + 4 * 8  # pinn_config_t layout


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pinns_b200 import Engine
    from pinns_b200._capi import PinnError
    with pytest.raises(PinnError) as e:
        Engine([2, 20, 20, 1], [-1, 0], [1, 1])
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_invalid_configurations_are_rejected_before_touching_the_gpu():
    from pinns_b200 import _capi
    cfg = _capi.PinnConfig()
    cfg.abi_version = 999
    h = C.c_void_p()
    assert _capi.lib.pinn_create(C.byref(cfg), C.byref(h)) == -1
    assert b"abi_version" in _capi.lib.pinn_last_error(None)
    cfg.abi_version = _capi.PINN_B200_ABI_VERSION
    cfg.n_layers = 3
    cfg.layers[0], cfg.layers[1], cfg.layers[2] = 3, 20, 1   # layers[0] must be 2 (x,t)
    assert _capi.lib.pinn_create(C.byref(cfg), C.byref(h)) == -1
    cfg.layers[0] = 2
    cfg.pde, cfg.loss = _capi.PDE_EULER, _capi.LOSS_V4_MSE   # Euler needs 3 outputs
    assert _capi.lib.pinn_create(C.byref(cfg), C.byref(h)) == -1
    assert b"Euler" in _capi.lib.pinn_last_error(None)


def test_product_never_imports_the_oracle():
    """a product path that routes through oracle/ voids every parity claim"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pinns_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
