"""The C-ABI libraries load and export every symbol include/gbenv.h declares; no compute without a GPU."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "gbenv.h").read_text()
    return sorted(set(re.findall(r"\b(gbenv_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("gbenv_create", "gbenv_destroy", "gbenv_reset", "gbenv_step", "gbenv_step_host", "gbenv_add_state_template", "gbenv_load_template",
                 "gbenv_save_state", "gbenv_run_action", "gbenv_read_mem", "gbenv_write_mem", "gbenv_get_info", "gbenv_reduce_info"):
        assert must in names


def test_cuda_library_exports_every_declared_symbol(built):
    dll = ctypes.CDLL(str(built["cuda"]))
    for name in _declared():
        assert hasattr(dll, name), f"libgbenv.so does not export {name}"


def test_oracle_library_exports_the_same_surface(built):
    dll = ctypes.CDLL(str(built["oracle"]))
    for name in _declared():
        assert hasattr(dll, name.replace("gbenv_", "oracle_", 1)), name


def test_python_binding_covers_the_header():
    from pokegym_b200 import _capi

    bound = {"gbenv_" + n for n in _capi.EXPORTED_SYMBOLS}
    assert bound == set(_declared())


def test_create_fails_loudly_without_a_gpu(built, roms):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pokegym_b200 import _capi

    lib = _capi.GbEnvLib(built["cuda"])
    with pytest.raises(_capi.GbEnvError, match="no CUDA device|CUDA"):
        _capi.Handle(lib, 4, roms("pokelike"))


def test_missing_library_is_an_error(tmp_path):
    from pokegym_b200 import _capi

    with pytest.raises(_capi.GbEnvError, match="no CPU fallback"):
        _capi.GbEnvLib(tmp_path / "libgbenv.so")


def test_info_names_module_matches_the_header():
    """pokegym_b200/_info_names.py is generated from include/gbenv_info.h so that an installed package needs no header."""
    from pokegym_b200 import info

    assert info.INFO_NAMES == info.names_from_header()
