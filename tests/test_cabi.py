"""The C-ABI library loads and exports every symbol include/ngan_b200.h declares (no compute calls: CPU)."""
import ctypes
import os

from neuron_gan_b200 import _lib


def test_header_parses_and_all_symbols_are_exported():
    protos = _lib.parse_header()
    assert len(protos) >= 39 and 'ngan_conv3x3_fwd' in protos and 'ngan_adam_multi' in protos
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), f'{name} declared in include/ngan_b200.h but not exported'
    assert lib.ngan_version() == 100


def test_argument_validation_without_gpu():
    lib = _lib.load()
    rc = lib.ngan_nchw_to_c8(None, None, 1, 12, 4, 4, None)      # C not a multiple of 8, null pointers
    assert rc == -1
    assert b'nchw_to_c8' in lib.ngan_last_error()


def test_adam_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.AdamTensor) == 80
    assert _lib.AdamTensor.n.offset == 40 and _lib.AdamTensor.step_size.offset == 48 and _lib.AdamTensor.dyn.offset == 72
