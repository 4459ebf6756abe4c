"""The C-ABI library loads without a GPU, exports every symbol include/wxb200.h declares, and
fails loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "wxb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wxb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from whisperx import _native
    lib = _native.load_library()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wxb200.h but not exported"
    assert set(names) == set(_native.EXPORTED_SYMBOLS)
    assert lib.wxb_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from whisperx import _native
    lib = _native.load_library()
    h = ctypes.c_void_p()
    assert lib.wxb_create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.wxb_last_error(None)
    with pytest.raises(_native.WxbError):
        _native.Context(0)
    import numpy as np
    import whisperx.audio as wa
    with pytest.raises(RuntimeError):
        wa.log_mel_spectrogram(np.zeros(16000, np.float32), 80, device="cpu")
