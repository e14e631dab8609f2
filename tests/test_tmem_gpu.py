"""The tensor-memory scratchpad of the sweep kernels (csrc/tmem.cuh): device self-test through the C ABI."""
import ctypes as C

import pytest

from idefix_b200 import _abi

pytestmark = pytest.mark.gpu


def test_tmem_scratchpad_roundtrip():
    bad = C.c_longlong(-1)
    assert _abi.lib().idfx_selftest_tmem(0, C.byref(bad)) == 0
    assert bad.value == 0
