"""The C-ABI shared library: it loads, exports every symbol include/pacybara_b200.h declares, and refuses to
compute without a CUDA device (no CPU fallback exists).  No compute calls here: this runs on the GPU-less box."""
import ctypes as C
import os
import re

import pytest

from pacybara_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pacybara_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcb_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"_pcb.so does not export {missing}"
    assert sorted(_lib.SYMBOLS) == names, "pacybara_b200/_lib.py binds a different symbol set than the header declares"


def test_abi_version_and_constants(lib):
    assert lib.pcb_abi_version() == 2
    text = open(HEADER).read()
    for name, value in (("PCB_STAT_PAIRS_SCORED", _lib.STAT["pairs_scored"]), ("PCB_STAT_T_REPLAY_NS", _lib.STAT["t_replay_ns"]),
                        ("PCB_STAT_PROBE_NS", _lib.STAT["probe_ns"])):
        assert re.search(rf"{name} = {value}\b", text), name
    assert "#define PCB_NEEDS_ALIGNMENT (-2)" in text and _lib.NEEDS_ALIGNMENT == -2
    assert "#define PCB_FILL INT32_MIN" in text and _lib.FILL == -(1 << 31)


def test_no_device_means_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.pcb_ctx_create(0, C.byref(h))
    assert rc == _lib.PCB_E_CUDA and not h.value
    assert b"no CPU implementation" in lib.pcb_last_error(None)
    from pacybara_b200.api import Context
    with pytest.raises(_lib.PcbError):
        Context(0)


def test_product_code_never_touches_the_oracle():
    """Nothing under pacybara_b200/ or bin/ may import, link or execute anything under oracle/."""
    offenders = []
    for base in ("pacybara_b200", "bin"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".R")) or "." not in f:
                    path = os.path.join(dirpath, f)
                    try:
                        text = open(path, errors="ignore").read()
                    except OSError:
                        continue
                    if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "libpcb_oracle" in text or "pcb_oracle.c" in text:
                        offenders.append(os.path.relpath(path, ROOT))
    assert not offenders, offenders
