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
    assert lib.pcb_abi_version() == 3
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


C_CALLER = r'''
#include <stdio.h>
#include <string.h>
#include "pacybara_b200.h"
/* a C (not C++) caller: the header must be plain C and the library must link without torch or the C++ runtime in sight */
int main(void) {
  pcb_ctx *ctx = NULL;
  int rc;
  if (pcb_abi_version() != PCB_ABI_VERSION) return 2;
  rc = pcb_ctx_create(0, &ctx);
  if (rc == PCB_OK) {                       /* a GPU is there: one tiny call through the boundary */
    const uint8_t seq[8] = {'A','C','G','T','A','C','G','T'};
    const int32_t len[2] = {4, 4};
    int32_t bor[2], ptr[3], reads[2], U = 0;
    rc = pcb_bucket(ctx, PCB_MEM_HOST, seq, NULL, len, 2, 4, bor, &U, ptr, reads, NULL);
    printf("gpu: rc=%d buckets=%d\n", rc, (int)U);
    pcb_ctx_destroy(ctx);
    return (rc == PCB_OK && U == 1) ? 0 : 3;
  }
  printf("no gpu: rc=%d (%s)\n", rc, pcb_last_error(NULL));
  return (rc == PCB_E_CUDA && strstr(pcb_last_error(NULL), "no CPU implementation")) ? 0 : 4;
}
'''


def _run_c_caller(tmp_path):
    import shutil
    import subprocess
    from pacybara_b200 import _lib as L
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if cc is None or not os.path.exists(L.SO):
        pytest.skip("no C compiler or library not built")
    src = tmp_path / "caller.c"
    src.write_text(C_CALLER)
    exe = str(tmp_path / "caller")
    so_dir = os.path.dirname(L.SO)
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                    L.SO, f"-Wl,-rpath,{so_dir}"], check=True, capture_output=True, text=True)
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)


def test_a_c_program_can_include_the_header_and_link_the_library(tmp_path):
    """C99, -Wall -Wextra -pedantic -Werror on the header; without a GPU the library must say so, not compute."""
    r = _run_c_caller(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_the_c_program_computes_on_the_gpu(tmp_path):
    """The same plain-C caller on the GPU box: its pcb_bucket call must go through (the driver's -m gpu run sees this)."""
    r = _run_c_caller(tmp_path)
    assert r.returncode == 0 and r.stdout.startswith("gpu: rc=0 buckets=1"), r.stdout + r.stderr
