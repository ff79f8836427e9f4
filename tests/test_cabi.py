"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header
declares, and refuses to run without a B200 (no CPU fallback)."""
import ctypes
import importlib
import os
import subprocess

import pytest
import torch


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load()
    names = pkg.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_binding_covers_header(pkg):
    lib_mod = importlib.import_module("menghini-neurips23-code_b200._lib")
    assert set(lib_mod._SIGNATURES) == set(pkg.declared_symbols())


def test_version_and_sizes_without_gpu(pkg):
    lib = pkg.load()
    assert b"sm_100a" in lib.gb_version()
    # pure host arithmetic entry points
    assert lib.gb_tape_bytes(2, 50, 768, 12) == (12 * 9 + 1) * 2 * 50 * 768 * 2
    assert lib.gb_tape_bytes(0, 50, 768, 12) == 0
    assert lib.gb_leaderboard_state_bytes(10, 16) == (8 + 30 + 2 * 160 + 2 * 17) * 4
    assert lib.gb_leaderboard_state_bytes(0, 16) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(pkg):
    lib = pkg.load()
    h = ctypes.c_void_p()
    assert lib.gb_create(ctypes.byref(h), 0) == -4  # GB_ERR_NO_DEVICE
    with pytest.raises(pkg.GripB200Error):
        pkg.Context(0)
    clip = importlib.import_module("menghini-neurips23-code_b200.clip")
    with pytest.raises(pkg.GripB200Error):
        clip.load("ViT-B/32", "cpu", state_dict={})
    eng = importlib.import_module("menghini-neurips23-code_b200.engine")
    with pytest.raises(pkg.GripB200Error):
        eng.Leaderboard(4, 2, "cpu")


def test_product_does_not_import_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "menghini-neurips23-code_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), os.path.join(dp, f)
