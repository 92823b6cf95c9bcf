"""The C-ABI library loads on a CPU-only host and exports every symbol include/b2sim.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2sim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2(?:sim|model)_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(engine_lib):
    import b2sim
    names = declared_symbols()
    assert len(names) > 40
    for name in names:
        assert hasattr(engine_lib, name), f"{name} is declared in include/b2sim.h but not exported"
    # the Python binding declares exactly the header's symbols
    assert sorted(b2sim._lib.SYMBOLS) == names


def test_no_cpu_fallback(engine_lib):
    """Without a GPU the engine refuses to create a simulator instead of falling back to the CPU."""
    if engine_lib.b2sim_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    h = engine_lib.b2sim_create(0, 4, 0.001, 1, 0)
    assert not h
    assert b"no CUDA device" in engine_lib.b2sim_last_error()


def test_product_does_not_touch_the_oracle():
    """Nothing under gym-ignition_b200/ may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "gym-ignition_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "b2oracle" not in text, f"{f} references the oracle library"
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
