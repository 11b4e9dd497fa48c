"""The boundary is a C ABI: a C99 program (tests/c/abi_smoke.c) compiles against include/gm2.h with
-Wall -Wextra -Werror -pedantic, links libgm2.so and exercises the host-only entry points plus the
"no GPU -> loud failure" rule.  No compute calls here.  CPU only."""
from __future__ import annotations

import os
import shutil
import subprocess

import pytest

from genome_minimizer_2_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c99_caller_compiles_links_and_runs(tmp_path):
    lib = build.build_native()
    exe = tmp_path / "abi_smoke"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                           "-o", str(exe), lib, f"-Wl,-rpath,{os.path.dirname(lib)}"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, f"check {r.returncode} failed\n{r.stdout}{r.stderr}"


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_self_contained_c(tmp_path):
    """include/gm2.h compiles on its own as C (no hidden dependency on C++ or on another header)."""
    src = tmp_path / "only_header.c"
    src.write_text('#include "gm2.h"\nint main(void) { return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only",
                           "-I", os.path.join(ROOT, "include"), str(src)])
