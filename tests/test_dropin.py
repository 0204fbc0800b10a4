"""Drop-in check of the host mirror: the reference's own command-line front end
(cfemm/fsolver/main.cpp) must compile UNCHANGED against this repo's headers and link against
libxfemm_b200.so, and the resulting binary must keep the reference's exit-code contract.
Needs the reference sources, so it only runs where /root/reference exists (never on the GPU box)."""
import os
import subprocess

import pytest

import xfemm_b200 as X
from util import stage_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MAIN = "/root/reference/cfemm/fsolver/main.cpp"


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="reference sources not present")
def test_reference_main_compiles_and_links_against_our_headers(tmp_path):
    exe = str(tmp_path / "fsolver_ref_main")
    host = os.path.join(ROOT, "xfemm_b200", "host")
    libdir = os.path.join(ROOT, "xfemm_b200")
    # quote-includes search the including file's directory first, so compile a scratch copy placed
    # outside the reference tree (the copy lives in pytest's tmp dir only)
    src = str(tmp_path / "main.cpp")
    with open(REF_MAIN) as fi, open(src, "w") as fo:
        fo.write(fi.read())
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-I" + host, src, "-o", exe, "-L" + libdir, "-lxfemm_b200",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # exit code 1: the .fem cannot be loaded (main.cpp:72-77)
    r = subprocess.run([exe, str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.returncode == 1
    if X.device_count() == 0:
        # exit code 2: the solve fails — here because there is no GPU and no CPU fallback (main.cpp:79-80)
        base = stage_case("tb", tmp_path)
        r = subprocess.run([exe, base], capture_output=True, text=True)
        assert r.returncode == 2


def test_own_cli_exit_codes(tmp_path):
    exe = os.path.join(ROOT, "xfemm_b200", "fsolver_b200")
    assert os.path.exists(exe)
    r = subprocess.run([exe, str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.returncode == 1
    r = subprocess.run([exe, "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "Too many arguments" in r.stdout
