"""Runs oracle/_ref/dropin_test: the reference Opm::EulerUpstream and the drop-in Opm::b200::EulerUpstream
(opm-porsol_b200/host) side by side in one C++ program, same GridInterface / ReservoirProperties /
BoundaryConditions objects (tests/cpp/dropin_test.cpp).  The binary is built where /root/reference exists."""
import os
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_test")


@pytest.mark.gpu
def test_dropin_header_matches_reference(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_test not built (needs /root/reference)")
    out = subprocess.run([BIN, str(tmp_path)], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    print(out.stderr)
    assert out.returncode == 0 and "DROPIN TEST PASSED" in out.stdout, out.stdout + out.stderr


def test_dropin_header_fails_loudly_without_a_gpu(tmp_path):
    """No CPU fallback behind the C++ drop-in either: without a CUDA device initObj throws
    'EulerUpstream (B200): no CUDA device (there is no CPU fallback)'.  Runs where there is no GPU (the build container)."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_test not built (needs /root/reference)")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: covered by test_dropin_header_matches_reference")
    except ImportError:
        pass
    out = subprocess.run([BIN, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0
    assert "no CUDA device (there is no CPU fallback)" in out.stderr and "DROPIN TEST PASSED" not in out.stdout


FLATTEN = os.path.join(ROOT, "oracle", "_ref", "flatten_test")


def test_dropin_headers_host_logic_against_recording_stub(tmp_path):
    """tests/cpp/flatten_test.cpp: the drop-in headers linked against a recording stub of the C ABI (no GPU): the
    flattening walk, fluid extraction, parameter keys, flux gathering, exception mapping, the residual mirror and the
    diagnostics wrappers, against what the reference's own grid / property / boundary-condition objects say."""
    if not os.path.exists(FLATTEN):
        pytest.skip("oracle/_ref/flatten_test not built (needs /root/reference)")
    out = subprocess.run([FLATTEN, str(tmp_path)], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0 and "FLATTEN TEST PASSED" in out.stdout, out.stdout + out.stderr
