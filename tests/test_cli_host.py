"""CPU tests of the C++ host pipeline `kamino-b200` (kamino's CLI above the C ABI): flag validation mirrors
/root/reference/src/lib.rs:206-227, and without a GPU the run fails loudly instead of taking a host path."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden" / "test_3diff"


@pytest.fixture(scope="module")
def cli():
    from kamino_b200 import build
    return str(build.build_host())


def _run(cli, *args):
    return subprocess.run([cli, *args], capture_output=True, text=True, timeout=120)


def test_flag_validation(cli, tmp_path):
    r = _run(cli)
    assert r.returncode != 0 and "input" in r.stderr.lower()
    r = _run(cli, "-i", str(GOLDEN / "input"), "-k", "22", "-o", str(tmp_path / "x"))      # lib.rs:218
    assert r.returncode != 0 and "k" in r.stderr
    r = _run(cli, "-i", str(GOLDEN / "input"), "-f", "0.5", "-o", str(tmp_path / "x"))     # lib.rs:213-216
    assert r.returncode != 0
    r = _run(cli, "-i", str(GOLDEN / "input"), "-k", "5", "-c", "6", "-o", str(tmp_path / "x"))   # lib.rs:219
    assert r.returncode != 0
    r = _run(cli, "-i", str(GOLDEN / "input"), "--genomes", "-o", str(tmp_path / "x"))
    assert r.returncode != 0 and "genomes" in r.stderr
    r = _run(cli, "-i", str(tmp_path / "does_not_exist"), "-o", str(tmp_path / "x"))
    assert r.returncode != 0


def test_no_gpu_means_error_not_fallback(cli, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(cli, "-i", str(GOLDEN / "input"), "--recode", "dayhoff6", "--length-middle", "35", "--nj",
             "-o", str(tmp_path / "kamino"))
    assert r.returncode != 0
    assert "CUDA" in r.stderr or "GPU" in r.stderr
    assert not list(tmp_path.glob("kamino_*"))
