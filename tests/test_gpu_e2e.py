"""End-to-end parity of the product pipeline (`kamino-b200`: C++ host + CUDA front end) against the
reference's goldens and against the oracle CLI on synthetic proteome sets: alignment, missing-data
table, partitions and NJ tree must be byte-identical."""
import subprocess
from pathlib import Path

import pytest

from kamino_b200 import synth

pytestmark = pytest.mark.gpu
FILES = ["kamino_alignment.fas", "kamino_missing.tsv", "kamino_partitions.tsv", "kamino_NJ.tree"]


@pytest.fixture(scope="module")
def product_cli():
    from kamino_b200 import build
    return build.build_host()


def _run(cli, args, host_fasta=False):
    """host_fasta: KAMINO_HOST_FASTA=1 — FASTA records split by the host threads instead of kmn_count_fasta."""
    import os
    env = dict(os.environ)
    env.pop("KAMINO_HOST_FASTA", None)
    if host_fasta:
        env["KAMINO_HOST_FASTA"] = "1"
    r = subprocess.run([str(cli), *map(str, args)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    return r


@pytest.mark.parametrize("mode,scheme", [("dir", "dayhoff6"), ("tsv", "kgb6")])
def test_golden_3diff_through_gpu(product_cli, golden_dir, tmp_path, mode, scheme):
    """tests/golden.rs:96-124 with both passes and the --nj counts on the GPU."""
    src = golden_dir / "test_3diff"
    args = ["-i", src / "input"] if mode == "dir" else ["-I", src / "input.tsv"]
    _run(product_cli, [*args, "--length-middle", 35, "--recode", scheme, "--nj", "-t", 2, "-o", tmp_path / "kamino"])
    for f in FILES:
        assert (tmp_path / f).read_text() == (src / "expected" / f).read_text(), f


def test_golden_genomes_through_gpu(product_cli, golden_dir, tmp_path):
    """tests/golden.rs:126-139: --genomes (host ORF prediction into a temporary directory) in front of the GPU
    pipeline; the temporary proteomes are gone afterwards (lib.rs:296)."""
    src = golden_dir / "test_genomes"
    _run(product_cli, ["-i", src / "input", "--genomes", "--length-middle", 50, "--recode", "sr6", "-t", 2,
                       "-o", tmp_path / "kamino"])
    for f in FILES[:3]:
        assert (tmp_path / f).read_text() == (src / "expected" / f).read_text(), f
    assert not list(tmp_path.glob("tmp_kamino_*"))


@pytest.fixture(scope="module")
def synthetic_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("synth")
    species = synth.species_set(24, 500, 300.0, seed=21)
    synth.write_fasta_dir(species, d / "in")
    return d


@pytest.mark.parametrize("extra", [
    [], ["-r", "dayhoff6", "-k", 11], ["-r", "kgb6", "-f", 0.7, "-l", 50], ["-k", 8, "-c", 2, "-m", 3],
    ["-k", 21, "-f", 0.6], ["--batch-bytes", 700000],
])
def test_synthetic_matches_oracle_cli(product_cli, oracle, synthetic_dir, tmp_path, extra):
    want, got = tmp_path / "want", tmp_path / "got"
    want.mkdir(); got.mkdir()
    common = ["-i", synthetic_dir / "in", "--nj", "-t", 4, *extra]
    ref_args = [a for i, a in enumerate(common)            # --batch-bytes is a product-only tuning flag
                if a != "--batch-bytes" and (i == 0 or common[i - 1] != "--batch-bytes")]
    r = oracle.run_cli([*ref_args, "-o", want / "kamino"])
    assert r.returncode == 0, r.stderr
    for host_fasta in (False, True):   # records found on the device (default) / by the host threads
        for f in FILES:
            (got / f).unlink(missing_ok=True)
        _run(product_cli, [*common, "-o", got / "kamino"], host_fasta=host_fasta)
        for f in FILES:
            assert (got / f).read_bytes() == (want / f).read_bytes(), (f, extra, host_fasta)
    aln = (got / "kamino_alignment.fas").read_text().splitlines()
    assert len(aln) >= 48 and len(aln[1]) > 0, "expected a non-empty alignment on related proteomes"


@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_multi_device_cli_matches_oracle_cli(product_cli, oracle, synthetic_dir, tmp_path, n_dev):
    """`--devices 0,1,...`: the species shard round-robin over the GPUs of ONE process (the reference's rayon
    loops over species, group_extraction.rs:558-567 / :589-608), one in-process table exchange, pass 2 and the
    --nj row blocks on every device: byte-identical to the oracle CLI and to the one-GPU run."""
    import torch
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    want, got, one = tmp_path / "want", tmp_path / "got", tmp_path / "one"
    for d in (want, got, one):
        d.mkdir()
    for extra in ([], ["-r", "dayhoff6", "-k", 11, "--batch-bytes", 700000]):
        common = ["-i", synthetic_dir / "in", "--nj", "-t", 4, *extra]
        ref_args = [a for i, a in enumerate(common) if a != "--batch-bytes" and (i == 0 or common[i - 1] != "--batch-bytes")]
        r = oracle.run_cli([*ref_args, "-o", want / "kamino"])
        assert r.returncode == 0, r.stderr
        _run(product_cli, [*common, "-o", one / "kamino"])
        _run(product_cli, [*common, "--devices", ",".join(map(str, range(n_dev))), "-o", got / "kamino"])
        for f in FILES:
            assert (got / f).read_bytes() == (want / f).read_bytes(), (f, n_dev, extra)
            assert (got / f).read_bytes() == (one / f).read_bytes(), (f, n_dev, extra)


def test_awkward_fasta_files_match_oracle_cli(product_cli, oracle, synthetic_dir, tmp_path):
    """CRLF files, leading empty lines, a missing final line feed, '>' inside headers and a file whose first
    line is not a header: same outputs / same error text as the oracle CLI, with the records found on the device."""
    src = sorted((synthetic_dir / "in").glob("*.faa"))[:10]
    d = tmp_path / "in"
    d.mkdir()
    for i, f in enumerate(src):
        data = f.read_bytes()
        if i == 1:
            data = data.replace(b"\n", b"\r\n")
        elif i == 2:
            data = b"\n\r\n" + data
        elif i == 3:
            data = data.rstrip(b"\n")
        elif i == 4:
            data = data.replace(b" hypothetical", b" >hypo>thetical")
        elif i == 5:
            data = data.replace(b"\n", b"\n\n", 50)
        elif i == 7:   # Unicode white space around / inside the header fields (str::trim, split_whitespace)
            data = data.replace(b" hypothetical protein ", "\u00a0 hypothetical\u2003protein\u3000 ".encode())
        elif i == 8:
            data = data.replace(b" [Synth sp]", "\u2028 [Synth sp]\u0085".encode())
        (d / f.name).write_bytes(data)
    want, got = tmp_path / "want", tmp_path / "got"
    want.mkdir(); got.mkdir()
    common = ["-i", d, "--nj", "-t", 4, "-f", 0.7]
    r = oracle.run_cli([*common, "-o", want / "kamino"])
    assert r.returncode == 0, r.stderr
    _run(product_cli, [*common, "-o", got / "kamino"])
    for f in FILES:
        assert (got / f).read_bytes() == (want / f).read_bytes(), f
    # a file that does not start with '>': both stop with the reader's message
    (d / src[6].name).write_bytes(b"MKV\n" + src[6].read_bytes())
    r_ref = oracle.run_cli([*common, "-o", want / "bad"])
    r_got = subprocess.run([str(product_cli), *map(str, [*common, "-o", got / "bad"])], capture_output=True, text=True)
    assert r_ref.returncode != 0 and r_got.returncode != 0
    assert "expected '>' at record start" in r_ref.stderr and "expected '>' at record start" in r_got.stderr
    assert src[6].name in r_got.stderr


def test_cli_errors(product_cli, golden_dir, tmp_path):
    src = golden_dir / "test_3diff" / "input"
    for bad in (["-k", 22], ["-f", 0.5], ["-t", 0], ["-c", 14], ["--genomes"], ["--frobnicate"]):
        r = subprocess.run([str(product_cli), "-i", str(src), "-o", str(tmp_path / "x"), *map(str, bad)],
                           capture_output=True, text=True)
        assert r.returncode != 0, bad
