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
    r = _run(cli, "-i", str(GOLDEN / "input"), "--genomes", "-o", str(tmp_path / "x"))   # proteins are not genomes
    assert r.returncode != 0 and "does not appear to contain DNA sequences" in r.stderr      # protein_prediction.rs:301-306
    assert not list(tmp_path.glob("tmp_kamino_*")), "the temporary proteome directory must not outlive the run"
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


# ---- --genomes: the ORF predictor in front of the pipeline (protein_prediction.rs) is host logic, testable here

GENOMES = ROOT / "tests" / "golden" / "test_genomes"


def test_reference_unit_tests_on_the_host_logic(cli):
    """The reference's own unit tests restated on the product's host functions (kamino-b200 --self-test, no GPU):
    group_filtering.rs:414-693 (consensus, row masking and every filter_groups case), group_sorting.rs:486-638 (sort / non-overlap / length
    selection), group_extraction.rs:666-677 (protein names), protein_prediction.rs:1152-1286; plus the record
    bookkeeping of the device-framed FASTA path against the seq_io-style host parser."""
    r = _run(cli, "--self-test")
    assert r.returncode == 0 and "self-test ok" in r.stdout, r.stderr


def test_golden_genomes_predictor_then_oracle(cli, tmp_path):
    """tests/golden.rs:126-139 (golden_genomes): the product's predictor writes the proteomes, the oracle runs
    the rest of the pipeline on them on the CPU, and the three outputs equal the reference's expected files."""
    import sys
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_ffi
    pred = tmp_path / "pred"
    r = _run(cli, "-i", str(GENOMES / "input"), "--genomes", "--predict-only", str(pred), "-t", "2")
    assert r.returncode == 0, r.stderr
    assert sorted(p.name for p in pred.iterdir()) == ["genome_A.faa", "genome_B.faa"]
    a = (pred / "genome_A.faa").read_text().splitlines()
    assert a[0] == ">protein_1" and all(len(line) <= 60 for line in a[1:])
    prot = "".join(a[1:])
    assert prot.startswith("M") and len(prot) >= 90 and "*" not in prot
    out = tmp_path / "out"
    out.mkdir()
    ro = oracle_ffi.load().run_cli(["-i", pred, "--length-middle", 50, "--recode", "sr6", "-t", 2, "-o", out / "kamino"])
    assert ro.returncode == 0, ro.stderr
    for f in ("kamino_alignment.fas", "kamino_missing.tsv", "kamino_partitions.tsv"):
        assert (out / f).read_bytes() == (GENOMES / "expected" / f).read_bytes(), f


def test_predictor_on_a_synthetic_genome(cli, tmp_path):
    """A 300 kb random genome with planted genes on both strands (enough long ORFs to train the adaptive codon
    model): planted genes are recovered, every protein obeys the emission rules, and the output does not
    depend on the thread count or on line wrapping / case of the input."""
    import numpy as np
    rng = np.random.default_rng(5)
    codons = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT" if a + b + c not in ("TAA", "TAG", "TGA")]
    comp = str.maketrans("ACGT", "TGCA")
    parts, genes = [], []
    for g in range(160):
        parts.append("".join(rng.choice(list("ACGT"), int(rng.integers(60, 400)))))
        body = "".join(rng.choice(codons, int(rng.integers(130, 400))))
        gene = "ATG" + body + "TAA"
        genes.append(gene)
        parts.append(gene if g % 2 == 0 else gene.translate(comp)[::-1])
    genome = "".join(parts)
    d1, d2 = tmp_path / "in1", tmp_path / "in2"
    d1.mkdir(); d2.mkdir()
    (d1 / "g.fna").write_text(">contig1 test\n" + "\n".join(genome[i:i + 70] for i in range(0, len(genome), 70)) + "\n")
    (d2 / "g.fna").write_text(">contig1\n" + genome.lower() + "\n")
    outs = []
    for d, t in ((d1, "1"), (d1, "4"), (d2, "2")):
        o = tmp_path / f"pred_{d.name}_{t}"
        r = _run(cli, "-i", str(d), "--genomes", "--predict-only", str(o), "-t", t)
        assert r.returncode == 0, r.stderr
        outs.append((o / "g.faa").read_text())
    assert outs[0] == outs[1] == outs[2]
    recs = outs[0].split(">")[1:]
    prots = ["".join(r.splitlines()[1:]) for r in recs]
    assert [r.splitlines()[0] for r in recs] == [f"protein_{i + 1}" for i in range(len(recs))]
    assert len(prots) >= 150
    assert all(p.startswith("M") and len(p) >= 90 and "*" not in p for p in prots)
    table = dict(zip([a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"],
                     "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"))
    found = 0
    for gene in genes:
        want = "".join(table[gene[i:i + 3]] for i in range(0, len(gene) - 3, 3))
        found += any(p.endswith(want[-80:]) for p in prots)     # the start may be an upstream alternative
    assert found >= 150, found


def _synthetic_genome(seed, n_genes, with_n=False):
    import numpy as np
    rng = np.random.default_rng(seed)
    codons = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT" if a + b + c not in ("TAA", "TAG", "TGA")]
    comp = str.maketrans("ACGT", "TGCA")
    parts = []
    for g in range(n_genes):
        parts.append("".join(rng.choice(list("ACGT"), int(rng.integers(20, 400)))))
        if g % 9 == 3:
            parts.append("AAGGAGG" + "".join(rng.choice(list("ACGT"), int(rng.integers(3, 12)))))   # Shine-Dalgarno-like
        start = ["ATG", "GTG", "TTG"][int(rng.integers(0, 3))]
        gene = start + "".join(rng.choice(codons, int(rng.integers(60, 420)))) + ["TAA", "TAG", "TGA"][g % 3]
        if with_n and g % 11 == 5:
            gene = gene[:150] + "N" * int(rng.integers(1, 5)) + gene[150:]
        parts.append(gene if g % 2 == 0 else gene.translate(comp)[::-1])
    return "".join(parts)


def test_predictor_matches_the_python_restatement(cli, tmp_path):
    """Differential test: the C++ host predictor against oracle/protein_prediction.py (written independently from
    the Rust source) on synthetic genomes — several contigs, N runs, lower case, CRLF, a contig shorter than a
    codon, enough long ORFs on one contig to train the adaptive codon model, and the golden genomes."""
    import sys
    sys.path.insert(0, str(ROOT))
    from oracle import protein_prediction as orc
    d = tmp_path / "in"
    d.mkdir()
    big = _synthetic_genome(11, 140)                     # adaptive model kicks in
    small = _synthetic_genome(12, 25, with_n=True)
    (d / "big.fna").write_text(">c1\n" + "\n".join(big[i:i + 80] for i in range(0, len(big), 80)) + "\n")
    (d / "multi.fa").write_text(">c1 first\r\n" + small.lower() + "\r\n>c2\r\nAC\r\n>c3\r\n" + _synthetic_genome(13, 12) + "\r\n")
    import gzip
    with gzip.open(d / "zipped.fas.GZ", "wb") as f:      # the predictor gunzips "gz" in any case (:361-365)
        f.write((">x\n" + _synthetic_genome(14, 30, with_n=True) + "\n").encode())
    for f in (GENOMES / "input").iterdir():
        (d / f.name).write_bytes(f.read_bytes())
    out = tmp_path / "pred"
    r = _run(cli, "-i", str(d), "--genomes", "--predict-only", str(out), "-t", "3")
    assert r.returncode == 0, r.stderr
    n_checked = 0
    for f in sorted(d.iterdir()):
        name = {"zipped.fas.GZ": "zipped.fas.GZ"}.get(f.name, f.name.split(".")[0])   # io.rs:164-183 strips only a lower-case ".gz"
        got = (out / f"{name}.faa").read_text()
        want = orc.predict_one_genome(f)
        assert got == want, f.name
        n_checked += got.count(">")
    assert n_checked > 150
    # the big contig does reach the adaptive codon model (otherwise that path would be compared vacuously)
    rc = "".join(orc.COMPLEMENT.get(ch, "N") for ch in reversed(big))
    gm = orc.gc_model(orc.f32(sum(ch in "GC" for ch in big) / len(big)))
    orfs = []
    orc.find_orfs_on_strand(big, 1, 270, gm, orfs)
    orc.find_orfs_on_strand(rc, -1, 270, gm, orfs)
    assert orc.train_adaptive(orfs, big, rc) is not None
