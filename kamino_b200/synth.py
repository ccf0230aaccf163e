"""Deterministic synthetic proteome sets for the benchmark and the parity tests.

The reference ships no generator and there is no network, so the named configurations of
BASELINE.json ("100 synthetic bacterial proteomes (~4k proteins x ~330 aa)", ...) are produced here:
a root proteome with i.i.d. LG-frequency residues and log-normal protein lengths, evolved along a
random tree by point substitutions, short indels, gene loss and a sprinkle of 'X' residues.
Everything is seeded; the same (config, seed) always yields the same bytes.
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass
from pathlib import Path

import numpy as np

AA = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
# LG stationary frequencies in AA order (reference: src/phylo.rs:27-31)
LG_PI = np.array([
    0.079066, 0.055941, 0.041977, 0.053052, 0.012937, 0.040767, 0.071586, 0.057337, 0.022355, 0.062157,
    0.099081, 0.064600, 0.022951, 0.042302, 0.044040, 0.061197, 0.053287, 0.012777, 0.027843, 0.070200,
])
LG_PI = LG_PI / LG_PI.sum()
SEED0 = 0x6B616D696E6F  # "kamino"


@dataclass
class Species:
    seq: np.ndarray   # uint8 residues (letters), all proteins concatenated
    lens: np.ndarray  # int64 protein lengths


@dataclass
class StagedBatch:
    """Batch in the layout kmn_count_batch takes."""
    residues: np.ndarray    # uint8: concatenated rec.seq() bytes (with line terminators)
    rec_off: np.ndarray     # uint64 [n_rec+1]
    sp_rec_off: np.ndarray  # uint64 [n_sp+1]
    n_residues: int         # sequence residues, whitespace excluded (the metric's unit)

    @property
    def n_sp(self) -> int:
        return len(self.sp_rec_off) - 1


def _random_residues(rng: np.random.Generator, n: int) -> np.ndarray:
    return AA[rng.choice(20, size=n, p=LG_PI)]


def _lengths(rng, n_prot, mean, lo, hi, sigma=0.6):
    mu = np.log(mean) - 0.5 * sigma * sigma
    return np.clip(rng.lognormal(mu, sigma, n_prot).astype(np.int64), lo, hi)


def root_proteome(rng, n_prot: int, mean_len: float, lo=50, hi=3000) -> Species:
    lens = _lengths(rng, n_prot, mean_len, lo, hi)
    return Species(_random_residues(rng, int(lens.sum())), lens)


def evolve(rng, sp: Species, p_sub: float, indel_rate=1 / 2000, p_loss=0.02, x_rate=0.0005) -> Species:
    seq, lens = sp.seq.copy(), sp.lens.copy()
    n = len(seq)
    # substitutions
    m = rng.binomial(n, p_sub)
    if m:
        seq[rng.integers(0, n, m)] = _random_residues(rng, m)
    prot = np.repeat(np.arange(len(lens)), lens)
    # single-residue deletions and insertions (keep every protein at >= 30 residues)
    nd = rng.binomial(n, indel_rate / 2)
    keep = np.ones(n, dtype=bool)
    if nd:
        dpos = rng.integers(0, n, nd)
        dpos = dpos[lens[prot[dpos]] > 40]
        keep[dpos] = False
    ni = rng.binomial(n, indel_rate / 2)
    ins_after = np.zeros(n, dtype=np.int64)
    if ni:
        np.add.at(ins_after, rng.integers(0, n, ni), 1)
    # gene loss
    lost = rng.random(len(lens)) < p_loss
    keep &= ~lost[prot]
    ins_after[~keep] = 0
    reps = keep.astype(np.int64) + ins_after
    new_seq = np.repeat(seq, reps)
    # every copy after the first of a repeated residue is an insertion: give it a fresh residue
    if ni:
        starts = np.cumsum(reps) - reps
        in_group = np.arange(len(new_seq), dtype=np.int64) - np.repeat(starts, reps)
        fresh = in_group >= 1
        k = int(fresh.sum())
        if k:
            new_seq[fresh] = _random_residues(rng, k)
    new_lens = np.bincount(prot, weights=reps, minlength=len(lens)).astype(np.int64)
    new_lens = new_lens[~lost]
    # ambiguous residues
    nx = rng.binomial(len(new_seq), x_rate)
    if nx:
        new_seq[rng.integers(0, len(new_seq), nx)] = ord("X")
    return Species(new_seq, new_lens)


def species_set(n_species: int, n_prot: int, mean_len: float, seed: int = SEED0,
                p_sub_range=(0.0005, 0.01), max_len=3000, max_depth: float | None = None) -> list[Species]:
    """Random tree: every new species branches off a uniformly chosen earlier one.  max_depth caps the summed
    substitution probability from the root to any species (a branch that would pass it gets the minimum length),
    which keeps the pairwise identity span independent of the number of species."""
    rng = np.random.default_rng(seed)
    root = root_proteome(rng, n_prot, mean_len, hi=max_len)
    out = [evolve(rng, root, p_sub_range[0])]
    depth = [p_sub_range[0]]
    while len(out) < n_species:
        i = int(rng.integers(0, len(out)))
        p = float(rng.uniform(*p_sub_range))
        if max_depth is not None and depth[i] + p > max_depth:
            p = p_sub_range[0]
        out.append(evolve(rng, out[i], p))
        depth.append(depth[i] + p)
    return out


# Per-branch substitution probabilities of the named configurations, with the root-to-tip sum capped so that the
# pairwise identity span does not depend on the number of species.  kamino's subject is sets of ISOLATES (north_star:
# "per-isolate k-mer table"): the bacterial sets are ~97.7-99.9 % identical (root-to-tip <= 1.2 %), the eukaryotic
# "within-phylum" sets ~89-99.8 %.  SURVEY.md 8d suggests 80-99 %; measured with the oracle CLI (400 proteins per
# proteome, default flags), wider spans leave nothing to compare at the sizes of configs[2] and configs[4], because a
# variant group needs a 13-mer shared by 85 % of the species:
#     root-to-tip cap   identity        100 species            1000 species
#     8 %  (p <= 3 %)   85-99.9 %       0 groups               0 groups
#     4 %  (p <= 2 %)   92.5-99.9 %     978 groups, 13.5 kaa   3 groups, 14 aa
#     2 %  (p <= 1 %)   96-99.9 %       1545 groups, 20.7 kaa  68 groups, 587 aa
#     1.2 % (p <= 0.6%) 97.7-99.9 %     1882 groups, 20.0 kaa  187 groups, 1.4 kaa
# so the cap is set where 1,000 and 5,000 proteomes still give an alignment worth checking bit for bit.
BACTERIAL_P_SUB, BACTERIAL_MAX_DEPTH = (0.0003, 0.006), 0.012
EUKARYOTIC_P_SUB, EUKARYOTIC_MAX_DEPTH = (0.001, 0.03), 0.06


def bacterial(n_species: int, seed: int = SEED0 + 1, n_prot=4000, mean_len=330.0) -> list[Species]:
    return species_set(n_species, n_prot, mean_len, seed, p_sub_range=BACTERIAL_P_SUB, max_depth=BACTERIAL_MAX_DEPTH)


def eukaryotic(n_species: int, seed: int = SEED0 + 3, n_prot=20000, mean_len=500.0) -> list[Species]:
    return species_set(n_species, n_prot, mean_len, seed, p_sub_range=EUKARYOTIC_P_SUB, max_depth=EUKARYOTIC_MAX_DEPTH,
                       max_len=6000)


def _wrap_species(sp: Species, width: int) -> tuple[np.ndarray, np.ndarray]:
    """seq() bytes of every record (60-column lines, '\n' terminated) + record byte lengths."""
    lens = sp.lens
    if width <= 0:
        return sp.seq, lens.copy()
    n_lines = (lens + width - 1) // width
    out_lens = lens + n_lines
    rec_out_start = np.cumsum(out_lens) - out_lens
    rec_in_start = np.cumsum(lens) - lens
    off = np.arange(len(sp.seq), dtype=np.int64) - np.repeat(rec_in_start, lens)
    out_idx = np.repeat(rec_out_start, lens) + off + off // width
    out = np.full(int(out_lens.sum()), ord("\n"), dtype=np.uint8)
    out[out_idx] = sp.seq
    return out, out_lens


def stage(species: list[Species], width: int = 60) -> StagedBatch:
    parts, rec_lens, sp_rec = [], [], [0]
    for sp in species:
        b, l = _wrap_species(sp, width)
        parts.append(b)
        rec_lens.append(l)
        sp_rec.append(sp_rec[-1] + len(l))
    residues = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    lens = np.concatenate(rec_lens) if rec_lens else np.zeros(0, np.int64)
    rec_off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=rec_off[1:])
    return StagedBatch(residues, rec_off, np.asarray(sp_rec, dtype=np.uint64),
                       int(sum(int(s.lens.sum()) for s in species)))


def write_fasta_dir(species: list[Species], out_dir: Path, width: int = 60, gzip_every: int = 2) -> list[Path]:
    """One FASTA per species; headers exercise the ' [' truncation (group_extraction.rs:397-404)."""
    out_dir.mkdir(parents=True, exist_ok=True)
    paths = []
    for sid, sp in enumerate(species):
        chunks = []
        pos = 0
        for pid, ln in enumerate(sp.lens):
            chunks.append(f">sp{sid}_p{pid} hypothetical protein {pid} [Synth sp]\n".encode())
            s = sp.seq[pos:pos + ln].tobytes()
            pos += ln
            for i in range(0, len(s), width):
                chunks.append(s[i:i + width] + b"\n")
        data = b"".join(chunks)
        gz = gzip_every and sid % gzip_every == 1
        p = out_dir / (f"sp{sid:05d}.faa" + (".gz" if gz else ""))
        if gz:
            with gzip.open(p, "wb", compresslevel=1) as f:
                f.write(data)
        else:
            p.write_bytes(data)
        paths.append(p)
    return paths


def write_fasta_dir_fast(species: list[Species], out_dir: Path, width: int = 60) -> list[Path]:
    """Same files as write_fasta_dir without gzip, but laid out with numpy (no per-protein Python loop): headers
    have a fixed width (">sp00017_p003999 hypothetical protein 003999 [Synth sp]"), so that thousands of proteomes
    can be written in seconds (configs[2] / configs[4] runs through the CLI)."""
    out_dir.mkdir(parents=True, exist_ok=True)
    paths = []
    for sid, sp in enumerate(species):
        lens = sp.lens.astype(np.int64)
        n = len(lens)
        pid = np.arange(n)
        d6 = (pid[:, None] // (10 ** np.arange(5, -1, -1))[None, :] % 10 + 48).astype(np.uint8)
        head_a = np.frombuffer(f">sp{sid:05d}_p".encode(), np.uint8)
        head_b = np.frombuffer(b" hypothetical protein ", np.uint8)
        head_c = np.frombuffer(b" [Synth sp]\n", np.uint8)
        hdr = np.concatenate([np.broadcast_to(head_a, (n, head_a.size)), d6, np.broadcast_to(head_b, (n, head_b.size)), d6,
                              np.broadcast_to(head_c, (n, head_c.size))], axis=1)
        hlen = hdr.shape[1]
        n_lines = (lens + width - 1) // width
        body_len = lens + n_lines
        rec_len = hlen + body_len
        rec_start = np.cumsum(rec_len) - rec_len
        out = np.full(int(rec_len.sum()), ord("\n"), dtype=np.uint8)
        out[(rec_start[:, None] + np.arange(hlen)[None, :]).ravel()] = hdr.ravel()
        in_start = np.cumsum(lens) - lens
        off = np.arange(len(sp.seq), dtype=np.int64) - np.repeat(in_start, lens)
        out[np.repeat(rec_start + hlen, lens) + off + off // width] = sp.seq
        p = out_dir / f"sp{sid:05d}.faa"
        out.tofile(p)
        paths.append(p)
    return paths


CONFIGS = {
    # name: (factory, n_species) — BASELINE.json configs[1..4]
    "bacterial_100": (bacterial, 100),
    "bacterial_1000": (bacterial, 1000),
    "eukaryotic_200": (eukaryotic, 200),
    "bacterial_5000": (bacterial, 5000),
}
