"""TEST INFRASTRUCTURE — CPU restatement of /root/reference/src/protein_prediction.rs (`--genomes`), in plain
Python for small cases.  Only tests/ may import it; the product (`kamino_b200/csrc/host/protein_prediction.cpp`)
never does.  Written from the Rust source independently of the C++ host code so that the two can be compared
on synthetic genomes (`tests/test_cli_host.py`), and pinned by the reference's golden case `tests/data/test_genomes`
(through the rest of the oracle pipeline).  Each function cites the reference lines it follows.

Parity note: `train_adaptive_coding_model` sorts its seed ORFs with an UNSTABLE sort by score only
(protein_prediction.rs:961); ties on the truncation boundary are implementation-defined in the reference, so
contigs that reach the adaptive model (>= 60 complete ORFs of >= 120 aa) are "parity unpinned" in that respect.
Here, as in the product, ties keep discovery order.
"""
from __future__ import annotations

import gzip
import math
import struct
from pathlib import Path

DEFAULT_MIN_AA = 90                      # :17
MAX_ALLOWED_OVERLAP_NT = 30              # :19
GC_ESTIMATE_MAX_VALID_BASES = 300_000    # :22
MIN_START_SEPARATION_NT = 12             # :26
START_FILTER_LIMIT = 8                   # :29
ALT_START_TOTAL_GAP = 6                  # :30
ALT_START_MAX = 4                        # :31
SD_SCAN_UPSTREAM_MAX = 20                # :32
SD_SCAN_UPSTREAM_MIN = 4                 # :33
ADAPTIVE_TRAIN_MIN_AA = 120              # :34
ADAPTIVE_MIN_TRAIN_ORFS = 60             # :35

NT2 = {"A": 0, "C": 1, "G": 2, "T": 3}   # :223-230 (upper case only)
COMPLEMENT = {"A": "T", "a": "t", "T": "A", "t": "a", "C": "G", "c": "g", "G": "C", "g": "c"}   # :232-244, else N
CODON_TO_AA = ("KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF")   # :214-219
START_BONUS = {14: 8, 46: 4, 62: 1}      # ATG / GTG / TTG (:87-94, :246-252)
START_RANK = {14: 0, 46: 1, 62: 1}       # :96-103
STOPS = (48, 50, 56)                     # TAA TAG TGA (:254-260)
WS = " \t\n\x0c\r"


def f32(x: float) -> float:
    """Round a Python float to the nearest f32 (the reference computes the GC regressions in f32)."""
    return struct.unpack("f", struct.pack("f", x))[0]


def codon_index(a: str, b: str, c: str):          # :1051-1059
    x, y, z = NT2.get(a), NT2.get(b), NT2.get(c)
    if x is None or y is None or z is None:
        return None
    return (x << 4) | (y << 2) | z


def gc_model(genome_gc: float):                   # :120-134, f32 arithmetic
    def reg(slope, icpt, lo, hi):
        v = f32(f32(f32(slope) * genome_gc) + f32(icpt))
        v = min(max(v, f32(lo)), f32(hi))
        r = f32(v * f32(1000.0))
        return int(math.floor(r + 0.5)) if r >= 0 else -int(math.floor(-r + 0.5))   # f32::round: half away from zero
    return reg(0.684392, 0.233505, 0.20, 0.90), reg(0.432433, 0.183878, 0.15, 0.75), reg(1.887194, -0.394644, 0.05, 0.98)


def sd_score(seq: str, start: int) -> int:        # :778-836
    if start < SD_SCAN_UPSTREAM_MIN + 4:
        return 0
    lo = max(start - SD_SCAN_UPSTREAM_MAX, 0)
    hi = max(start - SD_SCAN_UPSTREAM_MIN, 0)
    best = 0
    i = lo
    while i + 4 <= hi:
        local = 0
        if i + 5 <= hi:
            m5 = (seq[i] == "A") + (seq[i + 1] == "G") + (seq[i + 2] == "G") + (seq[i + 3] == "A") + (seq[i + 4] == "G")
            local = 5 if m5 == 5 else 3 if m5 == 4 else 0
        if local == 0:
            purines = sum(ch in "AG" for ch in seq[i:i + 4])
            local = 1 if purines == 4 else 0
        if local > 0:
            motif_len = 5 if local >= 3 else 4
            spacer = start - (i + motif_len)
            local += 2 if 6 <= spacer <= 10 else 1 if spacer in (5, 11) else 0
        best = max(best, local)
        i += 1
    return best


def clamp(v, lo, hi):
    return lo if v < lo else hi if v > hi else v


def idiv(a: int, b: int) -> int:
    """Rust integer division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def length_score(nt_len: int) -> int:             # :881-891
    aa = nt_len // 3
    if aa < 20:
        return -20
    return (60 * aa) // (aa + 80)


def calibrated_start_term(start_score: int, nt_len: int) -> int:   # :869-877
    if nt_len >= 600:
        return clamp(idiv(start_score, 2), -6, 7)
    if nt_len >= 300:
        return clamp(idiv(2 * start_score, 3), -8, 9)
    return clamp(start_score, -10, 10)


def total_orf_score(nt_len, has_start, start_score, gc_score, p5, p3) -> int:   # :840-865
    score = length_score(nt_len) + gc_score + calibrated_start_term(start_score, nt_len)
    if not has_start:
        score -= 10
    if p5:
        score -= 8
    if p3:
        score -= 8
    return score


def gcpos_score_pm(s: str, model) -> int:          # :895-935
    codons = len(s) // 3
    if codons < 20:
        return 0
    usable = s[:codons * 3]
    gc1 = sum(ch in "GC" for ch in usable[0::3])
    gc2 = sum(ch in "GC" for ch in usable[1::3])
    gc3 = sum(ch in "GC" for ch in usable[2::3])
    d1 = abs(gc1 * 1000 // codons - model[0])
    d2 = abs(gc2 * 1000 // codons - model[1])
    d3 = abs(gc3 * 1000 // codons - model[2])
    return 90 - (d1 + d2 + 2 * d3) // 25 + min(codons, 400) // 40


def suppress_nearby_starts(starts):               # :716-738; start = (pos, codon index, start_score)
    kept = []
    for s in starts:
        for j, k in enumerate(kept):
            if abs(s[0] - k[0]) < MIN_START_SEPARATION_NT:
                if s[2] > k[2] or (s[2] == k[2] and START_RANK[s[1]] < START_RANK[k[1]]):
                    kept[j] = s
                break
        else:
            kept.append(s)
    return kept


def emit_segment(seq, strand, frame, segment_end, cands, min_orf_nt, model, out):   # :586-712
    n = len(seq)
    frame_end = n - ((n - frame) % 3)
    touches_3p = segment_end == frame_end
    ranked = []
    for pos, codon, sscore in suppress_nearby_starts(cands):
        if pos >= segment_end:
            continue
        len_nt = segment_end - pos
        if len_nt < min_orf_nt:
            continue
        total = total_orf_score(len_nt, True, sscore, gcpos_score_pm(seq[pos:segment_end], model), False, touches_3p)
        ranked.append((-total, -sscore, START_RANK[codon], -len_nt, pos))
    if not ranked:
        return
    ranked.sort()                                  # :683-696 (a total order: positions are unique)
    top_total = -ranked[0][0]
    for idx, (neg_total, _, _, _, pos) in enumerate(ranked):
        if idx >= START_FILTER_LIMIT or idx >= ALT_START_MAX:
            break
        if idx > 0 and top_total - (-neg_total) > ALT_START_TOTAL_GAP:
            break
        start, end = (pos, segment_end) if strand == 1 else (n - segment_end, n - pos)   # :752-756
        out.append({"start": start, "end": end, "strand": strand, "scan_start": pos, "scan_end": segment_end,
                    "p5": False, "p3": touches_3p, "has_start": True, "score": -neg_total})


def find_orfs_on_strand(seq, strand, min_orf_nt, model, out):   # :492-582
    n = len(seq)
    for frame in range(3):
        if frame >= n:
            continue
        cands = []
        i = frame
        while i + 3 <= n:
            a, b, c = seq[i], seq[i + 1], seq[i + 2]
            if a == "N" or b == "N" or c == "N":
                emit_segment(seq, strand, frame, i, cands, min_orf_nt, model, out)
                cands = []
                i += 3
                continue
            idx = codon_index(a, b, c)
            if idx in START_BONUS:
                cands.append((i, idx, START_BONUS[idx] + sd_score(seq, i)))
            if idx in STOPS:
                emit_segment(seq, strand, frame, i, cands, min_orf_nt, model, out)
                cands = []
            i += 3
        emit_segment(seq, strand, frame, n - ((n - frame) % 3), cands, min_orf_nt, model, out)


def codon_counts(s: str, counts):                 # :1037-1047
    for i in range(0, (len(s) // 3) * 3, 3):
        idx = codon_index(s[i], s[i + 1], s[i + 2])
        if idx is not None:
            counts[idx] += 1


def train_adaptive(orfs, fwd, rev):               # :940-1011
    seed = [o for o in orfs if not o["p5"] and not o["p3"] and o["has_start"]
            and (o["end"] - o["start"]) // 3 >= ADAPTIVE_TRAIN_MIN_AA]
    if len(seed) < ADAPTIVE_MIN_TRAIN_ORFS:
        return None
    seed.sort(key=lambda o: -o["score"])           # stable here; unstable in the reference (module docstring)
    seed = seed[:min(max(len(seed) // 2, ADAPTIVE_MIN_TRAIN_ORFS), len(seed))]
    coding, background = [1] * 64, [1] * 64
    for o in seed:
        s = fwd if o["strand"] == 1 else rev
        nt = s[o["scan_start"]:o["scan_end"]]
        if (len(nt) // 3) * 3 >= 3:
            codon_counts(nt, coding)
    codon_counts(fwd, background)
    codon_counts(rev, background)
    ct, bt = sum(coding), sum(background)
    model = []
    for i in range(64):
        lod2 = math.log2((coding[i] / ct) / (background[i] / bt))
        model.append(int(clamp(lod2 * 1000.0, -2200.0, 2200.0)))   # `as i16` truncates toward zero
    return model


def adaptive_coding_score(nt: str, model) -> int:  # :1015-1033
    codons = len(nt) // 3
    if codons < 35:
        return 0
    total = 0
    for i in range(0, codons * 3, 3):
        idx = codon_index(nt[i], nt[i + 1], nt[i + 2])
        if idx is not None:
            total += model[idx]
    return clamp(idiv(idiv(total, codons), 25), -18, 18)


def greedy_non_overlapping(orfs):                 # :462-488
    order = sorted(orfs, key=lambda o: (int(o["p5"]) + int(o["p3"]), -o["score"], -(o["end"] - o["start"]), o["start"]))
    accepted = []
    for cand in order:
        if all(max(0, min(cand["end"], a["end"]) - max(cand["start"], a["start"])) <= MAX_ALLOWED_OVERLAP_NT for a in accepted):
            accepted.append(cand)
    return accepted


def translate_table11(nt: str, force_initial_m: bool) -> str:   # :1113-1139
    out, i = [], 0
    if force_initial_m and len(nt) >= 3:
        out.append("M")
        i = 3
    while i + 3 <= len(nt):
        idx = codon_index(nt[i], nt[i + 1], nt[i + 2])
        out.append(CODON_TO_AA[idx] if idx is not None else "X")
        i += 3
    return "".join(out)


def read_records(path: Path):
    """seq_io::fasta::Reader: a record starts at every line beginning with '>', seq() keeps the line breaks;
    gunzip iff the extension is 'gz' in any case (:357-372)."""
    raw = Path(path).read_bytes()
    if Path(path).suffix.lower() == ".gz":
        raw = gzip.decompress(raw)
    text = raw.decode("latin-1")
    pos = 0
    while pos < len(text) and (text[pos] == "\n" or text.startswith("\r\n", pos)):
        pos += 1 if text[pos] == "\n" else 2
    if pos >= len(text):
        return []
    if text[pos] != ">":
        raise ValueError(f"FASTA parse error in {path}: expected '>' at record start")
    recs, lines, cur = [], text[pos:].split("\n"), None
    for k, line in enumerate(lines):
        if line.startswith(">"):
            cur = []
            recs.append(cur)
        elif cur is not None:
            cur.append(line + ("\n" if k + 1 < len(lines) else ""))
    return ["".join(r) for r in recs]


def predict_one_genome(path: Path) -> str:         # :298-354 (+ :398-448); returns the text of `<name>.faa`
    recs = read_records(path)
    gc = valid = total = 0
    for rec in recs:
        if valid >= GC_ESTIMATE_MAX_VALID_BASES:
            break
        for b in rec:
            if b in WS:
                continue
            if total >= GC_ESTIMATE_MAX_VALID_BASES:
                break
            total += 1
            if b in "GCgc":
                gc += 1
                valid += 1
            elif b in "ATat":
                valid += 1
    genome_gc = f32(0.5) if valid == 0 else f32(f32(gc) / f32(valid))
    atgc = 0.0 if total == 0 else f32(f32(valid) / f32(total))
    if atgc < 0.5:
        raise ValueError(f"the file {path} does not appear to contain DNA sequences.")
    model = gc_model(genome_gc)
    out, count = [], 0
    for rec in recs:
        seq = "".join(ch for ch in rec if ch not in WS).upper()
        if len(seq) < 3:
            continue
        rc = "".join(COMPLEMENT.get(ch, "N") for ch in reversed(seq))
        orfs = []
        find_orfs_on_strand(seq, 1, DEFAULT_MIN_AA * 3, model, orfs)
        find_orfs_on_strand(rc, -1, DEFAULT_MIN_AA * 3, model, orfs)
        adaptive = train_adaptive(orfs, seq, rc)
        if adaptive is not None:
            for o in orfs:
                s = seq if o["strand"] == 1 else rc
                o["score"] += adaptive_coding_score(s[o["scan_start"]:o["scan_end"]], adaptive)
        for o in greedy_non_overlapping(orfs):
            s = seq if o["strand"] == 1 else rc
            protein = translate_table11(s[o["scan_start"]:o["scan_end"]], o["has_start"])
            count += 1
            out.append(f">protein_{count}\n")
            out.extend(protein[i:i + 60] + "\n" for i in range(0, len(protein), 60))
    return "".join(out)
