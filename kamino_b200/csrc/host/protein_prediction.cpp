// protein_prediction.cpp — `--genomes`: six-frame bacterial ORF prediction in front of the normal pipeline
// ("next" row 4 of SURVEY 8f; host logic, no GPU work).  Equivalent to
// /root/reference/src/protein_prediction.rs:265-1150: every nucleotide FASTA is scanned on both strands, the
// stop-free segments of each frame yield start candidates scored by start codon + Shine-Dalgarno motif,
// ORFs are ranked by length / codon-position GC / start terms (+ an adaptive codon-usage model when a contig
// has enough long ORFs), pruned greedily by overlap and translated with table 11 into `<name>.faa`.
// Integer arithmetic throughout except the three f32 regressions of GcModel and the f64 log-odds of the
// adaptive model, which are kept at the reference's precision (-ffp-contract=off).
//
// One documented divergence risk: the reference orders the adaptive model's seed ORFs with an UNSTABLE sort
// by score only (protein_prediction.rs:961), so which of several equally scored ORFs sit on the truncation
// boundary depends on the Rust standard library's sort; here ties keep their discovery order.
#include "kamino_host.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <zlib.h>

namespace kh {

namespace {

constexpr size_t DEFAULT_MIN_AA = 90;                    // :17
constexpr size_t MAX_ALLOWED_OVERLAP_NT = 30;            // :19
constexpr size_t GC_ESTIMATE_MAX_VALID_BASES = 300000;   // :22
constexpr size_t MIN_START_SEPARATION_NT = 12;           // :26
constexpr size_t START_FILTER_LIMIT = 8;                 // :29
constexpr int ALT_START_TOTAL_GAP = 6;                   // :30
constexpr size_t ALT_START_MAX = 4;                      // :31
constexpr size_t SD_SCAN_UPSTREAM_MAX = 20;              // :32
constexpr size_t SD_SCAN_UPSTREAM_MIN = 4;               // :33
constexpr size_t ADAPTIVE_TRAIN_MIN_AA = 120;            // :34
constexpr size_t ADAPTIVE_MIN_TRAIN_ORFS = 60;           // :35

enum StartCodon : uint8_t { ATG = 0, GTG = 1, TTG = 2 };
int start_bonus(StartCodon s) { return s == ATG ? 8 : s == GTG ? 4 : 1; }   // :87-94
int start_rank(StartCodon s) { return s == ATG ? 0 : 1; }                   // :96-103

struct GcModel { int exp1_pm, exp2_pm, exp3_pm; };        // :106-134

GcModel gc_model_from(float genome_gc) {
    auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); };
    const float gc1 = clampf(0.684392f * genome_gc + 0.233505f, 0.20f, 0.90f);
    const float gc2 = clampf(0.432433f * genome_gc + 0.183878f, 0.15f, 0.75f);
    const float gc3 = clampf(1.887194f * genome_gc - 0.394644f, 0.05f, 0.98f);
    return {int(uint16_t(std::round(gc1 * 1000.0f))), int(uint16_t(std::round(gc2 * 1000.0f))),
            int(uint16_t(std::round(gc3 * 1000.0f)))};
}

struct Orf {                                              // :136-172
    size_t start, end;        // forward-strand coordinates, half open
    int strand;
    size_t scan_start, scan_end;
    bool partial_5p, partial_3p, has_start_codon;
    int total_score;
    size_t len_nt() const { return end - start; }
    int partial_class() const { return int(partial_5p) + int(partial_3p); }
};

struct CandidateStart { size_t pos; StartCodon codon; int start_score; };
struct ValidStart { CandidateStart cand; size_t len_nt; };

int nt2(uint8_t b) { return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1; }   // :223-230
int codon_index(uint8_t a, uint8_t b, uint8_t c) {                                             // :1051-1059
    const int x = nt2(a), y = nt2(b), z = nt2(c);
    if (x < 0 || y < 0 || z < 0) return -1;
    return (x << 4) | (y << 2) | z;
}
const char CODON_TO_AA[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";   // :214-219

bool is_stop(uint8_t a, uint8_t b, uint8_t c) {           // :254-260, :1079-1085
    const int i = codon_index(a, b, c);
    return i == 48 || i == 50 || i == 56;
}
bool start_codon_of(uint8_t a, uint8_t b, uint8_t c, StartCodon* out) {   // :246-252, :1089-1095
    const int i = codon_index(a, b, c);
    if (i == 14) { *out = ATG; return true; }
    if (i == 46) { *out = GTG; return true; }
    if (i == 62) { *out = TTG; return true; }
    return false;
}
bool is_purine(uint8_t b) { return b == 'A' || b == 'G'; }
bool is_gc(uint8_t b) { return b == 'G' || b == 'C'; }

uint8_t complement(uint8_t b) {                           // :232-244
    switch (b) {
        case 'A': return 'T'; case 'a': return 't'; case 'T': return 'A'; case 't': return 'a';
        case 'C': return 'G'; case 'c': return 'g'; case 'G': return 'C'; case 'g': return 'c';
        default: return 'N';
    }
}

int sd_score(const std::vector<uint8_t>& seq, size_t start) {   // :778-836
    if (start < SD_SCAN_UPSTREAM_MIN + 4) return 0;
    const size_t lo = start > SD_SCAN_UPSTREAM_MAX ? start - SD_SCAN_UPSTREAM_MAX : 0;
    const size_t hi = start - SD_SCAN_UPSTREAM_MIN;
    int best = 0;
    for (size_t i = lo; i + 4 <= hi; i++) {
        int local = 0;
        if (i + 5 <= hi) {
            const int m5 = int(seq[i] == 'A') + int(seq[i + 1] == 'G') + int(seq[i + 2] == 'G') + int(seq[i + 3] == 'A') +
                           int(seq[i + 4] == 'G');
            local = m5 == 5 ? 5 : m5 == 4 ? 3 : 0;
        }
        if (local == 0) {
            const int pur = int(is_purine(seq[i])) + int(is_purine(seq[i + 1])) + int(is_purine(seq[i + 2])) + int(is_purine(seq[i + 3]));
            local = pur == 4 ? 1 : 0;
        }
        if (local > 0) {
            const size_t motif_len = local >= 3 ? 5 : 4;
            const size_t spacer = start - (i + motif_len);
            local += (spacer >= 6 && spacer <= 10) ? 2 : (spacer == 5 || spacer == 11) ? 1 : 0;
        }
        best = std::max(best, local);
    }
    return best;
}

int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

int length_score(size_t nt_len) {                         // :881-891
    const size_t aa = nt_len / 3;
    if (aa < 20) return -20;
    const int x = int(aa);
    return (60 * x) / (x + 80);
}
int calibrated_start_term(int start_score, size_t nt_len) {   // :869-877
    if (nt_len >= 600) return clampi(start_score / 2, -6, 7);
    if (nt_len >= 300) return clampi((2 * start_score) / 3, -8, 9);
    return clampi(start_score, -10, 10);
}
int total_orf_score(size_t nt_len, bool has_start, int start_score, int gc_score, bool p5, bool p3) {   // :840-865
    int score = length_score(nt_len) + gc_score + calibrated_start_term(start_score, nt_len);
    if (!has_start) score -= 10;
    if (p5) score -= 8;
    if (p3) score -= 8;
    return score;
}
int gcpos_score_pm(const uint8_t* s, size_t n, const GcModel& m) {   // :895-935
    const size_t codons = n / 3;
    if (codons < 20) return 0;
    size_t gc1 = 0, gc2 = 0, gc3 = 0;
    for (size_t i = 0; i < codons * 3; i += 3) {
        gc1 += is_gc(s[i]);
        gc2 += is_gc(s[i + 1]);
        gc3 += is_gc(s[i + 2]);
    }
    const int o1 = int((gc1 * 1000) / codons), o2 = int((gc2 * 1000) / codons), o3 = int((gc3 * 1000) / codons);
    const int pen = std::abs(o1 - m.exp1_pm) + std::abs(o2 - m.exp2_pm) + 2 * std::abs(o3 - m.exp3_pm);
    return 90 - pen / 25 + int(std::min<size_t>(codons, 400) / 40);
}

void suppress_nearby_starts(const std::vector<CandidateStart>& starts, std::vector<CandidateStart>& kept) {   // :716-738
    kept.clear();
    for (const CandidateStart& s : starts) {
        bool merged = false;
        for (CandidateStart& k : kept) {
            const size_t dist = s.pos > k.pos ? s.pos - k.pos : k.pos - s.pos;
            if (dist < MIN_START_SEPARATION_NT) {
                if (s.start_score > k.start_score || (s.start_score == k.start_score && start_rank(s.codon) < start_rank(k.codon))) k = s;
                merged = true;
                break;
            }
        }
        if (!merged) kept.push_back(s);
    }
}

void push_orf(int strand, size_t seq_len, size_t scan_start, size_t scan_end, bool p3, int total, std::vector<Orf>& out) {   // :741-770
    Orf o;
    o.strand = strand;
    o.start = strand == 1 ? scan_start : seq_len - scan_end;
    o.end = strand == 1 ? scan_end : seq_len - scan_start;
    o.scan_start = scan_start;
    o.scan_end = scan_end;
    o.partial_5p = false;
    o.partial_3p = p3;
    o.has_start_codon = true;
    o.total_score = total;
    out.push_back(o);
}

// one stop-free segment of a frame: emit_segment_orfs + collect_valid_starts + emit_complete_orfs (:586-712)
void emit_segment(const std::vector<uint8_t>& seq, int strand, size_t frame, size_t segment_end,
                  const std::vector<CandidateStart>& cands, size_t min_orf_nt, const GcModel& gm,
                  std::vector<CandidateStart>& kept, std::vector<Orf>& out) {
    const size_t len = seq.size();
    const size_t frame_end = len - ((len - frame) % 3);
    const bool touches_3p = segment_end == frame_end;
    suppress_nearby_starts(cands, kept);
    struct Ranked { ValidStart vs; int total; };
    std::vector<Ranked> ranked;
    for (const CandidateStart& s : kept) {
        if (s.pos >= segment_end) continue;
        const size_t len_nt = segment_end - s.pos;
        if (len_nt < min_orf_nt) continue;
        const int gc = gcpos_score_pm(seq.data() + s.pos, len_nt, gm);
        ranked.push_back({{s, len_nt}, total_orf_score(len_nt, true, s.start_score, gc, false, touches_3p)});
    }
    if (ranked.empty()) return;
    std::sort(ranked.begin(), ranked.end(), [](const Ranked& a, const Ranked& b) {   // :683-696, a total order
        if (a.total != b.total) return a.total > b.total;
        if (a.vs.cand.start_score != b.vs.cand.start_score) return a.vs.cand.start_score > b.vs.cand.start_score;
        if (start_rank(a.vs.cand.codon) != start_rank(b.vs.cand.codon)) return start_rank(a.vs.cand.codon) < start_rank(b.vs.cand.codon);
        if (a.vs.len_nt != b.vs.len_nt) return a.vs.len_nt > b.vs.len_nt;
        return a.vs.cand.pos < b.vs.cand.pos;
    });
    const int top = ranked[0].total;
    for (size_t i = 0; i < ranked.size(); i++) {
        if (i >= START_FILTER_LIMIT || i >= ALT_START_MAX) break;
        if (i > 0 && top - ranked[i].total > ALT_START_TOTAL_GAP) break;
        push_orf(strand, len, ranked[i].vs.cand.pos, segment_end, touches_3p, ranked[i].total, out);
    }
}

void find_orfs_on_strand(const std::vector<uint8_t>& seq, int strand, size_t min_orf_nt, const GcModel& gm, std::vector<Orf>& out) {   // :492-582
    const size_t len = seq.size();
    std::vector<CandidateStart> kept, cands;
    for (size_t frame = 0; frame < 3; frame++) {
        if (frame >= len) continue;
        cands.clear();
        size_t i = frame;
        while (i + 3 <= len) {
            const uint8_t a = seq[i], b = seq[i + 1], c = seq[i + 2];
            if (a == 'N' || b == 'N' || c == 'N') {
                emit_segment(seq, strand, frame, i, cands, min_orf_nt, gm, kept, out);
                cands.clear();
                i += 3;
                continue;
            }
            StartCodon sc;
            if (start_codon_of(a, b, c, &sc)) cands.push_back({i, sc, start_bonus(sc) + sd_score(seq, i)});
            if (is_stop(a, b, c)) {
                emit_segment(seq, strand, frame, i, cands, min_orf_nt, gm, kept, out);
                cands.clear();
            }
            i += 3;
        }
        const size_t frame_end = len - ((len - frame) % 3);
        emit_segment(seq, strand, frame, frame_end, cands, min_orf_nt, gm, kept, out);
    }
}

void accumulate_codon_counts(const uint8_t* s, size_t n, size_t* counts) {   // :1037-1047
    for (size_t i = 0; i + 3 <= (n / 3) * 3; i += 3) {
        const int idx = codon_index(s[i], s[i + 1], s[i + 2]);
        if (idx >= 0) counts[idx]++;
    }
}

// train_adaptive_coding_model (:940-1011); false = not enough training ORFs
bool train_adaptive(const std::vector<Orf>& orfs, const std::vector<uint8_t>& fwd, const std::vector<uint8_t>& rev, int16_t* score_pm) {
    std::vector<Orf> seed;
    for (const Orf& o : orfs)
        if (!o.partial_5p && !o.partial_3p && o.has_start_codon && o.len_nt() / 3 >= ADAPTIVE_TRAIN_MIN_AA) seed.push_back(o);
    if (seed.size() < ADAPTIVE_MIN_TRAIN_ORFS) return false;
    std::stable_sort(seed.begin(), seed.end(), [](const Orf& a, const Orf& b) { return a.total_score > b.total_score; });
    const size_t keep = std::min(std::max(seed.size() / 2, ADAPTIVE_MIN_TRAIN_ORFS), seed.size());
    seed.resize(keep);
    size_t coding[64], background[64];
    std::fill(coding, coding + 64, size_t(1));
    std::fill(background, background + 64, size_t(1));
    for (const Orf& o : seed) {
        const std::vector<uint8_t>& s = o.strand == 1 ? fwd : rev;
        const size_t n = o.scan_end - o.scan_start;
        if ((n / 3) * 3 >= 3) accumulate_codon_counts(s.data() + o.scan_start, (n / 3) * 3, coding);
    }
    if ((fwd.size() / 3) * 3 >= 3) accumulate_codon_counts(fwd.data(), (fwd.size() / 3) * 3, background);
    if ((rev.size() / 3) * 3 >= 3) accumulate_codon_counts(rev.data(), (rev.size() / 3) * 3, background);
    size_t ct = 0, bt = 0;
    for (int i = 0; i < 64; i++) { ct += coding[i]; bt += background[i]; }
    for (int i = 0; i < 64; i++) {
        const double pc = double(coding[i]) / double(ct), pb = double(background[i]) / double(bt);
        double v = std::log2(pc / pb) * 1000.0;
        v = v < -2200.0 ? -2200.0 : (v > 2200.0 ? 2200.0 : v);
        score_pm[i] = int16_t(v);   // `as i16`: toward zero
    }
    return true;
}

int adaptive_coding_score(const uint8_t* nt, size_t n, const int16_t* score_pm) {   // :1015-1033
    const size_t codons = n / 3;
    if (codons < 35) return 0;
    int sum = 0;
    for (size_t i = 0; i < codons * 3; i += 3) {
        const int idx = codon_index(nt[i], nt[i + 1], nt[i + 2]);
        if (idx >= 0) sum += score_pm[idx];
    }
    return clampi((sum / int(codons)) / 25, -18, 18);
}

std::vector<Orf> greedy_non_overlapping(std::vector<Orf>& orfs) {   // :462-488
    std::sort(orfs.begin(), orfs.end(), [](const Orf& a, const Orf& b) {
        if (a.partial_class() != b.partial_class()) return a.partial_class() < b.partial_class();
        if (a.total_score != b.total_score) return a.total_score > b.total_score;
        if (a.len_nt() != b.len_nt()) return a.len_nt() > b.len_nt();
        return a.start < b.start;
    });
    std::vector<Orf> accepted;
    for (const Orf& cand : orfs) {
        bool ok = true;
        for (const Orf& a : accepted) {
            const size_t s = std::max(cand.start, a.start), e = std::min(cand.end, a.end);
            if (e > s && e - s > MAX_ALLOWED_OVERLAP_NT) { ok = false; break; }
        }
        if (ok) accepted.push_back(cand);
    }
    return accepted;
}

std::string translate_table11(const uint8_t* nt, size_t n, bool force_initial_m) {   // :1113-1139
    std::string p;
    p.reserve(n / 3);
    size_t i = 0;
    if (force_initial_m && n >= 3) { p.push_back('M'); i = 3; }
    for (; i + 3 <= n; i += 3) {
        const int idx = codon_index(nt[i], nt[i + 1], nt[i + 2]);
        p.push_back(idx >= 0 ? CODON_TO_AA[idx] : 'X');
    }
    return p;
}

bool is_ws(uint8_t c) { return c == 0x20 || c == 0x09 || c == 0x0A || c == 0x0C || c == 0x0D; }

// open_fasta_reader (:357-372): gunzip iff the extension is "gz" in any case
std::vector<uint8_t> slurp_genome(const std::string& path) {
    const size_t slash = path.find_last_of('/');
    const std::string file = slash == std::string::npos ? path : path.substr(slash + 1);
    const size_t dot = file.find_last_of('.');
    std::string ext = dot == std::string::npos || dot == 0 ? std::string() : file.substr(dot + 1);
    for (char& ch : ext) ch = char(std::tolower((unsigned char)ch));
    std::vector<uint8_t> data;
    if (ext == "gz") {
        gzFile g = gzopen(path.c_str(), "rb");
        if (!g) throw std::runtime_error("Failed to open input file: \"" + path + "\"");
        if (gzdirect(g)) { gzclose(g); throw std::runtime_error("Failed to read FASTA record: invalid gzip header in \"" + path + "\""); }
        gzbuffer(g, 1 << 20);
        std::vector<uint8_t> buf(1 << 20);
        for (;;) {
            const int got = gzread(g, buf.data(), unsigned(buf.size()));
            if (got < 0) { gzclose(g); throw std::runtime_error("Failed to read FASTA record: gzip error in \"" + path + "\""); }
            if (got == 0) break;
            data.insert(data.end(), buf.begin(), buf.begin() + got);
        }
        gzclose(g);
    } else {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("Failed to open input file: \"" + path + "\"");
        std::vector<uint8_t> buf(1 << 20);
        size_t got;
        while ((got = std::fread(buf.data(), 1, buf.size(), f)) > 0) data.insert(data.end(), buf.begin(), buf.begin() + got);
        std::fclose(f);
    }
    return data;
}

// predict_one_genome (:298-354) + estimate_input_genome_gc (:398-448)
void predict_one_genome(const std::string& in_path, const std::string& out_path) {
    const size_t min_orf_nt = DEFAULT_MIN_AA * 3;
    const Proteome recs = parse_proteome(slurp_genome(in_path), in_path);   // seq_io records; sequences keep their line breaks
    size_t gc = 0, valid = 0, total = 0;
    for (size_t r = 0; r < recs.n_rec(); r++) {
        if (valid >= GC_ESTIMATE_MAX_VALID_BASES) break;
        for (uint64_t i = recs.rec_begin[r]; i < recs.rec_end[r]; i++) {
            const uint8_t b = recs.seq[i];
            if (is_ws(b)) continue;
            if (total >= GC_ESTIMATE_MAX_VALID_BASES) break;
            total++;
            if (b == 'G' || b == 'C' || b == 'g' || b == 'c') { gc++; valid++; }
            else if (b == 'A' || b == 'T' || b == 'a' || b == 't') valid++;
        }
    }
    const float genome_gc = valid == 0 ? 0.5f : float(gc) / float(valid);
    const float atgc = total == 0 ? 0.0f : float(valid) / float(total);
    if (atgc < 0.5f) throw std::runtime_error("the file " + in_path + " does not appear to contain DNA sequences.");
    const GcModel gm = gc_model_from(genome_gc);

    FILE* out = std::fopen(out_path.c_str(), "wb");
    if (!out) throw std::runtime_error("Failed to create output file: \"" + out_path + "\"");
    size_t protein_count = 0;
    std::vector<uint8_t> seq, rc;
    std::vector<Orf> orfs;
    for (size_t r = 0; r < recs.n_rec(); r++) {
        seq.clear();
        for (uint64_t i = recs.rec_begin[r]; i < recs.rec_end[r]; i++) {
            const uint8_t b = recs.seq[i];
            if (is_ws(b)) continue;
            seq.push_back((b >= 'a' && b <= 'z') ? uint8_t(b - 32) : b);
        }
        if (seq.size() < 3) continue;
        rc.assign(seq.size(), 'N');
        for (size_t i = 0; i < seq.size(); i++) rc[seq.size() - 1 - i] = complement(seq[i]);
        orfs.clear();
        find_orfs_on_strand(seq, 1, min_orf_nt, gm, orfs);
        find_orfs_on_strand(rc, -1, min_orf_nt, gm, orfs);
        int16_t model[64];
        if (train_adaptive(orfs, seq, rc, model))
            for (Orf& o : orfs) {
                const std::vector<uint8_t>& s = o.strand == 1 ? seq : rc;
                o.total_score += adaptive_coding_score(s.data() + o.scan_start, o.scan_end - o.scan_start, model);
            }
        for (const Orf& o : greedy_non_overlapping(orfs)) {
            const std::vector<uint8_t>& s = o.strand == 1 ? seq : rc;
            const std::string protein = translate_table11(s.data() + o.scan_start, o.scan_end - o.scan_start, o.has_start_codon);
            protein_count++;
            std::fprintf(out, ">protein_%zu\n", protein_count);
            for (size_t i = 0; i < protein.size(); i += 60) {
                std::fwrite(protein.data() + i, 1, std::min<size_t>(60, protein.size() - i), out);
                std::fputc('\n', out);
            }
        }
    }
    if (std::fclose(out) != 0) throw std::runtime_error("Failed to flush output file: \"" + out_path + "\"");
}

Orf complete_orf(size_t a, size_t b, int strand, int score) {
    Orf o{};
    o.start = o.scan_start = a;
    o.end = o.scan_end = b;
    o.strand = strand;
    o.has_start_codon = true;
    o.total_score = score;
    return o;
}

std::vector<uint8_t> repeat(const char* unit, size_t times) {
    std::vector<uint8_t> v;
    for (size_t i = 0; i < times; i++) v.insert(v.end(), unit, unit + std::strlen(unit));
    return v;
}

}  // namespace

// The reference's own unit tests of this module (protein_prediction.rs:1152-1286), restated on the functions
// above; returns a description of the first failing check, or an empty string.
std::string prediction_self_test() {
#define KH_CHECK(cond) do { if (!(cond)) return std::string("failed: ") + #cond; } while (0)
    // lookup_tables_encode_bases_complements_starts_and_stops
    KH_CHECK(nt2('A') == 0 && nt2('C') == 1 && nt2('G') == 2 && nt2('T') == 3 && nt2('N') < 0 && nt2('a') < 0);
    KH_CHECK(complement('A') == 'T' && complement('T') == 'A' && complement('C') == 'G' && complement('G') == 'C');
    KH_CHECK(complement('a') == 't' && complement('n') == 'N');
    StartCodon sc;
    KH_CHECK(start_codon_of('A', 'T', 'G', &sc) && sc == ATG);
    KH_CHECK(start_codon_of('G', 'T', 'G', &sc) && sc == GTG);
    KH_CHECK(start_codon_of('T', 'T', 'G', &sc) && sc == TTG);
    KH_CHECK(!start_codon_of('A', 'C', 'G', &sc));
    KH_CHECK(is_stop('T', 'A', 'A') && is_stop('T', 'A', 'G') && is_stop('T', 'G', 'A') && !is_stop('T', 'G', 'G'));
    const int gct = codon_index('G', 'C', 'T'), aaa = codon_index('A', 'A', 'A');
    {   // adaptive_model_training_uses_high_scoring_complete_orfs_and_background
        const std::vector<uint8_t> rich = repeat("GCT", 130), poor = repeat("AAA", 130);
        std::vector<uint8_t> fwd;
        std::vector<Orf> orfs;
        for (size_t i = 0; i < ADAPTIVE_MIN_TRAIN_ORFS; i++) {
            const size_t st = fwd.size();
            fwd.insert(fwd.end(), rich.begin(), rich.end());
            orfs.push_back(complete_orf(st, fwd.size(), 1, 10000 - int(i)));
        }
        for (size_t i = 0; i < ADAPTIVE_MIN_TRAIN_ORFS; i++) {
            const size_t st = fwd.size();
            fwd.insert(fwd.end(), poor.begin(), poor.end());
            orfs.push_back(complete_orf(st, fwd.size(), 1, int(i)));
        }
        const std::vector<uint8_t> rev = repeat("CCC", ADAPTIVE_MIN_TRAIN_ORFS * 130);
        int16_t model[64];
        KH_CHECK(train_adaptive(orfs, fwd, rev, model));
        KH_CHECK(model[gct] > 0 && model[aaa] < 0);
    }
    {   // adaptive_model_training_rejects_insufficient_or_partial_seed_orfs
        const std::vector<uint8_t> fwd = repeat("ATG", ADAPTIVE_MIN_TRAIN_ORFS * ADAPTIVE_TRAIN_MIN_AA);
        const std::vector<uint8_t> rev = repeat("CAT", ADAPTIVE_MIN_TRAIN_ORFS * ADAPTIVE_TRAIN_MIN_AA);
        int16_t model[64];
        std::vector<Orf> too_few(ADAPTIVE_MIN_TRAIN_ORFS - 1, complete_orf(0, ADAPTIVE_TRAIN_MIN_AA * 3, 1, 0));
        KH_CHECK(!train_adaptive(too_few, fwd, rev, model));
        std::vector<Orf> partial(ADAPTIVE_MIN_TRAIN_ORFS, complete_orf(0, ADAPTIVE_TRAIN_MIN_AA * 3, 1, 0));
        for (Orf& o : partial) o.partial_5p = true;
        KH_CHECK(!train_adaptive(partial, fwd, rev, model));
    }
    {   // adaptive_scores_are_clamped_and_added_per_strand
        int16_t model[64] = {};
        model[gct] = 2200;
        model[aaa] = -2200;
        const std::vector<uint8_t> g34 = repeat("GCT", 34), g35 = repeat("GCT", 35), a35 = repeat("AAA", 35);
        KH_CHECK(adaptive_coding_score(g34.data(), g34.size(), model) == 0);
        KH_CHECK(adaptive_coding_score(g35.data(), g35.size(), model) == 18);
        KH_CHECK(adaptive_coding_score(a35.data(), a35.size(), model) == -18);
    }
    {   // accumulate_codon_counts_counts_only_complete_valid_codons
        size_t counts[64] = {};
        const char* s = "ATGGCTNNNT";
        accumulate_codon_counts(reinterpret_cast<const uint8_t*>(s), 10, counts);
        size_t sum = 0;
        for (size_t c : counts) sum += c;
        KH_CHECK(counts[codon_index('A', 'T', 'G')] == 1 && counts[gct] == 1 && sum == 2);
    }
    // scoring arithmetic (:840-935): spot values computed by hand from the reference's formulas
    KH_CHECK(length_score(57) == -20 && length_score(270) == (60 * 90) / 170 && length_score(3000) == (60 * 1000) / 1080);
    KH_CHECK(calibrated_start_term(13, 600) == 6 && calibrated_start_term(13, 300) == 8 && calibrated_start_term(13, 299) == 10);
    KH_CHECK(total_orf_score(270, true, 8, 90, false, true) == 31 + 90 + 8 - 8);
    {   // Shine-Dalgarno: AGGAG 7 nt upstream of the start scores 5 + 2
        std::vector<uint8_t> s = repeat("C", 40);
        const size_t start = 30;
        std::memcpy(s.data() + start - 7 - 5, "AGGAG", 5);
        KH_CHECK(sd_score(s, start) == 7);
        KH_CHECK(sd_score(s, 7) == 0);
    }
    {   // GcModel regressions at 50 % GC (f32): 0.575701 / 0.4000945 / 0.548953
        const GcModel m = gc_model_from(0.5f);
        KH_CHECK(m.exp1_pm == 576 && m.exp2_pm == 400 && m.exp3_pm == 549);
    }
    KH_CHECK(translate_table11(reinterpret_cast<const uint8_t*>("GTGAAATAGNNN"), 12, true) == "MK*X");
#undef KH_CHECK
    return std::string();
}

// predict_proteomes (:265-294): one `<name>.faa` per genome in tmp_dir, on `threads` host threads; the first
// error in input order is reported.
std::vector<SpeciesFile> predict_proteomes(const std::vector<SpeciesFile>& genomes, const std::string& tmp_dir, size_t threads) {
    std::vector<SpeciesFile> out(genomes.size());
    std::vector<std::string> errs(genomes.size());
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < genomes.size();) {
            out[i] = {genomes[i].name, tmp_dir + "/" + genomes[i].name + ".faa"};
            try { predict_one_genome(genomes[i].path, out[i].path); }
            catch (const std::exception& e) { errs[i] = *e.what() ? e.what() : "error"; }
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 0; t < std::max<size_t>(1, std::min(threads, genomes.size())); t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
    for (const std::string& e : errs)
        if (!e.empty()) throw std::runtime_error(e);
    return out;
}

}  // namespace kh
