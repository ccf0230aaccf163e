// kamino_host.hpp — host side of the B200 pipeline (C++ above the C ABI of include/kamino_b200.h).
//
// kamino's CLI, input discovery, extraction bookkeeping, group sorting / filtering and output writers
// stay host logic (north_star: "reference-equivalent host logic over the GPU-produced tables"); the
// data-parallel front end (recode, k-mers, Bloom, count-min table, occupancy scan, pair counts) is
// only ever computed by libkamino_b200.so — there is no host implementation of it in this tree.
// Names mirror the reference's modules; each function cites the reference lines it is equivalent to
// (paths under /root/reference).  Data structures are deliberately flat (sorted observation tables
// instead of hash maps) so the grouping can later move to the device.
#pragma once
#include "kamino_b200.h"

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <thread>
#include <vector>

namespace kh {

enum Scheme : int { Dayhoff6 = 0, SR6 = 1, KGB6 = 2 };

struct Options {                     // src/lib.rs:119-175
    std::string input_dir, input_file, output = "kamino";
    bool has_dir = false, has_file = false, genomes = false, nj = false;
    long k = -1, constant = -1;
    float min_freq = 0.85f;
    size_t length_middle = 35, mask = 5, threads = 1;
    Scheme recode = SR6;
    int device = 0;                  // new, non-breaking: CUDA ordinal (env KAMINO_DEVICE)
    std::vector<int> devices;        // new, non-breaking: --devices 0,1,... (env KAMINO_DEVICES): the species shard round-robin
                                     // over these GPUs, one table exchange per job; default: {device}
    size_t batch_bytes = size_t(1) << 30;   // residues per kmn_count_batch call
    bool quiet = false;
    std::string predict_only;        // new, hidden: with --genomes, write the predicted proteomes here and stop (no GPU)
};

struct SpeciesFile { std::string name, path; };      // src/io.rs:51-58

// The GPUs of one run: one kmn_ctx per device (rank r = devices[r]), created once and shared by the stage
// driver (both passes) and the --nj pair counts.  With several devices the contexts are wired for the
// in-process table exchange (kmn_exchange_setup_local).
struct GpuSet {
    std::vector<kmn_ctx*> ctx;
    std::vector<int> devices;
    GpuSet() = default;
    GpuSet(const GpuSet&) = delete;
    GpuSet& operator=(const GpuSet&) = delete;
    ~GpuSet();
    void create(const std::vector<int>& devices, int k, int scheme);
    size_t size() const { return ctx.size(); }
};

// io.rs
std::vector<SpeciesFile> inputs_from_directory(const std::string& dir);   // io.rs:16-79
std::vector<SpeciesFile> inputs_from_table(const std::string& tsv);       // io.rs:82-162
std::string species_name_of(const std::string& path);                      // io.rs:164-183

// One proteome as the front end wants it.  Host-parsed (load_proteome): `seq` = all rec.seq() bytes back to
// back.  Device-framed (load_fasta_bytes + index_records): `seq` = the file's bytes as read, the records
// located by kmn_count_fasta.  Either way record r's sequence is seq[rec_begin[r] .. rec_end[r]).
struct Proteome {
    std::vector<uint8_t> seq;
    std::vector<uint64_t> rec_begin, rec_end;
    std::vector<std::string> heads;    // header line without '>' / CR
    size_t n_rec() const { return rec_begin.size(); }
};
Proteome load_proteome(const std::string& path);      // io.rs:185-200 + seq_io::fasta::Reader
// The file's (gunzipped) bytes, a line feed appended if the last line lacks one (kmn_count_fasta's contract).
std::vector<uint8_t> load_fasta_bytes(const std::string& path);
// seq_io's parse of bytes already in memory; throws the reader's format error for `path`.
Proteome parse_proteome(const std::vector<uint8_t>& bytes, const std::string& path);
// Record table of a device-framed proteome: hdr[r] = offset of record r's '>' in p.seq (kmn_fetch_records).
void index_records(Proteome& p, const uint64_t* hdr, size_t n_rec, uint64_t base);

// protein_prediction.rs:265-294: `--genomes` — one predicted `<name>.faa` per nucleotide FASTA, written to tmp_dir.
std::vector<SpeciesFile> predict_proteomes(const std::vector<SpeciesFile>& genomes, const std::string& tmp_dir, size_t threads);

std::string extraction_self_test();   // group_extraction.rs protein-name / header checks; empty = passed
std::string host_logic_self_test();   // group_filtering.rs / group_sorting.rs unit tests on the host functions; empty = passed
std::string fasta_index_self_test();  // index_records vs parse_proteome on awkward FASTA texts; empty = passed
std::string prediction_self_test();   // the reference's unit tests of protein_prediction.rs; empty = all passed

// One observed block between an anchor pair (group_extraction.rs:118-127), flat form.
struct Observation {
    uint64_t left, right;              // anchor pair (group_extraction.rs:117)
    uint32_t sid;                      // species index
    uint32_t protein;                  // index into Extraction::protein_names
    uint64_t aa_off;                   // offset into Extraction::arena
    uint32_t len;                      // block length (<= u16::MAX enforced)
    uint64_t order;                    // emission order (stable tie-break == reference Vec order)
};

struct Extraction {                    // RawExtractedGroups, group_extraction.rs:510-525
    std::vector<std::string> species_names, protein_names;
    std::vector<uint8_t> arena;
    std::vector<Observation> obs;      // sorted by (left, right, order) once extraction is finished
    std::vector<kmn_group> groups;     // one run of obs per anchor pair, length selection done on the device
    size_t k = 0, min_needed = 0, n_species = 0, n_groups = 0;
};

struct Candidate {                     // RawDirectCandidate, group_sorting.rs:79-88
    uint64_t left, right;
    size_t coverage, raw_len;
    std::vector<uint32_t> rows;        // indices into Extraction::obs with the selected length
};

struct Alignment {                     // AlignmentResult, group_filtering.rs:12-21
    std::vector<std::string> species_names, partition_names;
    std::vector<std::vector<uint8_t>> concat;
    struct Part { size_t start, end, len; };
    std::vector<Part> partitions;
};

// group_extraction.rs:527-660 with both passes on the GPU (kmn_count_batch / kmn_scan_batch).
Extraction extract_groups(const std::vector<SpeciesFile>& inputs, size_t k, float min_freq, size_t length_middle,
                          Scheme scheme, const Options& opt, GpuSet& gpus);
// group_sorting.rs:329-380
std::vector<Candidate> sort_and_deduplicate_groups(const Extraction& ex, Scheme scheme);
// group_filtering.rs:292-363
Alignment filter_groups(const Extraction& ex, const std::vector<Candidate>& cands, float min_freq, size_t constant,
                        size_t mask, Scheme scheme, size_t threads);
// output.rs:32-95 (+ phylo.rs:274-301 with the pair counts from kmn_pair_counts)
void write_outputs(const std::string& out_base, const Alignment& al, bool nj, GpuSet& gpus, size_t threads, size_t* total_len,
                   double* pct_missing);
// lib.rs:206-298
void run(const Options& opt);
int cli(int argc, char** argv);

uint8_t recode(uint8_t b, Scheme s);

// std::sort on `threads` host threads: sorted chunks, then pairwise in-place merges (same result as one
// std::sort for a strict weak order whose ties are broken by a unique key).
template <typename T, typename Cmp>
void parallel_sort(std::vector<T>& v, Cmp cmp, size_t threads) {
    const size_t n = v.size();
    size_t parts = 1;
    while (parts * 2 <= threads && n / (parts * 2) >= (size_t(1) << 15)) parts *= 2;
    if (parts == 1) { std::sort(v.begin(), v.end(), cmp); return; }
    std::vector<size_t> cut(parts + 1);
    for (size_t i = 0; i <= parts; i++) cut[i] = n * i / parts;
    {
        std::vector<std::thread> pool;
        for (size_t i = 0; i < parts; i++)
            pool.emplace_back([&, i] { std::sort(v.begin() + cut[i], v.begin() + cut[i + 1], cmp); });
        for (auto& t : pool) t.join();
    }
    for (size_t width = 1; width < parts; width *= 2) {
        std::vector<std::thread> pool;
        for (size_t i = 0; i + width < parts; i += 2 * width)
            pool.emplace_back([&, i] {
                std::inplace_merge(v.begin() + cut[i], v.begin() + cut[i + width], v.begin() + cut[std::min(parts, i + 2 * width)], cmp);
            });
        for (auto& t : pool) t.join();
    }
}

// Rust's str::trim / trim_end / split_whitespace / char::is_whitespace follow the Unicode White_Space property
// (group_extraction.rs:398-402,429-430; group_filtering.rs:153-155,179; io.rs:122-123), not just ASCII.  On valid
// UTF-8 the white-space characters are exactly these byte sequences:
//   09..0D, 20 | C2 85, C2 A0 | E1 9A 80 | E2 80 80..8A, E2 80 A8, E2 80 A9, E2 80 AF | E2 81 9F | E3 80 80
namespace uws {
// length in bytes of the white-space character that STARTS at s[i], 0 if none does
inline size_t at(const std::string& s, size_t i) {
    const size_t n = s.size();
    const unsigned char c = (unsigned char)s[i];
    if (c == 0x20 || (c >= 0x09 && c <= 0x0D)) return 1;
    if (c == 0xC2 && i + 1 < n) {
        const unsigned char d = (unsigned char)s[i + 1];
        return (d == 0x85 || d == 0xA0) ? 2 : 0;
    }
    if (i + 2 >= n) return 0;
    const unsigned char d = (unsigned char)s[i + 1], e = (unsigned char)s[i + 2];
    if (c == 0xE1) return (d == 0x9A && e == 0x80) ? 3 : 0;
    if (c == 0xE2) {
        if (d == 0x80) return ((e >= 0x80 && e <= 0x8A) || e == 0xA8 || e == 0xA9 || e == 0xAF) ? 3 : 0;
        return (d == 0x81 && e == 0x9F) ? 3 : 0;
    }
    if (c == 0xE3) return (d == 0x80 && e == 0x80) ? 3 : 0;
    return 0;
}
// length in bytes of the white-space character that ENDS just before s[e], 0 if none does
inline size_t before(const std::string& s, size_t e) {
    for (size_t len = 1; len <= 3 && len <= e; len++)
        if (at(s, e - len) == len) return len;
    return 0;
}
inline std::string trim_end(const std::string& s) {
    size_t e = s.size();
    for (size_t l; e > 0 && (l = before(s, e)) != 0;) e -= l;
    return s.substr(0, e);
}
inline std::string trim(const std::string& s) {
    size_t b = 0, e = s.size();
    for (size_t l; b < e && (l = at(s, b)) != 0;) b += l;
    for (size_t l; e > b && (l = before(s, e)) != 0;) e -= l;
    return s.substr(b, e - b);
}
}  // namespace uws

// Wall-clock stage timer: prints "[time] <what>: <s>" to stderr when KAMINO_TIMING is set.
struct StageTimer {
    const char* what;
    double t0;
    explicit StageTimer(const char* w);
    ~StageTimer();
};   // recode.rs:31-80 (host copy for the column filters of group_filtering.rs)

}  // namespace kh
