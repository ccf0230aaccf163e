// kamino_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A plain C++17 restatement of rderelle/kamino 2.0.0 (Rust) for the data-parallel
// front end and everything downstream of it that the reference's golden tests pin.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product (kamino_b200/) never links or calls it.
//
// Parity status: the reference cannot be compiled here (no cargo/rustc in the image),
// so there is no oracle/_ref.  The oracle is PINNED by
//   * the reference's own end-to-end goldens tests/data/test_3diff (dir + TSV case,
//     Dayhoff6 + KGB6, alignment / missing / partitions, plus the checked-in NJ tree),
//     copied as data fixtures under tests/golden/test_3diff;
//   * the derived known-answer vectors of SURVEY.md §8c (mix64, first 13-mers,
//     Bloom/CMS indices, 3diff CMS histogram, start/end anchors);
//   * the reference's own unit tests of the host logic (group_extraction.rs:666-748,
//     group_sorting.rs:438-659, group_filtering.rs:413-693), restated in selftest.cpp;
//   * for --genomes (protein_prediction.py, plain Python): the golden tests/data/test_genomes
//     and the module's unit tests (protein_prediction.rs:1152-1286).
// Collision / saturation / Bloom-order behaviour is not covered by any reference test
// ("parity unpinned by the reference's own tests"); it follows the source literally.
//
// Every function cites the reference file:line (paths relative to /root/reference).
#pragma once
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include <unordered_map>
#include <stdexcept>

namespace kmo {

// char::is_whitespace / str::trim / split_whitespace of Rust's std follow the Unicode White_Space property.
// Decodes the scalar value that starts at s[i] (the callers hold valid UTF-8) and tests the property's ranges;
// returns the character's byte length when it is white space, else 0.
inline size_t rs_ws_len(const std::string& s, size_t i) {
    const unsigned char b0 = (unsigned char)s[i];
    uint32_t cp;
    size_t len;
    if (b0 < 0x80) { cp = b0; len = 1; }
    else if ((b0 >> 5) == 0x6 && i + 1 < s.size()) { cp = ((b0 & 0x1Fu) << 6) | ((unsigned char)s[i + 1] & 0x3Fu); len = 2; }
    else if ((b0 >> 4) == 0xE && i + 2 < s.size()) {
        cp = ((b0 & 0x0Fu) << 12) | (((unsigned char)s[i + 1] & 0x3Fu) << 6) | ((unsigned char)s[i + 2] & 0x3Fu);
        len = 3;
    } else return 0;   // 4-byte characters: none is white space
    const bool ws = (cp >= 0x09 && cp <= 0x0D) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 ||
                    (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
    return ws ? len : 0;
}
inline std::string rs_trim_start(const std::string& s) {
    size_t b = 0;
    while (b < s.size()) {
        const size_t l = rs_ws_len(s, b);
        if (!l) break;
        b += l;
    }
    return s.substr(b);
}
inline std::string rs_trim_end(const std::string& s) {
    size_t e = s.size();
    for (;;) {
        // the last character starts at the last byte that is not a UTF-8 continuation byte
        size_t st = e;
        while (st > 0 && (((unsigned char)s[st - 1]) & 0xC0) == 0x80) st--;
        if (st == 0) break;
        st--;
        if (rs_ws_len(s, st) != e - st || e == st) break;
        e = st;
    }
    return s.substr(0, e);
}
inline std::string rs_trim(const std::string& s) { return rs_trim_end(rs_trim_start(s)); }


// ---------------------------------------------------------------- recode.rs
constexpr uint32_t RECODE_BITS_PER_SYMBOL = 3;          // src/recode.rs:5
enum RecodeScheme : int { Dayhoff6 = 0, SR6 = 1, KGB6 = 2 };  // src/recode.rs:9-16
uint8_t recode_byte(uint8_t b, RecodeScheme s);         // src/recode.rs:31-37
const char* scheme_name(RecodeScheme s);                // src/recode.rs:18-28

// ---------------------------------------------------------- proba_filter.rs
constexpr size_t BLOOM_BITS = size_t(1) << 25;          // src/proba_filter.rs:13
constexpr size_t BLOOM_HASHES = 2;                      // src/proba_filter.rs:15
constexpr size_t CMS_WIDTH = size_t(1) << 26;           // src/proba_filter.rs:17
constexpr size_t CMS_DEPTH = 4;                         // src/proba_filter.rs:19
extern const uint64_t SEEDS[6];                         // src/proba_filter.rs:21-28
uint64_t mix64(uint64_t x);                             // src/proba_filter.rs:30-37

struct BloomFilter {                                    // src/proba_filter.rs:40-75
    std::vector<uint64_t> bits;
    uint64_t mask;
    explicit BloomFilter(size_t nbits);
    bool test_and_set(uint64_t key);
    void clear();
};

struct CountMinSketch {                                 // src/proba_filter.rs:78-113
    size_t width;
    std::vector<uint16_t> table;
    std::vector<uint64_t> seeds;
    CountMinSketch(size_t width, size_t depth);
    uint32_t estimate(uint64_t key) const;
};

struct AtomicCountMinSketch {                           // src/proba_filter.rs:119-178
    size_t width;
    size_t cells;
    uint16_t* table;     // accessed with __atomic builtins (AtomicU16 semantics)
    std::vector<uint64_t> seeds;
    AtomicCountMinSketch(size_t width, size_t depth);
    ~AtomicCountMinSketch();
    AtomicCountMinSketch(const AtomicCountMinSketch&) = delete;
    AtomicCountMinSketch& operator=(const AtomicCountMinSketch&) = delete;
    void increment(uint64_t key);
    CountMinSketch snapshot() const;
};

// -------------------------------------------------------------------- io.rs
struct SpeciesInput { std::string name; std::string path; };   // src/io.rs:51-58
std::vector<std::string> collect_fasta_files(const std::string& dir);             // io.rs:16-49
std::vector<SpeciesInput> collect_species_inputs_from_dir(const std::string& d);  // io.rs:61-79
std::vector<SpeciesInput> collect_species_inputs_from_table(const std::string& t);// io.rs:82-162
std::string species_name_from_path(const std::string& p);                          // io.rs:164-183

// One FASTA record as seq_io 0.3 exposes it: head = header line without '>' and
// without the trailing CR/LF; seq = raw bytes of the following lines up to the next
// line that starts with '>' (line terminators INCLUDED, kamino skips them itself).
struct FastaRecord { std::string head; std::string seq; };
// open_fasta (io.rs:185-200) + seq_io::fasta::Reader: gunzip iff the path ends ".gz".
std::vector<FastaRecord> read_fasta(const std::string& path);
// Whole (decompressed) file as bytes; used by the staged-input benchmark helpers.
std::string read_all_maybe_gz(const std::string& path);
std::vector<FastaRecord> parse_fasta_bytes(const std::string& bytes, const std::string& what);

// ------------------------------------------------------- group_extraction.rs
constexpr size_t START_CLUSTER_SPAN = 10;               // group_extraction.rs:21
struct BlockAa { size_t start; uint16_t len; };         // group_extraction.rs:67-78
struct BlockArena {                                     // group_extraction.rs:80-114
    std::vector<uint8_t> data;
    BlockAa push(const uint8_t* p, size_t n);
    size_t append(const BlockArena& o);
    const uint8_t* get(BlockAa b) const { return data.data() + b.start; }
};
struct AnchorPair {
    uint64_t left, right;
    bool operator==(const AnchorPair& o) const { return left == o.left && right == o.right; }
};
struct AnchorPairHash {
    size_t operator()(const AnchorPair& p) const {
        uint64_t h = p.left * 0x9e3779b97f4a7c15ull ^ (p.right + 0x7f4a7c15ull + (p.left << 6));
        h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
        return (size_t)h;
    }
};
struct BlockObs { uint16_t sid; uint32_t protein_id; BlockAa aa; };   // :118-127
using RawDirectGroups = std::unordered_map<AnchorPair, std::vector<BlockObs>, AnchorPairHash>;
struct KmerHit { uint64_t kmer; size_t start, end; uint32_t occupancy; bool shared; }; // :131-139

std::vector<KmerHit> select_best_starts_per_local_cluster(const std::vector<KmerHit>& starts); // :151-175
std::vector<std::pair<size_t, size_t>> pair_anchors(const std::vector<KmerHit>& start_anchors,
                                                    const std::vector<KmerHit>& end_anchors, size_t length_middle);
void extract_bubble_pairs(const uint8_t* aa, const std::vector<KmerHit>& starts,
                          const std::vector<KmerHit>& ends, size_t length_middle, uint16_t sid,
                          uint32_t protein_id, RawDirectGroups& groups, BlockArena& arena);   // :178-242
// Boundary rule only (group_extraction.rs:264-303): fills start candidates / end anchors
// and (optionally) the occupancy of every k-mer position of the segment.
void scan_valid_segment(const uint8_t* recoded, size_t len, size_t k, uint64_t k_mask,
                        const CountMinSketch& cms, uint32_t min_needed,
                        std::vector<KmerHit>& start_candidates, std::vector<KmerHit>& end_anchors,
                        std::vector<uint16_t>* occ_out);
void extract_from_valid_segment(const uint8_t* aa, const uint8_t* recoded, size_t len, size_t k,
                                uint64_t k_mask, const CountMinSketch& cms, uint32_t min_needed,
                                size_t length_middle, uint16_t sid, uint32_t protein_id,
                                RawDirectGroups& groups, BlockArena& arena);                   // :245-322
struct SpeciesExtraction {                                                                     // :324-333
    size_t sid; BlockArena arena; RawDirectGroups groups; std::vector<std::string> protein_names;
};
std::string truncate_protein_name(const std::string& raw);                                    // :397-404
void merge_species_extraction(SpeciesExtraction&& ex, RawDirectGroups& groups, BlockArena& arena,
                              std::vector<std::string>& protein_names);                       // :335-357
// Pass 1 for one species over already-parsed records (group_extraction.rs:359-395).
// Returns the number of k-mers that passed the Bloom filter (CMS increments / CMS_DEPTH).
uint64_t count_species_records(const std::vector<FastaRecord>& recs, size_t k, uint64_t k_mask,
                               RecodeScheme scheme, AtomicCountMinSketch& cmsa, BloomFilter& bloom);
SpeciesExtraction extract_species_records(size_t sid, const std::string& species_name,
                                          const std::vector<FastaRecord>& recs, size_t k,
                                          uint64_t k_mask, const CountMinSketch& cms,
                                          uint32_t min_needed, size_t length_middle,
                                          RecodeScheme scheme);                                 // :407-507
struct RawExtractedGroups {                                                                    // :510-525
    std::vector<std::string> species_names; BlockArena arena; RawDirectGroups groups;
    std::vector<std::string> protein_names; size_t k = 0, min_needed = 0, n_species = 0;
};
uint32_t min_needed_for(float min_freq, size_t n);                                             // :544
uint64_t k_mask_for(size_t k);                                                                 // :545-549
RawExtractedGroups extract_groups(const std::vector<SpeciesInput>& inputs, size_t k, float min_freq,
                                  size_t length_middle, RecodeScheme scheme, size_t threads);  // :527-660

// ---------------------------------------------------------- group_sorting.rs
struct RawDirectCandidate {                                                // group_sorting.rs:79-88
    AnchorPair key; size_t coverage; size_t raw_len; std::vector<BlockObs> selected;
};
size_t direct_overlap_k(size_t k);                                         // :12-22
// returns false when rejected (tie / low support / short block)            // :90-143
bool select_unique_best_supported_length(AnchorPair key, const std::vector<BlockObs>& obs,
                                         size_t n_species, size_t min_needed, size_t k,
                                         RawDirectCandidate& out);
std::vector<RawDirectCandidate> retain_nonoverlapping_raw_candidates(
    std::vector<RawDirectCandidate> cands, const BlockArena& arena, size_t overlap_k,
    RecodeScheme scheme);                                                  // :246-286
std::vector<std::vector<uint8_t>> consensus_rows_by_isolate(const std::vector<BlockObs>& obs,
                                                            const BlockArena& arena,
                                                            size_t n_species, size_t len); // :288-308
struct SortedGroups {                                                      // :310-327
    std::vector<std::string> species_names; BlockArena arena;
    std::vector<RawDirectCandidate> raw_candidates; std::vector<std::string> protein_names;
    size_t k = 0, min_needed = 0, n_species = 0; RecodeScheme scheme = SR6;
};
SortedGroups sort_and_deduplicate_groups(RawExtractedGroups&& raw, RecodeScheme scheme); // :329-380

// -------------------------------------------------------- group_filtering.rs
struct AlignmentResult {                                                   // group_filtering.rs:12-21
    std::vector<std::string> species_names; std::vector<std::vector<uint8_t>> concat;
    struct Part { size_t start, end, len; };
    std::vector<Part> partitions; std::vector<std::string> partition_names;
};
std::vector<int> majority_recoded_consensus(const std::vector<std::vector<uint8_t>>& rows,
                                            size_t len, RecodeScheme s);   // :37-67 (-1 = None)
void mask_rows_with_long_divergence(std::vector<std::vector<uint8_t>>& rows,
                                    const std::vector<int>& consensus, size_t mask,
                                    RecodeScheme s);                       // :69-113
bool is_polymorphic_column(const std::vector<std::vector<uint8_t>>& rows, size_t col,
                           RecodeScheme s);                                // :115-133
float missing_fraction(const std::vector<std::vector<uint8_t>>& rows, size_t col, size_t n); // :135-141
std::string consensus_protein_name(const std::vector<BlockObs>& obs,
                                   const std::vector<std::string>& protein_names); // :143-221
AlignmentResult filter_groups(SortedGroups&& sorted, float min_freq, size_t constant, size_t mask,
                              size_t threads);                             // :292-363

// ------------------------------------------------------------------ phylo.rs
extern const double LG_PI[20];                                             // phylo.rs:27-31
double f81_constant();                                                     // phylo.rs:34-43
uint8_t aa_index_or_20(uint8_t b);                                         // phylo.rs:47-79
void pair_counts(const uint8_t* a, const uint8_t* b, size_t len, uint64_t& valid, uint64_t& mism); // :84-95
double lg_f81_from_counts(uint64_t valid, uint64_t mism);                  // phylo.rs:96-105
std::vector<double> compute_distance_matrix(const std::vector<std::vector<uint8_t>>& mapped,
                                            size_t threads);               // phylo.rs:114-141
std::string nj_tree_newick(const std::vector<std::string>& names,
                           const std::vector<std::vector<uint8_t>>& seqs, size_t threads); // :274-301
// NJ + Newick from an externally supplied (2n-1)^2 matrix (phylo.rs:145-272).
std::string nj_newick_from_matrix(const std::vector<std::string>& names, std::vector<double>& dist);

// ----------------------------------------------------------------- output.rs
struct OutputPaths { std::string fas, tsv, parts, tree; };
OutputPaths output_paths(const std::string& out_base);                     // output.rs:9-30
void write_outputs(const std::string& out_base, const AlignmentResult& res, bool nj, size_t threads,
                   size_t* total_len, double* pct_missing);                // output.rs:32-95

// -------------------------------------------------------------------- lib.rs
struct Args {                                                              // lib.rs:119-175
    std::string input, input_file; bool has_input = false, has_input_file = false;
    bool genomes = false; long k = -1; float min_freq = 0.85f; std::string output = "kamino";
    long constant = -1; size_t length_middle = 35; size_t mask = 5; size_t threads = 1;
    RecodeScheme recode = SR6; bool nj = false;
};
void run_with_args(const Args& a);                                         // lib.rs:206-298
int cli_main(int argc, char** argv);                                       // main.rs:5-9 + clap

// Rust std::path helpers the io/output code depends on.
std::string rs_file_name(const std::string& p);
std::string rs_parent(const std::string& p, bool* has_parent);
std::string rs_extension(const std::string& file_name);   // "" when none
std::string rs_file_stem(const std::string& file_name);

// selftest.cpp: the reference's unit tests of the host logic on these functions; "" = all passed,
// else the name of the first failing reference test
std::string self_test();

}  // namespace kmo
