// capi.cpp — CPU ORACLE (test infrastructure): extern "C" surface for ctypes (tests/,
// __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference ONLY).
// The batch layout mirrors the product ABI (include/kamino_b200.h) so that the parity tests
// feed byte-identical inputs to both sides.
#include "kamino_oracle.hpp"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

using namespace kmo;

static thread_local std::string g_err;
static int fail(const std::exception& e) { g_err = e.what(); return 1; }

struct orc_hit {  // == kmn_hit
    uint64_t kmer;
    uint32_t rec;        // record index within the species
    uint32_t seg_start;  // whitespace-stripped offset of the valid segment inside the record
    uint32_t start;      // k-mer start inside the segment (KmerHit.start)
    uint32_t occ;        // CountMinSketch::estimate
};

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

uint8_t orc_recode_byte(uint8_t b, int scheme) { return recode_byte(b, RecodeScheme(scheme)); }
uint64_t orc_mix64(uint64_t x) { return mix64(x); }
uint64_t orc_k_mask(int k) { return k_mask_for(size_t(k)); }
uint32_t orc_min_needed(float min_freq, uint64_t n) { return min_needed_for(min_freq, size_t(n)); }

// First rolling k-mers of a residue string (group_extraction.rs:373-388); returns count written.
uint64_t orc_roll_kmers(const uint8_t* seq, uint64_t len, int k, int scheme, uint64_t* out, uint64_t cap) {
    uint64_t roll = 0, n = 0, mask = k_mask_for(size_t(k));
    size_t have = 0;
    for (uint64_t i = 0; i < len; i++) {
        uint8_t b = seq[i];
        if (b == 0x20 || b == 0x09 || b == 0x0A || b == 0x0C || b == 0x0D) continue;
        uint8_t up = (b >= 'a' && b <= 'z') ? uint8_t(b - 32) : b;
        uint8_t a = recode_byte(up, RecodeScheme(scheme));
        if (a == 255) { roll = 0; have = 0; continue; }
        roll = ((roll << 3) | a) & mask;
        if (have < size_t(k)) have++;
        if (have == size_t(k)) { if (n < cap) out[n] = roll; n++; }
    }
    return n;
}

void* orc_bloom_new(uint64_t bits) { return new BloomFilter(size_t(bits)); }
int orc_bloom_test_and_set(void* b, uint64_t key) { return static_cast<BloomFilter*>(b)->test_and_set(key) ? 1 : 0; }
void orc_bloom_free(void* b) { delete static_cast<BloomFilter*>(b); }
// Bloom probe indices (proba_filter.rs:60-64) for known-answer tests.
void orc_bloom_indices(uint64_t key, uint64_t* idx0, uint64_t* idx1) {
    uint64_t h1 = mix64(key ^ SEEDS[0]), h2 = mix64(key ^ SEEDS[1]) | 1;
    uint64_t mask = BLOOM_BITS - 1;
    *idx0 = h1 & mask;
    *idx1 = (h1 + h2) & mask;
}
void orc_cms_indices(uint64_t key, uint64_t* idx4) {  // proba_filter.rs:107,148
    for (int r = 0; r < 4; r++) idx4[r] = mix64(key ^ SEEDS[2 + r]) & (CMS_WIDTH - 1);
}

void* orc_cmsa_new() {
    try { return new AtomicCountMinSketch(CMS_WIDTH, CMS_DEPTH); } catch (const std::exception& e) { fail(e); return nullptr; }
}
void orc_cmsa_free(void* c) { delete static_cast<AtomicCountMinSketch*>(c); }
void orc_cmsa_increment(void* c, uint64_t key) { static_cast<AtomicCountMinSketch*>(c)->increment(key); }
void orc_cmsa_reset(void* c) {
    auto* a = static_cast<AtomicCountMinSketch*>(c);
    std::memset(a->table, 0, a->cells * sizeof(uint16_t));
}
// Raw view of the atomic table (== snapshot contents once counting has finished).
const uint16_t* orc_cmsa_table(void* c) { return static_cast<AtomicCountMinSketch*>(c)->table; }
uint64_t orc_cmsa_cells(void* c) { return static_cast<AtomicCountMinSketch*>(c)->cells; }
void* orc_cmsa_snapshot(void* c) {
    try { return new CountMinSketch(static_cast<AtomicCountMinSketch*>(c)->snapshot()); }
    catch (const std::exception& e) { fail(e); return nullptr; }
}
void* orc_cms_from_table(const uint16_t* table) {
    try {
        auto* c = new CountMinSketch(CMS_WIDTH, CMS_DEPTH);
        std::memcpy(c->table.data(), table, c->table.size() * sizeof(uint16_t));
        return c;
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
const uint16_t* orc_cms_table(void* c) { return static_cast<CountMinSketch*>(c)->table.data(); }
uint32_t orc_cms_estimate(void* c, uint64_t key) { return static_cast<CountMinSketch*>(c)->estimate(key); }
void orc_cms_free(void* c) { delete static_cast<CountMinSketch*>(c); }

// Pass 1 over a staged batch: species-parallel threads, shared atomic sketch, one private
// Bloom per species (group_extraction.rs:557-567 + :359-395).
int orc_count_batch(void* cmsa_, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                    const uint64_t* sp_rec_off, uint32_t n_sp, int k, int scheme, int threads,
                    uint64_t* new_kmers_per_species) {
    try {
        (void)n_rec;
        auto* cmsa = static_cast<AtomicCountMinSketch*>(cmsa_);
        uint64_t mask = k_mask_for(size_t(k));
        std::atomic<uint32_t> next{0};
        auto worker = [&]() {
            BloomFilter bloom(BLOOM_BITS);
            for (;;) {
                uint32_t s = next.fetch_add(1);
                if (s >= n_sp) break;
                bloom.clear();
                // same loop as count_species_records, run straight over the staged bytes
                uint64_t n_new = 0;
                for (uint64_t r = sp_rec_off[s]; r < sp_rec_off[s + 1]; r++) {
                    uint64_t roll = 0;
                    size_t have = 0;
                    for (uint64_t i = rec_off[r]; i < rec_off[r + 1]; i++) {
                        uint8_t b = residues[i];
                        if (b == 0x20 || b == 0x09 || b == 0x0A || b == 0x0C || b == 0x0D) continue;
                        uint8_t up = (b >= 'a' && b <= 'z') ? uint8_t(b - 32) : b;
                        uint8_t a = recode_byte(up, RecodeScheme(scheme));
                        if (a == 255) { roll = 0; have = 0; continue; }
                        roll = ((roll << RECODE_BITS_PER_SYMBOL) | uint64_t(a)) & mask;
                        if (have < size_t(k)) have++;
                        if (have == size_t(k) && !bloom.test_and_set(roll)) { cmsa->increment(roll); n_new++; }
                    }
                }
                if (new_kmers_per_species) new_kmers_per_species[s] = n_new;
            }
        };
        int nt = threads < 1 ? 1 : threads;
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// Pass 2 boundary scan for one species of a staged batch (group_extraction.rs:446-498 segmentation
// + :264-303 boundary rule).  occ_per_residue (optional): one u16 per whitespace-stripped residue of
// the species (records concatenated); estimate() of the k-mer ENDING there under pass-1 validity
// (have == k), 0 where no k-mer ends.  Hit lists only cover segments >= 2k+1 like the reference.
int orc_scan_species(void* cms_, const uint8_t* residues, const uint64_t* rec_off, uint64_t rec_begin,
                     uint64_t rec_end, int k_, int scheme, uint32_t min_needed, orc_hit* starts,
                     uint64_t cap_starts, uint64_t* n_starts, orc_hit* ends, uint64_t cap_ends,
                     uint64_t* n_ends, uint16_t* occ_per_residue) {
    try {
        auto* cms = static_cast<CountMinSketch*>(cms_);
        size_t k = size_t(k_);
        uint64_t mask = k_mask_for(k);
        uint64_t ns = 0, ne = 0, stripped_base = 0;
        std::vector<uint8_t> recoded;
        std::vector<KmerHit> sc, ea;
        for (uint64_t r = rec_begin; r < rec_end; r++) {
            uint32_t rec = uint32_t(r - rec_begin);
            recoded.clear();
            uint32_t off_in_rec = 0, seg_start = 0;
            auto flush = [&]() {
                if (recoded.size() >= 2 * k + 1 && k > 0) {
                    sc.clear(); ea.clear();
                    scan_valid_segment(recoded.data(), recoded.size(), k, mask, *cms, min_needed, sc, ea, nullptr);
                    for (auto& h : sc) { if (ns < cap_starts) starts[ns] = orc_hit{h.kmer, rec, seg_start, uint32_t(h.start), h.occupancy}; ns++; }
                    for (auto& h : ea) { if (ne < cap_ends) ends[ne] = orc_hit{h.kmer, rec, seg_start, uint32_t(h.start), h.occupancy}; ne++; }
                }
            };
            uint64_t roll = 0; size_t have = 0;
            for (uint64_t i = rec_off[r]; i < rec_off[r + 1]; i++) {
                uint8_t b = residues[i];
                if (b == 0x20 || b == 0x09 || b == 0x0A || b == 0x0C || b == 0x0D) continue;
                uint8_t up = (b >= 'a' && b <= 'z') ? uint8_t(b - 32) : b;
                uint8_t rv = recode_byte(up, RecodeScheme(scheme));
                if (rv == 255) {
                    flush();
                    recoded.clear();
                    seg_start = off_in_rec + 1;
                    roll = 0; have = 0;
                    if (occ_per_residue) occ_per_residue[stripped_base + off_in_rec] = 0;
                } else {
                    recoded.push_back(rv);
                    roll = ((roll << 3) | rv) & mask;
                    if (have < k) have++;
                    if (occ_per_residue)
                        occ_per_residue[stripped_base + off_in_rec] = have == k ? uint16_t(cms->estimate(roll)) : 0;
                }
                off_in_rec++;
            }
            flush();
            stripped_base += off_in_rec;
        }
        *n_starts = ns;
        *n_ends = ne;
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// Anchor pairs of one species (group_extraction.rs:245-322 without the arena): for every valid segment
// scan_valid_segment -> select_best_starts_per_local_cluster -> pair_anchors, in reference order.
struct orc_pair { uint64_t left, right; uint32_t rec, seg_start, left_start, len; };   // == kmn_pair
int orc_pair_species(void* cms_, const uint8_t* residues, const uint64_t* rec_off, uint64_t rec_begin,
                     uint64_t rec_end, int k_, int scheme, uint32_t min_needed, uint64_t length_middle,
                     orc_pair* out, uint64_t cap, uint64_t* n_out) {
    try {
        auto* cms = static_cast<CountMinSketch*>(cms_);
        size_t k = size_t(k_);
        uint64_t mask = k_mask_for(k);
        uint64_t n = 0;
        std::vector<uint8_t> recoded;
        std::vector<KmerHit> sc, ea;
        for (uint64_t r = rec_begin; r < rec_end; r++) {
            uint32_t rec = uint32_t(r - rec_begin);
            recoded.clear();
            uint32_t off_in_rec = 0, seg_start = 0;
            auto flush = [&]() {
                if (recoded.size() >= 2 * k + 1 && k > 0) {
                    sc.clear(); ea.clear();
                    scan_valid_segment(recoded.data(), recoded.size(), k, mask, *cms, min_needed, sc, ea, nullptr);
                    std::vector<KmerHit> sel = select_best_starts_per_local_cluster(sc);
                    for (const auto& pr : pair_anchors(sel, ea, size_t(length_middle))) {
                        const KmerHit& left = sel[pr.first];
                        const KmerHit& right = ea[pr.second];
                        if (n < cap) out[n] = orc_pair{left.kmer, right.kmer, rec, seg_start, uint32_t(left.start),
                                                       uint32_t(right.end - left.start)};
                        n++;
                    }
                }
            };
            for (uint64_t i = rec_off[r]; i < rec_off[r + 1]; i++) {
                uint8_t b = residues[i];
                if (b == 0x20 || b == 0x09 || b == 0x0A || b == 0x0C || b == 0x0D) continue;
                uint8_t up = (b >= 'a' && b <= 'z') ? uint8_t(b - 32) : b;
                uint8_t rv = recode_byte(up, RecodeScheme(scheme));
                if (rv == 255) {
                    flush();
                    recoded.clear();
                    seg_start = off_in_rec + 1;
                } else {
                    recoded.push_back(rv);
                }
                off_in_rec++;
            }
            flush();
        }
        *n_out = n;
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// --nj integer counts (phylo.rs:84-95) for all pairs i<j; valid/mismatch are n*n row-major,
// only the strict upper triangle is written.
int orc_pair_counts(const uint8_t* mapped, uint32_t n, uint64_t L, uint32_t* valid, uint32_t* mismatch, int threads) {
    try {
        std::atomic<uint32_t> next{0};
        auto worker = [&]() {
            for (;;) {
                uint32_t i = next.fetch_add(1);
                if (i >= n) break;
                for (uint32_t j = i + 1; j < n; j++) {
                    uint64_t v, m;
                    pair_counts(mapped + uint64_t(i) * L, mapped + uint64_t(j) * L, size_t(L), v, m);
                    valid[uint64_t(i) * n + j] = uint32_t(v);
                    mismatch[uint64_t(i) * n + j] = uint32_t(m);
                }
            }
        };
        int nt = threads < 1 ? 1 : threads;
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
uint8_t orc_aa_index(uint8_t b) { return aa_index_or_20(b); }
double orc_f81_constant() { return f81_constant(); }
double orc_lg_f81_from_counts(uint64_t valid, uint64_t mism) { return lg_f81_from_counts(valid, mism); }
// NJ + Newick from an n x n distance matrix given as counts-derived f64 (copied into (2n-1)^2).
int orc_nj_newick(const char* const* names, uint32_t n, const double* dist_nxn, char* out, uint64_t cap) {
    try {
        std::vector<std::string> nm(names, names + n);
        size_t dim = 2 * size_t(n) - 1;
        std::vector<double> d(dim * dim, 0.0);
        for (uint32_t i = 0; i < n; i++)
            for (uint32_t j = 0; j < n; j++) d[i * dim + j] = dist_nxn[size_t(i) * n + j];
        std::string s = nj_newick_from_matrix(nm, d);
        if (s.size() + 1 > cap) throw std::runtime_error("newick buffer too small");
        std::memcpy(out, s.c_str(), s.size() + 1);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// merge_species_extraction (group_extraction.rs:335-357) + select_unique_best_supported_length
// (group_sorting.rs:90-143) on n observations given in emission order (species-major): the checker of the
// device grouping (kmn_group_observations).  Groups come out in ascending (left, right); per group:
// observations, accepted (0/1), best length and coverage (only meaningful when accepted).
// sel[] (n entries) lists, group after group, the emission indices of the SELECTED observations of the
// accepted groups; n_sel = how many.
int orc_group_observations(uint64_t n, const uint64_t* left, const uint64_t* right, const uint32_t* sid,
                           const uint32_t* len, uint64_t n_species, uint32_t min_needed, int k,
                           uint64_t* g_left, uint64_t* g_right, uint32_t* g_count, uint32_t* g_ok,
                           uint32_t* g_len, uint32_t* g_cov, uint64_t cap_groups, uint64_t* n_groups,
                           uint64_t* sel, uint64_t* n_sel) {
    try {
        RawDirectGroups groups;                       // the reference's HashMap<(u64,u64), Vec<BlockObs>>
        for (uint64_t i = 0; i < n; i++) {
            if (len[i] > 65535) throw std::runtime_error("block length exceeds u16");   // group_extraction.rs:92-94
            BlockObs o{uint16_t(sid[i]), 0u, BlockAa{size_t(i), uint16_t(len[i])}};      // aa.start carries the emission index
            groups[AnchorPair{left[i], right[i]}].push_back(o);
        }
        std::vector<AnchorPair> keys;
        keys.reserve(groups.size());
        for (auto& kv : groups) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end(), [](const AnchorPair& a, const AnchorPair& b) {
            return a.left != b.left ? a.left < b.left : a.right < b.right;
        });
        if (keys.size() > cap_groups) throw std::runtime_error("orc_group_observations: group buffers too small");
        uint64_t ns = 0;
        for (size_t g = 0; g < keys.size(); g++) {
            const auto& obs = groups[keys[g]];
            RawDirectCandidate c;
            const bool ok = select_unique_best_supported_length(keys[g], obs, size_t(n_species), size_t(min_needed),
                                                                size_t(k), c);
            g_left[g] = keys[g].left;
            g_right[g] = keys[g].right;
            g_count[g] = uint32_t(obs.size());
            g_ok[g] = ok ? 1u : 0u;
            g_len[g] = ok ? uint32_t(c.raw_len) : 0u;
            g_cov[g] = ok ? uint32_t(c.coverage) : 0u;
            if (ok)
                for (const BlockObs& o : c.selected) sel[ns++] = uint64_t(o.aa.start);
        }
        *n_groups = keys.size();
        *n_sel = ns;
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// Full pipeline with the reference CLI (argv[0] ignored).
int orc_run(int argc, char** argv) { return cli_main(argc, argv); }

// FASTA framing helper for tests/bench: parse a (possibly .gz) file with seq_io semantics and
// return the concatenated seq() bytes + record offsets.  Two-call protocol: pass null buffers
// to get sizes.
int orc_read_fasta(const char* path, uint8_t* residues, uint64_t cap_res, uint64_t* n_res,
                   uint64_t* rec_off, uint64_t cap_rec, uint64_t* n_rec) {
    try {
        auto recs = read_fasta(path);
        uint64_t tot = 0;
        for (auto& r : recs) tot += r.seq.size();
        *n_res = tot;
        *n_rec = recs.size();
        if (!residues || !rec_off) return 0;
        if (tot > cap_res || recs.size() + 1 > cap_rec) throw std::runtime_error("orc_read_fasta: buffers too small");
        uint64_t o = 0;
        for (size_t i = 0; i < recs.size(); i++) {
            rec_off[i] = o;
            std::memcpy(residues + o, recs[i].seq.data(), recs[i].seq.size());
            o += recs[i].seq.size();
        }
        rec_off[recs.size()] = o;
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// The reference's unit tests of group_extraction / group_sorting / group_filtering on the oracle's functions
// (selftest.cpp).  0 = all passed; otherwise the failing test's name is the error message.
int orc_self_test(void) {
    try {
        const std::string r = self_test();
        if (!r.empty()) throw std::runtime_error("reference unit test failed: " + r);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

}  // extern "C"
