// group_extraction.cpp — CPU ORACLE (test infrastructure).
// Restates /root/reference/src/group_extraction.rs (pass 1 count, pass 2 scan, anchor
// pairing, deterministic merge).
#include "kamino_oracle.hpp"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>

namespace kmo {

static bool is_ascii_whitespace(uint8_t b) {  // Rust u8::is_ascii_whitespace: SP, TAB, LF, FF, CR
    return b == 0x20 || b == 0x09 || b == 0x0A || b == 0x0C || b == 0x0D;
}
static uint8_t to_ascii_uppercase(uint8_t b) { return (b >= 'a' && b <= 'z') ? uint8_t(b - 32) : b; }

BlockAa BlockArena::push(const uint8_t* p, size_t n) {  // group_extraction.rs:90-100
    size_t start = data.size();
    if (n > 0xFFFF)
        throw std::runtime_error("amino-acid block length " + std::to_string(n) + " exceeds u16::MAX");
    data.insert(data.end(), p, p + n);
    return BlockAa{start, uint16_t(n)};
}
size_t BlockArena::append(const BlockArena& o) {  // group_extraction.rs:102-106
    size_t off = data.size();
    data.insert(data.end(), o.data.begin(), o.data.end());
    return off;
}

static bool is_better_start(const KmerHit& c, const KmerHit& best) {  // :146-149
    return c.occupancy > best.occupancy || (c.occupancy == best.occupancy && c.start < best.start);
}

std::vector<KmerHit> select_best_starts_per_local_cluster(const std::vector<KmerHit>& starts) {  // :151-175
    std::vector<KmerHit> selected;
    if (starts.empty()) return selected;
    size_t cluster_start_pos = starts[0].start;
    KmerHit best = starts[0];
    for (size_t i = 1; i < starts.size(); i++) {
        const KmerHit& cand = starts[i];
        if (cand.start < cluster_start_pos + START_CLUSTER_SPAN) {
            if (is_better_start(cand, best)) best = cand;
        } else {
            selected.push_back(best);
            cluster_start_pos = cand.start;
            best = cand;
        }
    }
    selected.push_back(best);
    return selected;
}

// The anchor search of extract_bubble_pairs (group_extraction.rs:188-231): for every selected start anchor,
// in order, its best end anchor (index into end_anchors) or none.
std::vector<std::pair<size_t, size_t>> pair_anchors(const std::vector<KmerHit>& start_anchors,
                                                    const std::vector<KmerHit>& end_anchors, size_t length_middle) {
    std::vector<std::pair<size_t, size_t>> out;
    size_t first_end = 0;
    for (size_t li = 0; li < start_anchors.size(); li++) {
        const KmerHit& left = start_anchors[li];
        while (first_end < end_anchors.size() && end_anchors[first_end].start <= left.end) first_end++;
        size_t j = first_end;
        bool have_best = false;
        size_t best_j = 0;
        KmerHit best{};
        while (j < end_anchors.size()) {
            const KmerHit& right = end_anchors[j];
            if (right.start <= left.end) { j++; continue; }
            size_t middle_len = right.start - left.end;
            if (middle_len > length_middle) break;
            bool replace = !have_best || right.occupancy > best.occupancy ||
                           (right.occupancy == best.occupancy && right.start < best.start);
            if (replace) { best = right; best_j = j; have_best = true; }
            j++;
        }
        if (have_best) out.emplace_back(li, best_j);
    }
    return out;
}

void extract_bubble_pairs(const uint8_t* aa, const std::vector<KmerHit>& start_anchors,
                          const std::vector<KmerHit>& end_anchors, size_t length_middle,
                          uint16_t sid, uint32_t protein_id, RawDirectGroups& groups,
                          BlockArena& arena) {  // :178-242
    for (const auto& pr : pair_anchors(start_anchors, end_anchors, length_middle)) {
        const KmerHit& left = start_anchors[pr.first];
        const KmerHit& best = end_anchors[pr.second];
        BlockAa blk = arena.push(aa + left.start, best.end - left.start);
        groups[AnchorPair{left.kmer, best.kmer}].push_back(BlockObs{sid, protein_id, blk});
    }
}

void scan_valid_segment(const uint8_t* recoded, size_t len, size_t k, uint64_t k_mask,
                        const CountMinSketch& cms, uint32_t min_needed,
                        std::vector<KmerHit>& start_candidates, std::vector<KmerHit>& end_anchors,
                        std::vector<uint16_t>* occ_out) {  // :264-303
    bool have_prev = false;
    KmerHit previous{};
    uint64_t roll = 0;
    size_t have = 0;
    for (size_t pos = 0; pos < len; pos++) {
        roll = ((roll << RECODE_BITS_PER_SYMBOL) | uint64_t(recoded[pos])) & k_mask;
        if (have < k) have++;
        if (have == k) {
            uint32_t occ = cms.estimate(roll);
            bool shared = occ >= min_needed;
            size_t start = pos + 1 - k;
            KmerHit cur{roll, start, start + k, occ, shared};
            if (occ_out) occ_out->push_back(uint16_t(occ));
            if (have_prev) {
                if (previous.shared && previous.occupancy > cur.occupancy)
                    start_candidates.push_back(previous);
                if (cur.shared && previous.occupancy < cur.occupancy) end_anchors.push_back(cur);
            }
            previous = cur;
            have_prev = true;
        }
    }
}

void extract_from_valid_segment(const uint8_t* aa, const uint8_t* recoded, size_t len, size_t k,
                                uint64_t k_mask, const CountMinSketch& cms, uint32_t min_needed,
                                size_t length_middle, uint16_t sid, uint32_t protein_id,
                                RawDirectGroups& groups, BlockArena& arena) {  // :245-322
    if (len < k || k == 0) return;
    std::vector<KmerHit> start_candidates, end_anchors;
    scan_valid_segment(recoded, len, k, k_mask, cms, min_needed, start_candidates, end_anchors, nullptr);
    std::vector<KmerHit> start_anchors = select_best_starts_per_local_cluster(start_candidates);
    extract_bubble_pairs(aa, start_anchors, end_anchors, length_middle, sid, protein_id, groups, arena);
}

void merge_species_extraction(SpeciesExtraction&& ex, RawDirectGroups& groups, BlockArena& arena,
                                     std::vector<std::string>& protein_names) {  // :335-357
    size_t protein_offset = protein_names.size();
    if (protein_offset + ex.protein_names.size() > 0xFFFFFFFFull)
        throw std::runtime_error("too many protein records across all species to index with u32");
    for (auto& n : ex.protein_names) protein_names.push_back(std::move(n));
    size_t arena_offset = arena.append(ex.arena);
    for (auto& kv : ex.groups) {
        auto& dst = groups[kv.first];
        for (BlockObs obs : kv.second) {
            obs.protein_id += uint32_t(protein_offset);
            obs.aa.start += arena_offset;
            dst.push_back(obs);
        }
    }
}

uint64_t count_species_records(const std::vector<FastaRecord>& recs, size_t k, uint64_t k_mask,
                               RecodeScheme scheme, AtomicCountMinSketch& cmsa, BloomFilter& bloom) {  // :359-395
    uint64_t n_new = 0;
    for (const FastaRecord& rec : recs) {
        uint64_t roll = 0;
        size_t have = 0;
        for (unsigned char b : rec.seq) {
            if (is_ascii_whitespace(b)) continue;
            uint8_t a = recode_byte(to_ascii_uppercase(b), scheme);
            if (a == 255) { roll = 0; have = 0; continue; }
            roll = ((roll << RECODE_BITS_PER_SYMBOL) | uint64_t(a)) & k_mask;
            if (have < k) have++;
            if (have == k && !bloom.test_and_set(roll)) { cmsa.increment(roll); n_new++; }
        }
    }
    return n_new;
}

std::string truncate_protein_name(const std::string& raw) {  // :397-404
    std::string t = rs_trim(raw);
    size_t p = t.find(" [");
    if (p == std::string::npos) return t;
    return rs_trim_end(t.substr(0, p));
}

static bool valid_utf8(const std::string& s) {
    size_t i = 0, n = s.size();
    while (i < n) {
        unsigned char c = (unsigned char)s[i];
        if (c < 0x80) { i++; continue; }
        size_t need; uint32_t cp;
        if (c >= 0xC2 && c <= 0xDF) { need = 1; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { need = 2; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { need = 3; cp = c & 0x07; }
        else return false;
        for (size_t j = 1; j <= need; j++) {
            if (i + j >= n) return false;
            unsigned char d = (unsigned char)s[i + j];
            if ((d & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (d & 0x3F);
        }
        if (need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
        if (need == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
        i += need + 1;
    }
    return true;
}

SpeciesExtraction extract_species_records(size_t sid, const std::string& species_name,
                                          const std::vector<FastaRecord>& recs, size_t k,
                                          uint64_t k_mask, const CountMinSketch& cms,
                                          uint32_t min_needed, size_t length_middle,
                                          RecodeScheme scheme) {  // :407-507
    SpeciesExtraction out;
    out.sid = sid;
    size_t min_block_len = 2 * k + 1;
    std::vector<uint8_t> aa_segment, recoded_segment;
    for (const FastaRecord& rec : recs) {
        // seq_io: id = head up to the first space; desc = the rest (if a space exists).
        size_t sp = rec.head.find(' ');
        std::string id = sp == std::string::npos ? rec.head : rec.head.substr(0, sp);
        bool has_desc = sp != std::string::npos;
        std::string desc = has_desc ? rec.head.substr(sp + 1) : std::string();
        if (!valid_utf8(id) || (has_desc && !valid_utf8(desc)))
            throw std::runtime_error("invalid UTF-8 in FASTA header of species " + species_name);
        std::string protein_record_id = rs_trim(id);
        std::string protein_desc = rs_trim(desc);
        std::string raw_name = protein_desc.empty() ? protein_record_id
                                                    : protein_record_id + " " + protein_desc;
        size_t protein_id_sz = out.protein_names.size();
        if (protein_id_sz > 0xFFFFFFFFull)
            throw std::runtime_error("too many protein records in species " + species_name +
                                     " to index with u32");
        out.protein_names.push_back(truncate_protein_name(raw_name));
        uint32_t protein_id = uint32_t(protein_id_sz);
        aa_segment.clear();
        recoded_segment.clear();
        auto flush = [&]() {
            if (recoded_segment.size() >= min_block_len)
                extract_from_valid_segment(aa_segment.data(), recoded_segment.data(),
                                           recoded_segment.size(), k, k_mask, cms, min_needed,
                                           length_middle, uint16_t(sid), protein_id, out.groups,
                                           out.arena);
        };
        for (unsigned char b : rec.seq) {
            if (is_ascii_whitespace(b)) continue;
            uint8_t up = to_ascii_uppercase(b);
            uint8_t rv = recode_byte(up, scheme);
            if (rv == 255) {
                flush();
                aa_segment.clear();
                recoded_segment.clear();
            } else {
                aa_segment.push_back(up);
                recoded_segment.push_back(rv);
            }
        }
        flush();
    }
    return out;
}

uint32_t min_needed_for(float min_freq, size_t n) {  // :544 — f32 arithmetic
    float nn = float(n < 1 ? size_t(1) : n);
    volatile float prod = min_freq * nn;
    return uint32_t(std::ceil(float(prod)));
}
uint64_t k_mask_for(size_t k) {  // :545-549
    size_t bits = k * RECODE_BITS_PER_SYMBOL;
    return bits >= 64 ? ~uint64_t(0) : ((uint64_t(1) << bits) - 1);
}

RawExtractedGroups extract_groups(const std::vector<SpeciesInput>& inputs, size_t k, float min_freq,
                                  size_t length_middle, RecodeScheme scheme, size_t threads) {  // :527-660
    size_t n = inputs.size();
    if (n > 0xFFFF)
        throw std::runtime_error("too many species (" + std::to_string(n) +
                                 ") for compact direct block observations; maximum is 65535");
    uint32_t min_needed = min_needed_for(min_freq, n);
    uint64_t k_mask = k_mask_for(k);
    size_t n_threads = threads < 1 ? 1 : threads;

    // Stage 2: species-parallel counting into the shared atomic sketch (:553-571).
    std::fprintf(stderr, " . pass 1: estimate k-mer occupancies\n");
    AtomicCountMinSketch cmsa(CMS_WIDTH, CMS_DEPTH);
    std::vector<std::string> errors(n);
    {
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            BloomFilter bloom(BLOOM_BITS);
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                try {
                    bloom.clear();  // a fresh BloomFilter per species (:370)
                    auto recs = read_fasta(inputs[i].path);
                    count_species_records(recs, k, k_mask, scheme, cmsa, bloom);
                } catch (const std::exception& e) {
                    errors[i] = e.what();
                    if (errors[i].empty()) errors[i] = "error";
                }
            }
        };
        std::vector<std::thread> pool;
        for (size_t t = 0; t < std::min(n_threads, std::max<size_t>(n, 1)); t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    for (auto& e : errors)
        if (!e.empty()) throw std::runtime_error(e);

    // Stage 3: immutable snapshot (:576).
    CountMinSketch cms = cmsa.snapshot();

    // Stage 4: per-species extraction, merged strictly in sid order (:578-650).
    std::fprintf(stderr, " . pass 2: extract sequences\n");
    RawExtractedGroups out;
    std::vector<SpeciesExtraction> results(n);
    std::vector<std::string> errors2(n);
    {
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                try {
                    auto recs = read_fasta(inputs[i].path);
                    results[i] = extract_species_records(i, inputs[i].name, recs, k, k_mask, cms,
                                                         min_needed, length_middle, scheme);
                } catch (const std::exception& e) {
                    errors2[i] = e.what();
                    if (errors2[i].empty()) errors2[i] = "error";
                }
            }
        };
        std::vector<std::thread> pool;
        for (size_t t = 0; t < std::min(n_threads, std::max<size_t>(n, 1)); t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    for (auto& e : errors2)
        if (!e.empty()) throw std::runtime_error(e);
    for (size_t i = 0; i < n; i++)
        merge_species_extraction(std::move(results[i]), out.groups, out.arena, out.protein_names);
    for (auto& in : inputs) out.species_names.push_back(in.name);
    out.k = k;
    out.min_needed = min_needed;
    out.n_species = n;
    return out;
}

}  // namespace kmo
