// group_sorting.cpp — CPU ORACLE (test infrastructure).
// Restates /root/reference/src/group_sorting.rs.
#include "kamino_oracle.hpp"
#include <algorithm>
#include <unordered_set>

namespace kmo {

size_t direct_overlap_k(size_t k) {  // group_sorting.rs:12-22
    return k < 10 ? 2 * k + 1 : 21;
}

bool select_unique_best_supported_length(AnchorPair key, const std::vector<BlockObs>& obs,
                                         size_t n_species, size_t min_needed, size_t k,
                                         RawDirectCandidate& out) {  // group_sorting.rs:90-143
    struct LengthSupport { size_t len; std::vector<uint64_t> species; };
    std::vector<LengthSupport> by_len;
    size_t words = (n_species + 63) / 64;
    if (words == 0) words = 1;
    for (const BlockObs& o : obs) {
        size_t len = o.aa.len;
        LengthSupport* bucket = nullptr;
        for (auto& b : by_len)
            if (b.len == len) { bucket = &b; break; }
        if (!bucket) {
            by_len.push_back(LengthSupport{len, std::vector<uint64_t>(words, 0)});
            bucket = &by_len.back();
        }
        bucket->species[o.sid >> 6] |= uint64_t(1) << (o.sid & 63);
    }
    size_t best_len = 0, best_c = 0;
    bool tie = false;
    for (auto& b : by_len) {
        size_t c = 0;
        for (uint64_t w : b.species) c += size_t(__builtin_popcountll(w));
        if (c > best_c) { best_len = b.len; best_c = c; tie = false; }
        else if (c == best_c) tie = true;
    }
    if (tie) return false;
    if (best_c < min_needed) return false;
    if (best_len < 2 * k + 1) return false;
    out.key = key;
    out.coverage = best_c;
    out.raw_len = best_len;
    out.selected.clear();
    for (const BlockObs& o : obs)
        if (size_t(o.aa.len) == best_len) out.selected.push_back(o);
    return true;
}

static uint64_t kmer_mask(size_t k) {  // group_sorting.rs:145-153
    size_t bit_len = k * RECODE_BITS_PER_SYMBOL;
    return bit_len >= 64 ? ~uint64_t(0) : ((uint64_t(1) << bit_len) - 1);
}

// group_sorting.rs:155-200 (query) and :202-244 (insert) share the same rolling loop.
template <typename F>
static bool for_each_path_kmer(const std::vector<BlockObs>& observations, const BlockArena& arena,
                               size_t k, RecodeScheme scheme, F&& f) {
    uint64_t mask = kmer_mask(k);
    for (const BlockObs& obs : observations) {
        const uint8_t* aa = arena.get(obs.aa);
        size_t len = obs.aa.len;
        if (len < k) continue;
        uint64_t val = 0;
        size_t have = 0;
        for (size_t i = 0; i < len; i++) {
            uint8_t c = recode_byte(aa[i], scheme);
            if (c == 255) { val = 0; have = 0; continue; }
            val = ((val << RECODE_BITS_PER_SYMBOL) | uint64_t(c)) & mask;
            if (have < k) have++;
            if (have == k && f(val)) return true;
        }
    }
    return false;
}

std::vector<RawDirectCandidate> retain_nonoverlapping_raw_candidates(
    std::vector<RawDirectCandidate> cands, const BlockArena& arena, size_t overlap_k,
    RecodeScheme scheme) {  // group_sorting.rs:246-286
    std::sort(cands.begin(), cands.end(), [](const RawDirectCandidate& a, const RawDirectCandidate& b) {
        if (a.coverage != b.coverage) return a.coverage > b.coverage;
        if (a.raw_len != b.raw_len) return a.raw_len > b.raw_len;
        if (a.key.left != b.key.left) return a.key.left < b.key.left;
        return a.key.right < b.key.right;
    });
    std::unordered_set<uint64_t> used_global;
    std::vector<RawDirectCandidate> kept;
    for (auto& cand : cands) {
        bool overlaps = for_each_path_kmer(cand.selected, arena, overlap_k, scheme, [&](uint64_t v) {
            return used_global.count(v) != 0;
        });
        if (overlaps) continue;
        for_each_path_kmer(cand.selected, arena, overlap_k, scheme, [&](uint64_t v) {
            used_global.insert(v);
            return false;
        });
        kept.push_back(std::move(cand));
    }
    return kept;
}

std::vector<std::vector<uint8_t>> consensus_rows_by_isolate(const std::vector<BlockObs>& observations,
                                                            const BlockArena& arena,
                                                            size_t n_species, size_t len) {  // :288-308
    std::vector<std::vector<uint8_t>> rows(n_species, std::vector<uint8_t>(len, '-'));
    for (const BlockObs& obs : observations) {
        auto& row = rows[obs.sid];
        const uint8_t* aa = arena.get(obs.aa);
        size_t m = std::min(len, size_t(obs.aa.len));
        for (size_t i = 0; i < m; i++) {
            uint8_t& cell = row[i];
            if (cell == '-') cell = aa[i];
            else if (cell == aa[i]) {}
            else cell = 'X';
        }
    }
    return rows;
}

SortedGroups sort_and_deduplicate_groups(RawExtractedGroups&& raw, RecodeScheme scheme) {  // :329-380
    SortedGroups out;
    std::vector<RawDirectCandidate> raw_candidates;
    raw_candidates.reserve(raw.groups.size());
    for (auto& kv : raw.groups) {
        RawDirectCandidate c;
        if (select_unique_best_supported_length(kv.first, kv.second, raw.n_species, raw.min_needed,
                                                raw.k, c))
            raw_candidates.push_back(std::move(c));
    }
    out.raw_candidates = retain_nonoverlapping_raw_candidates(std::move(raw_candidates), raw.arena,
                                                              direct_overlap_k(raw.k), scheme);
    out.species_names = std::move(raw.species_names);
    out.arena = std::move(raw.arena);
    out.protein_names = std::move(raw.protein_names);
    out.k = raw.k;
    out.min_needed = raw.min_needed;
    out.n_species = raw.n_species;
    out.scheme = scheme;
    return out;
}

}  // namespace kmo
