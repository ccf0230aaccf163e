// group_filtering.cpp — CPU ORACLE (test infrastructure).
// Restates /root/reference/src/group_filtering.rs.  Non-ASCII Unicode whitespace in protein
// names (char::is_whitespace) is not modelled — ASCII whitespace only.
#include "kamino_oracle.hpp"
#include <algorithm>
#include <atomic>
#include <map>
#include <thread>
#include <unordered_map>
#include <unordered_set>

namespace kmo {

static inline bool is_missing_or_ambiguous(uint8_t b) { return b == '-' || b == 'X'; }  // :32-35
static std::string trim(const std::string& s) { return rs_trim(s); }   // str::trim: Unicode White_Space

std::vector<int> majority_recoded_consensus(const std::vector<std::vector<uint8_t>>& rows, size_t len,
                                            RecodeScheme s) {  // :37-67
    std::vector<int> consensus(len, -1);
    for (size_t c = 0; c < len; c++) {
        size_t counts[6] = {0, 0, 0, 0, 0, 0};
        for (auto& row : rows) {
            uint8_t b = row[c];
            if (is_missing_or_ambiguous(b)) continue;
            uint8_t r = recode_byte(b, s);
            if (r != 255) counts[r]++;
        }
        int best_state = -1;
        size_t best_count = 0;
        for (int st = 0; st < 6; st++)
            if (counts[st] > best_count) { best_state = st; best_count = counts[st]; }
        consensus[c] = best_state;
    }
    return consensus;
}

void mask_rows_with_long_divergence(std::vector<std::vector<uint8_t>>& rows,
                                    const std::vector<int>& consensus, size_t mask,
                                    RecodeScheme s) {  // :69-113
    if (mask == 0) return;
    for (auto& row : rows) {
        size_t consecutive = 0;
        bool should_mask = false;
        size_t m = std::min(row.size(), consensus.size());
        for (size_t c = 0; c < m; c++) {
            int cs = consensus[c];
            if (cs < 0) { consecutive = 0; continue; }
            uint8_t aa = row[c];
            if (is_missing_or_ambiguous(aa)) { consecutive = 0; continue; }
            uint8_t rs = recode_byte(aa, s);
            if (rs == 255) { consecutive = 0; continue; }
            if (int(rs) == cs) consecutive = 0;
            else {
                consecutive++;
                if (consecutive >= mask) { should_mask = true; break; }
            }
        }
        if (should_mask) std::fill(row.begin(), row.end(), uint8_t('X'));
    }
}

bool is_polymorphic_column(const std::vector<std::vector<uint8_t>>& rows, size_t col, RecodeScheme s) {  // :115-133
    int first = -1;
    for (auto& row : rows) {
        uint8_t b = row[col];
        if (is_missing_or_ambiguous(b)) continue;
        uint8_t r = recode_byte(b, s);
        if (r == 255) continue;
        if (first < 0) first = r;
        else if (first != int(r)) return true;
    }
    return false;
}

float missing_fraction(const std::vector<std::vector<uint8_t>>& rows, size_t col, size_t n) {  // :135-141
    size_t miss = 0;
    for (auto& r : rows)
        if (is_missing_or_ambiguous(r[col])) miss++;
    return float(miss) / float(n);
}

std::string consensus_protein_name(const std::vector<BlockObs>& observations,
                                   const std::vector<std::string>& protein_names) {  // :143-221
    std::vector<std::string> names;
    for (const BlockObs& obs : observations) {
        if (size_t(obs.protein_id) >= protein_names.size()) continue;
        std::string t = trim(protein_names[obs.protein_id]);
        size_t p = 0, wl = 0;
        while (p < t.size() && (wl = rs_ws_len(t, p)) == 0) p++;
        if (p >= t.size()) continue;  // split_once(char::is_whitespace) found no whitespace
        std::string desc = trim(t.substr(p + wl));
        if (!desc.empty()) names.push_back(desc);
    }
    if (names.empty()) return "";
    struct WordStats { size_t count, position_sum, first_seen_order; };
    std::unordered_map<std::string, WordStats> stats;
    size_t next_seen = 0;
    for (auto& name : names) {
        std::unordered_set<std::string> seen_in_name;
        size_t position = 0, i = 0;
        while (i < name.size()) {
            for (size_t l; i < name.size() && (l = rs_ws_len(name, i)) != 0;) i += l;
            if (i >= name.size()) break;
            size_t j = i;
            while (j < name.size() && rs_ws_len(name, j) == 0) j++;
            std::string word = name.substr(i, j - i);
            i = j;
            size_t pos = position++;
            if (!seen_in_name.insert(word).second) continue;
            auto it = stats.find(word);
            if (it != stats.end()) { it->second.count++; it->second.position_sum += pos; }
            else stats.emplace(word, WordStats{1, pos, next_seen++});
        }
    }
    std::vector<std::pair<std::string, WordStats>> maj;
    for (auto& kv : stats)
        if (kv.second.count * 2 >= names.size()) maj.push_back(kv);
    std::sort(maj.begin(), maj.end(), [](const auto& a, const auto& b) {
        size_t l = a.second.position_sum * b.second.count, r = b.second.position_sum * a.second.count;
        if (l != r) return l < r;
        if (a.second.first_seen_order != b.second.first_seen_order)
            return a.second.first_seen_order < b.second.first_seen_order;
        return a.first < b.first;
    });
    std::string out;
    for (size_t i = 0; i < maj.size(); i++) {
        if (i) out += " ";
        out += maj[i].first;
    }
    return out;
}

struct CandidateGroup { std::vector<uint8_t> rows; size_t kept_len = 0; std::string name; bool ok = false; };

static CandidateGroup build_candidate_group(const RawDirectCandidate& raw, const BlockArena& arena,
                                            const std::vector<std::string>& protein_names, size_t n,
                                            size_t k, size_t constant, float min_freq, size_t mask,
                                            RecodeScheme scheme) {  // :224-290
    CandidateGroup out;
    out.name = consensus_protein_name(raw.selected, protein_names);
    auto rows = consensus_rows_by_isolate(raw.selected, arena, n, raw.raw_len);
    size_t mid_start = k;
    if (raw.raw_len < k) return out;
    size_t mid_end = raw.raw_len - k;
    if (mid_end <= mid_start) return out;
    if (mask > 0) {
        auto consensus = majority_recoded_consensus(rows, raw.raw_len, scheme);
        mask_rows_with_long_divergence(rows, consensus, mask, scheme);
    }
    const float thresh = 1.0f - min_freq;  // f32, group_filtering.rs:254,274
    std::vector<size_t> retained_middle;
    for (size_t c = mid_start; c < mid_end; c++)
        if (missing_fraction(rows, c, n) <= thresh) retained_middle.push_back(c);
    if (retained_middle.empty()) return out;
    long first_poly = -1, last_poly = -1;
    for (size_t i = 0; i < retained_middle.size(); i++)
        if (is_polymorphic_column(rows, retained_middle[i], scheme)) {
            if (first_poly < 0) first_poly = long(i);
            last_poly = long(i);
        }
    if (first_poly < 0) return out;
    std::vector<size_t> keep_cols(retained_middle.begin() + first_poly,
                                  retained_middle.begin() + last_poly + 1);
    size_t tail_keep = std::min(constant, k);
    for (size_t c = mid_end; c < std::min(mid_end + tail_keep, raw.raw_len); c++)
        if (missing_fraction(rows, c, n) <= thresh) keep_cols.push_back(c);
    out.kept_len = keep_cols.size();
    out.rows.reserve(n * out.kept_len);
    for (auto& row : rows)
        for (size_t c : keep_cols) out.rows.push_back(row[c]);
    out.ok = true;
    return out;
}

AlignmentResult filter_groups(SortedGroups&& sorted, float min_freq, size_t constant, size_t mask,
                              size_t threads) {  // :292-363
    AlignmentResult res;
    size_t n = sorted.n_species;
    size_t m = sorted.raw_candidates.size();
    std::vector<CandidateGroup> built(m);
    {
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= m) break;
                built[i] = build_candidate_group(sorted.raw_candidates[i], sorted.arena,
                                                 sorted.protein_names, n, sorted.k, constant, min_freq,
                                                 mask, sorted.scheme);
            }
        };
        size_t nt = std::max<size_t>(1, std::min(threads, std::max<size_t>(m, 1)));
        std::vector<std::thread> pool;
        for (size_t t = 0; t < nt; t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    res.concat.assign(n, {});
    size_t pos = 0;
    for (auto& cand : built) {
        if (!cand.ok) continue;
        size_t blen = cand.kept_len;
        for (size_t sid = 0; sid < n; sid++)
            res.concat[sid].insert(res.concat[sid].end(), cand.rows.begin() + sid * blen,
                                   cand.rows.begin() + (sid + 1) * blen);
        // (pos, pos + blen - 1, blen): blen >= 1 always (a polymorphic middle column is kept)
        res.partitions.push_back({pos, pos + blen - 1, blen});
        res.partition_names.push_back(cand.name);
        pos += blen;
    }
    res.species_names = std::move(sorted.species_names);
    return res;
}

}  // namespace kmo
