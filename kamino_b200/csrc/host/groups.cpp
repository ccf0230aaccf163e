// groups.cpp — candidate selection, non-overlap filter, column filtering and alignment assembly on
// the flat observation table.  Equivalent to /root/reference/src/group_sorting.rs:10-380 and
// /root/reference/src/group_filtering.rs:32-363.
#include "kamino_host.hpp"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <unordered_set>

namespace kh {

uint8_t recode(uint8_t b, Scheme s) {   // recode.rs:31-80
    static const char* classes[3][6] = {
        {"C", "AGPST", "DENQ", "HKR", "ILMV", "FWY"},
        {"APST", "DENG", "QKR", "MIVL", "WC", "FYH"},
        {"AGPS", "DENQHKRT", "MIL", "W", "FY", "CV"},
    };
    static uint8_t lut[3][256];
    static const bool init = [] {
        std::memset(lut, 255, sizeof lut);
        for (int sc = 0; sc < 3; sc++)
            for (int c = 0; c < 6; c++)
                for (const char* p = classes[sc][c]; *p; p++) {
                    lut[sc][uint8_t(*p)] = uint8_t(c);
                    lut[sc][uint8_t(*p) + 32] = uint8_t(c);
                }
        return true;
    }();
    (void)init;
    return lut[int(s)][b];
}

namespace {

inline bool gap_or_x(uint8_t b) { return b == '-' || b == 'X'; }   // group_filtering.rs:32-35

// Rolling recoded k-mers of a block (group_sorting.rs:155-244); f returns true to stop early.
template <typename F>
bool roll_block(const uint8_t* aa, size_t len, size_t k, Scheme scheme, F&& f) {
    if (len < k) return false;
    const uint64_t mask = 3 * k >= 64 ? ~0ull : ((1ull << (3 * k)) - 1);
    uint64_t val = 0;
    size_t have = 0;
    for (size_t i = 0; i < len; i++) {
        const uint8_t c = recode(aa[i], scheme);
        if (c == 255) { val = 0; have = 0; continue; }
        val = ((val << 3) | c) & mask;
        if (have < k) have++;
        if (have == k && f(val)) return true;
    }
    return false;
}

// Open-addressing set of u64 keys (linear probing, power-of-two capacity): the greedy non-overlap filter
// inserts tens of millions of recoded k-mers, where std::unordered_set's node allocations dominate.
class KmerSet {
    std::vector<uint64_t> slot_;
    std::vector<uint8_t> full_;
    size_t n_ = 0, mask_ = 0;
    static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; return x; }
    void grow() {
        std::vector<uint64_t> old;
        std::vector<uint8_t> of;
        old.swap(slot_);
        of.swap(full_);
        const size_t cap = old.empty() ? (size_t(1) << 16) : old.size() * 2;
        slot_.assign(cap, 0);
        full_.assign(cap, 0);
        mask_ = cap - 1;
        n_ = 0;
        for (size_t i = 0; i < old.size(); i++) if (of[i]) insert(old[i]);
    }
public:
    bool contains(uint64_t k) const {
        if (slot_.empty()) return false;
        for (size_t i = mix(k) & mask_;; i = (i + 1) & mask_) {
            if (!full_[i]) return false;
            if (slot_[i] == k) return true;
        }
    }
    void insert(uint64_t k) {
        if ((n_ + 1) * 2 > slot_.size()) grow();
        for (size_t i = mix(k) & mask_;; i = (i + 1) & mask_) {
            if (!full_[i]) { full_[i] = 1; slot_[i] = k; n_++; return; }
            if (slot_[i] == k) return;
        }
    }
};

}  // namespace

std::vector<Candidate> sort_and_deduplicate_groups(const Extraction& ex, Scheme scheme) {   // group_sorting.rs:329-380
    std::vector<Candidate> cands;
    const auto& obs = ex.obs;
    // support of every block length = number of DISTINCT species (group_sorting.rs:90-143); buckets are visited
    // in first-appearance order, ties reject the group.  Only for the runs the device left undecided.
    auto select_on_host = [&](size_t g0, size_t g1, size_t& best_len, size_t& best_c) -> bool {
        std::vector<std::pair<uint32_t, uint32_t>> len_sid;
        std::vector<uint32_t> lens_in_order;
        for (size_t i = g0; i < g1; i++) {
            len_sid.emplace_back(obs[i].len, obs[i].sid);
            if (std::find(lens_in_order.begin(), lens_in_order.end(), obs[i].len) == lens_in_order.end())
                lens_in_order.push_back(obs[i].len);
        }
        std::sort(len_sid.begin(), len_sid.end());
        len_sid.erase(std::unique(len_sid.begin(), len_sid.end()), len_sid.end());
        best_len = 0;
        best_c = 0;
        bool tie = false;
        for (uint32_t len : lens_in_order) {
            const size_t c = size_t(std::count_if(len_sid.begin(), len_sid.end(), [len](const auto& p) { return p.first == len; }));
            if (c > best_c) { best_len = len; best_c = c; tie = false; }
            else if (c == best_c) tie = true;
        }
        return !tie && best_c >= ex.min_needed && best_len >= 2 * ex.k + 1;
    };
    for (const kmn_group& g : ex.groups) {   // runs of equal (left, right) with the device's length selection
        const size_t g0 = g.first, g1 = size_t(g.first) + g.count;
        size_t best_len = g.best_len, best_c = g.coverage;
        bool ok = (g.flags & KMN_GROUP_VALID) != 0;
        if (g.flags & KMN_GROUP_HOST) ok = select_on_host(g0, g1, best_len, best_c);
        if (!ok) continue;
        Candidate c;
        c.left = obs[g0].left;
        c.right = obs[g0].right;
        c.coverage = best_c;
        c.raw_len = best_len;
        for (size_t i = g0; i < g1; i++) if (obs[i].len == best_len) c.rows.push_back(uint32_t(i));
        cands.push_back(std::move(c));
    }
    // strongest first: coverage desc, length desc, key asc (group_sorting.rs:252-258)
    std::sort(cands.begin(), cands.end(), [](const Candidate& a, const Candidate& b) {
        if (a.coverage != b.coverage) return a.coverage > b.coverage;
        if (a.raw_len != b.raw_len) return a.raw_len > b.raw_len;
        if (a.left != b.left) return a.left < b.left;
        return a.right < b.right;
    });
    // greedy non-overlap on recoded overlap_k-mers (group_sorting.rs:12-22, :246-286)
    const size_t overlap_k = ex.k < 10 ? 2 * ex.k + 1 : 21;
    KmerSet used;
    std::vector<Candidate> kept;
    for (Candidate& c : cands) {
        bool clash = false;
        for (uint32_t r : c.rows) {
            const Observation& o = obs[r];
            if (roll_block(ex.arena.data() + o.aa_off, o.len, overlap_k, scheme, [&](uint64_t v) { return used.contains(v); })) {
                clash = true;
                break;
            }
        }
        if (clash) continue;
        for (uint32_t r : c.rows) {
            const Observation& o = obs[r];
            roll_block(ex.arena.data() + o.aa_off, o.len, overlap_k, scheme, [&](uint64_t v) { used.insert(v); return false; });
        }
        kept.push_back(std::move(c));
    }
    return kept;
}

namespace {


// Majority-rule label over the descriptions of the supporting proteins (group_filtering.rs:143-221).
std::string consensus_protein_name(const Extraction& ex, const Candidate& c) {
    std::vector<std::string> names;
    for (uint32_t r : c.rows) {
        const uint32_t pid = ex.obs[r].protein;
        if (pid >= ex.protein_names.size()) continue;
        const std::string t = uws::trim(ex.protein_names[pid]);
        size_t p = 0, wl = 0;   // split_once(char::is_whitespace): first white-space CHARACTER (Unicode), :153-155
        while (p < t.size() && (wl = uws::at(t, p)) == 0) p++;
        if (p >= t.size()) continue;
        std::string d = uws::trim(t.substr(p + wl));
        if (!d.empty()) names.push_back(std::move(d));
    }
    if (names.empty()) return "";
    struct Stat { size_t count, pos_sum, first_seen; };
    std::unordered_map<std::string, Stat> stats;
    size_t seen_order = 0;
    for (const std::string& nm : names) {
        std::unordered_set<std::string> in_name;
        size_t pos = 0;
        for (size_t i = 0; i < nm.size();) {   // split_whitespace (:179)
            for (size_t l; i < nm.size() && (l = uws::at(nm, i)) != 0;) i += l;
            if (i >= nm.size()) break;
            size_t j = i;
            while (j < nm.size() && uws::at(nm, j) == 0) j++;
            std::string w = nm.substr(i, j - i);
            i = j;
            const size_t at = pos++;
            if (!in_name.insert(w).second) continue;
            auto it = stats.find(w);
            if (it == stats.end()) stats.emplace(std::move(w), Stat{1, at, seen_order++});
            else { it->second.count++; it->second.pos_sum += at; }
        }
    }
    std::vector<std::pair<std::string, Stat>> maj;
    for (auto& kv : stats) if (kv.second.count * 2 >= names.size()) maj.push_back(kv);
    std::sort(maj.begin(), maj.end(), [](const auto& a, const auto& b) {
        const size_t l = a.second.pos_sum * b.second.count, r = b.second.pos_sum * a.second.count;
        if (l != r) return l < r;
        if (a.second.first_seen != b.second.first_seen) return a.second.first_seen < b.second.first_seen;
        return a.first < b.first;
    });
    std::string out;
    for (size_t i = 0; i < maj.size(); i++) out += (i ? " " : "") + maj[i].first;
    return out;
}

// majority recoded state per column of the n x L row matrix, lowest state wins ties, -1 = no valid state
// (group_filtering.rs:37-67)
std::vector<int> majority_recoded_consensus(const uint8_t* rows, size_t n, size_t L, Scheme scheme) {
    std::vector<int> cons(L, -1);
    for (size_t col = 0; col < L; col++) {
        size_t cnt[6] = {0, 0, 0, 0, 0, 0};
        for (size_t s = 0; s < n; s++) {
            const uint8_t b = rows[s * L + col];
            if (gap_or_x(b)) continue;
            const uint8_t rc = recode(b, scheme);
            if (rc != 255) cnt[rc]++;
        }
        size_t best = 0;
        for (int st = 0; st < 6; st++) if (cnt[st] > best) { best = cnt[st]; cons[col] = st; }
    }
    return cons;
}

// rows with >= mask consecutive recoded mismatches against the consensus are blanked with 'X'
// (group_filtering.rs:69-113); mask == 0 disables it
void mask_rows_with_long_divergence(uint8_t* rows, size_t n, size_t L, const std::vector<int>& cons, size_t mask, Scheme scheme) {
    if (mask == 0) return;
    for (size_t s = 0; s < n; s++) {
        uint8_t* row = rows + s * L;
        size_t run = 0;
        bool hit = false;
        for (size_t col = 0; col < L; col++) {
            if (cons[col] < 0 || gap_or_x(row[col])) { run = 0; continue; }
            const uint8_t rc = recode(row[col], scheme);
            if (rc == 255 || int(rc) == cons[col]) { run = 0; continue; }
            if (++run >= mask) { hit = true; break; }
        }
        if (hit) std::memset(row, 'X', L);
    }
}

struct Built { std::vector<uint8_t> rows; size_t kept = 0; std::string name; bool ok = false; };

Built build_group(const Extraction& ex, const Candidate& c, size_t n, size_t k, size_t constant, float min_freq,
                  size_t mask, Scheme scheme) {   // group_filtering.rs:224-290
    Built out;
    out.name = consensus_protein_name(ex, c);
    const size_t L = c.raw_len;
    // one consensus row per isolate: '-' -> first aa; equal -> keep; different -> sticky 'X' (group_sorting.rs:288-308)
    std::vector<uint8_t> rows(n * L, '-');
    for (uint32_t r : c.rows) {
        const Observation& o = ex.obs[r];
        uint8_t* row = rows.data() + size_t(o.sid) * L;
        const uint8_t* aa = ex.arena.data() + o.aa_off;
        for (size_t i = 0; i < std::min<size_t>(L, o.len); i++) {
            if (row[i] == '-') row[i] = aa[i];
            else if (row[i] != aa[i]) row[i] = 'X';
        }
    }
    if (L < k || L - k <= k) return out;
    const size_t mid0 = k, mid1 = L - k;
    if (mask > 0) mask_rows_with_long_divergence(rows.data(), n, L, majority_recoded_consensus(rows.data(), n, L, scheme), mask, scheme);
    const float thresh = 1.0f - min_freq;   // f32 (group_filtering.rs:254,274)
    auto missing_ok = [&](size_t col) {
        size_t miss = 0;
        for (size_t s = 0; s < n; s++) miss += gap_or_x(rows[s * L + col]);
        return float(miss) / float(n) <= thresh;
    };
    auto polymorphic = [&](size_t col) {   // group_filtering.rs:115-133
        int first = -1;
        for (size_t s = 0; s < n; s++) {
            const uint8_t b = rows[s * L + col];
            if (gap_or_x(b)) continue;
            const uint8_t rc = recode(b, scheme);
            if (rc == 255) continue;
            if (first < 0) first = rc;
            else if (first != int(rc)) return true;
        }
        return false;
    };
    std::vector<size_t> mid;
    for (size_t col = mid0; col < mid1; col++) if (missing_ok(col)) mid.push_back(col);
    if (mid.empty()) return out;
    long p0 = -1, p1 = -1;
    for (size_t i = 0; i < mid.size(); i++) if (polymorphic(mid[i])) { if (p0 < 0) p0 = long(i); p1 = long(i); }
    if (p0 < 0) return out;
    std::vector<size_t> keep(mid.begin() + p0, mid.begin() + p1 + 1);
    for (size_t col = mid1; col < std::min(mid1 + std::min(constant, k), L); col++) if (missing_ok(col)) keep.push_back(col);
    out.kept = keep.size();
    out.rows.reserve(n * out.kept);
    for (size_t s = 0; s < n; s++) for (size_t col : keep) out.rows.push_back(rows[s * L + col]);
    out.ok = true;
    return out;
}

}  // namespace

Alignment filter_groups(const Extraction& ex, const std::vector<Candidate>& cands, float min_freq, size_t constant,
                        size_t mask, Scheme scheme, size_t threads) {   // group_filtering.rs:292-363
    const size_t n = ex.n_species, m = cands.size();
    std::vector<Built> built(m);
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < m;) built[i] = build_group(ex, cands[i], n, ex.k, constant, min_freq, mask, scheme);
    };
    std::vector<std::thread> pool;
    for (size_t t = 0; t < std::max<size_t>(1, std::min(threads, std::max<size_t>(m, 1))); t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
    Alignment al;
    al.species_names = ex.species_names;
    al.concat.assign(n, {});
    size_t pos = 0;
    for (const Built& b : built) {   // candidate order == raw-candidate order (:339)
        if (!b.ok) continue;
        for (size_t s = 0; s < n; s++)
            al.concat[s].insert(al.concat[s].end(), b.rows.begin() + s * b.kept, b.rows.begin() + (s + 1) * b.kept);
        al.partitions.push_back({pos, pos + b.kept - 1, b.kept});
        al.partition_names.push_back(b.name);
        pos += b.kept;
    }
    return al;
}

// ---- the reference's unit tests of this logic, restated on the functions above (kamino-b200 --self-test) -------
// group_filtering.rs:484-693 (every filter_groups case) and group_sorting.rs:486-638 (the sort / non-overlap /
// length-selection cases).  Observations are built by hand; groups carry KMN_GROUP_HOST so that the length
// selection runs here and not on the device.

namespace {

struct TestGroup { uint64_t left, right; std::vector<std::string> rows; std::vector<uint32_t> sids; };

Extraction test_extraction(const std::vector<TestGroup>& groups, size_t k, size_t min_needed, size_t n_species) {
    Extraction ex;
    ex.k = k;
    ex.min_needed = min_needed;
    ex.n_species = n_species;
    for (size_t s = 0; s < n_species; s++) ex.species_names.push_back("species_" + std::to_string(s));
    ex.protein_names.assign(n_species, "protein");
    std::vector<TestGroup> sorted = groups;   // obs are kept sorted by (left, right, order)
    std::stable_sort(sorted.begin(), sorted.end(), [](const TestGroup& a, const TestGroup& b) {
        return a.left != b.left ? a.left < b.left : a.right < b.right;
    });
    for (const TestGroup& g : sorted) {
        kmn_group kg{};
        kg.first = uint32_t(ex.obs.size());
        kg.count = uint32_t(g.rows.size());
        kg.flags = KMN_GROUP_HOST;
        for (size_t i = 0; i < g.rows.size(); i++) {
            Observation o{};
            o.left = g.left;
            o.right = g.right;
            o.sid = g.sids.empty() ? uint32_t(i) : g.sids[i];
            o.protein = o.sid;
            o.aa_off = ex.arena.size();
            o.len = uint32_t(g.rows[i].size());
            o.order = ex.obs.size();
            ex.arena.insert(ex.arena.end(), g.rows[i].begin(), g.rows[i].end());
            ex.obs.push_back(o);
        }
        ex.groups.push_back(kg);
    }
    ex.n_groups = ex.groups.size();
    return ex;
}

struct FilterCase {
    const char* name;
    std::vector<std::string> rows;
    size_t k;
    float min_freq;
    size_t constant, mask;
    std::vector<std::string> want;   // empty: the group is discarded
};

}  // namespace

std::string host_logic_self_test() {
    const std::vector<FilterCase> cases = {
        {"retained_middle_starts_at_first_retained_polymorphic_column", {"ACCEF", "ACCHF", "ACCEF", "ACCHF"}, 1, 1.0f, 0, 0, {"E", "H", "E", "H"}},
        {"retained_middle_trims_past_raw_only_polymorphism", {"AAPCF", "APDCF", "AAPCF", "APDCF"}, 1, 1.0f, 0, 0, {"P", "D", "P", "D"}},
        {"ambiguous_x_does_not_create_early_polymorphism", {"ACCEF", "ACCHF", "AXCEF", "ACXHF"}, 1, 0.75f, 0, 0, {"E", "H", "E", "H"}},
        {"discards_when_only_retained_middle_columns_are_conserved", {"ACCEF", "ACCDF", "ACC-F", "ACC-F"}, 1, 0.75f, 0, 0, {}},
        {"right_anchor_tail_is_appended_after_polymorphic_middle", {"ACEFT", "ACHFT", "ACEFT", "ACHFT"}, 1, 1.0f, 1, 0, {"ET", "HT", "ET", "HT"}},
        {"polymorphic_right_anchor_tail_cannot_rescue_conserved_middle", {"ACCD", "ACCE", "ACCD", "ACCE"}, 1, 1.0f, 1, 0, {}},
        {"discards_when_only_raw_polymorphic_middle_column_fails_missingness", {"ACGT", "ADGT", "A-GT", "A-GT"}, 1, 0.75f, 0, 0, {}},
        {"retains_group_with_retained_polymorphic_middle_column", {"ACGT", "ADGT", "ACGT", "ADGT"}, 1, 0.75f, 1, 0, {"CT", "DT", "CT", "DT"}},
        {"constant_tail_cannot_rescue_without_retained_polymorphic_middle", {"AC-T", "AD-T", "A--T", "A--T"}, 1, 0.75f, 1, 0, {}},
        {"middle_is_trimmed_on_both_sides_before_appending_constants", {"AAAQFCPQRRR", "AAAQDCDQRRR"}, 3, 1.0f, 3, 0, {"FCPRRR", "DCDRRR"}},
        {"missingness_filtering_happens_before_polymorphic_trimming", {"AFCPR", "ADCDR", "A-CPR", "A-CDR"}, 1, 0.75f, 0, 0, {"P", "D", "P", "D"}},
        {"raw_amino_acid_differences_in_same_recoded_state_do_not_count_as_polymorphism", {"AACA", "APCA"}, 1, 1.0f, 0, 0, {}},
        {"gaps_and_x_are_ignored_when_detecting_polymorphism", {"ACPR", "AXPR", "A-DR", "ACDR"}, 1, 0.5f, 0, 0, {"P", "P", "D", "D"}},
        {"constant_zero_emits_only_trimmed_polymorphic_middle_region", {"AAAQFCPQRRR", "AAAQDCDQRRR"}, 3, 1.0f, 0, 0, {"FCP", "DCD"}},
        {"one_polymorphic_middle_column_with_constant_three_has_length_four", {"AAACRRR", "AAADRRR"}, 3, 1.0f, 3, 0, {"CRRR", "DRRR"}},
        {"divergent_row_masking_happens_before_polymorphic_middle_filtering", {"ACDEFGK", "ACDEFGK", "APPPPPK"}, 1, 1.0f, 0, 5, {}},
        {"mask_zero_disables_divergent_row_masking", {"ACDEFGK", "ACDEFGK", "APPPPPK"}, 1, 1.0f, 0, 0, {"CDEFG", "CDEFG", "PPPPP"}},
        {"constant_tail_columns_failing_missingness_are_not_appended", {"ACT", "AD-", "AC-", "AD-"}, 1, 0.75f, 1, 0, {"C", "D", "C", "D"}},
    };
    for (const FilterCase& fc : cases) {
        const size_t n = fc.rows.size();
        const Extraction ex = test_extraction({{1, 2, fc.rows, {}}}, fc.k, 1, n);
        Candidate c;
        c.left = 1;
        c.right = 2;
        c.coverage = n;
        c.raw_len = fc.rows[0].size();
        for (uint32_t i = 0; i < n; i++) c.rows.push_back(i);
        const Alignment al = filter_groups(ex, {c}, fc.min_freq, fc.constant, fc.mask, SR6, 1);
        const std::string where = std::string("failed: group_filtering::") + fc.name;
        if (fc.want.empty()) {
            if (!al.partitions.empty()) return where;
            for (const auto& row : al.concat) if (!row.empty()) return where;
            continue;
        }
        const size_t len = fc.want[0].size();
        if (al.partitions.size() != 1 || al.partitions[0].start != 0 || al.partitions[0].end != len - 1 || al.partitions[0].len != len) return where;
        for (size_t s = 0; s < n; s++)
            if (std::string(al.concat[s].begin(), al.concat[s].end()) != fc.want[s]) return where;
    }
    {   // group_filtering.rs:414-481: majority_recoded_consensus and mask_rows_with_long_divergence
        const std::string m = "AAXDDBD--";   // rows AAX / DDB / D--
        const std::vector<int> got = majority_recoded_consensus(reinterpret_cast<const uint8_t*>(m.data()), 3, 3, SR6);
        if (got != std::vector<int>{1, 0, -1}) return "failed: group_filtering::majority_recoded_consensus_ignores_missing_invalid_and_ties_by_state_value";
        struct MaskCase { const char* name; std::string row; std::vector<int> cons; size_t mask; std::string want; };
        const std::vector<MaskCase> mc = {
            {"same_recoded_state_differences_are_not_masked", "PPPPP", {0, 0, 0, 0, 0}, 3, "PPPPP"},
            {"fewer_than_mask_consecutive_recoded_differences_are_not_masked", "ADDAA", {0, 0, 0, 0, 0}, 3, "ADDAA"},
            {"exactly_mask_consecutive_recoded_differences_mask_entire_row", "ADDDA", {0, 0, 0, 0, 0}, 3, "XXXXX"},
            {"missing_ambiguous_and_invalid_positions_break_difference_runs", "DDAD-DXBD", {0, 0, 0, 0, 0, 0, 0, 0, 0}, 3, "DDAD-DXBD"},
            {"consensus_positions_without_valid_recoded_state_break_difference_runs", "DDDDA", {0, 0, -1, 0, 0}, 3, "DDDDA"},
            {"mask_zero_returns_without_changing_rows", "DDDDD", {0, 0, 0, 0, 0}, 0, "DDDDD"},
        };
        for (const MaskCase& c : mc) {
            std::string row = c.row;
            mask_rows_with_long_divergence(reinterpret_cast<uint8_t*>(row.data()), 1, row.size(), c.cons, c.mask, SR6);
            if (row != c.want) return std::string("failed: group_filtering::") + c.name;
        }
    }
    auto keys_of = [](const std::vector<Candidate>& v) {
        std::vector<std::pair<uint64_t, uint64_t>> k;
        for (const Candidate& c : v) k.emplace_back(c.left, c.right);
        return k;
    };
    using Keys = std::vector<std::pair<uint64_t, uint64_t>>;
    {   // sort_nonoverlap_uses_shorter_overlap_k_when_k_less_than_11 (group_sorting.rs:486-521)
        const std::string b = "ACDEFGHIKLMNPQRSTVW";
        const Extraction ex = test_extraction({{10, 20, {b, b}, {}}, {30, 40, {b, b}, {}}}, 9, 2, 2);
        if (keys_of(sort_and_deduplicate_groups(ex, SR6)) != Keys{{10, 20}}) return "failed: group_sorting::sort_nonoverlap_uses_shorter_overlap_k_when_k_less_than_11";
    }
    {   // nonoverlap_filter_drops_weaker_shared_kmer_and_retains_empty_kmer_candidates (:540-558): coverage 2 beats 1
        const std::string shared = "ACDEFGHIKLMNPQRSTVWYA", xs(21, 'X');
        const Extraction ex = test_extraction({{20, 21, {shared}, {}}, {30, 31, {xs}, {}}, {10, 11, {shared, shared}, {}}}, 10, 1, 2);
        if (keys_of(sort_and_deduplicate_groups(ex, SR6)) != Keys{{10, 11}, {30, 31}}) return "failed: group_sorting::nonoverlap_filter_drops_weaker_shared_kmer";
    }
    {   // streaming_nonoverlap_preserves_full_path_overlap_semantics (:561-576), overlap k = 5 <=> k = 2
        const Extraction ex = test_extraction({{20, 21, {"ZZACDEFZZ", "ZZACDEFZZ", "ZZACDEFZZ"}, {}}, {40, 41, {"GGGGG"}, {}},
                                               {30, 31, {"XXXXX"}, {}}, {10, 11, {"ACDEF", "ACDEF", "ACDEF", "ACDEF"}, {}}}, 2, 1, 4);
        if (keys_of(sort_and_deduplicate_groups(ex, SR6)) != Keys{{10, 11}, {30, 31}, {40, 41}}) return "failed: group_sorting::streaming_nonoverlap_preserves_full_path_overlap_semantics";
    }
    {   // length_filter_rejects_ties_low_support_and_short_blocks (:579-603)
        const Extraction tie = test_extraction({{1, 2, {"ACDEF", "ACDEF", "ACDEFG", "ACDEFG"}, {}}}, 2, 2, 4);
        const Extraction low = test_extraction({{1, 2, {"ACDEF"}, {}}}, 2, 2, 4);
        const Extraction shrt = test_extraction({{1, 2, {"ACD", "ACD"}, {}}}, 2, 2, 2);
        if (!sort_and_deduplicate_groups(tie, SR6).empty() || !sort_and_deduplicate_groups(low, SR6).empty() ||
            !sort_and_deduplicate_groups(shrt, SR6).empty())
            return "failed: group_sorting::length_filter_rejects_ties_low_support_and_short_blocks";
    }
    {   // sort_retains_original_raw_anchor_key_without_refinement (:606-615)
        const Extraction ex = test_extraction({{31, 37, {"*AC**FGHIKL*", "*AC**DGHIKL*"}, {}}}, 2, 2, 2);
        const std::vector<Candidate> v = sort_and_deduplicate_groups(ex, SR6);
        if (v.size() != 1 || v[0].left != 31 || v[0].right != 37 || v[0].raw_len != 12) return "failed: group_sorting::sort_retains_original_raw_anchor_key_without_refinement";
    }
    {   // sort_nonoverlap_keeps_strongest_overlapping_group_deterministically (:618-637): equal strength, smaller key wins
        const std::string shared = "ACDEFGHIKLMNPQRSTVWYA", ys(21, 'Y');
        const Extraction ex = test_extraction({{50, 60, {shared, shared}, {}}, {10, 20, {shared, shared}, {}}, {30, 40, {ys}, {}}}, 10, 1, 2);
        if (keys_of(sort_and_deduplicate_groups(ex, SR6)) != Keys{{10, 20}, {30, 40}}) return "failed: group_sorting::sort_nonoverlap_keeps_strongest_overlapping_group_deterministically";
    }
    {   // consensus_rows_by_isolate_uses_arena_backed_observations (:526-537), seen through filter_groups: two
        // observations of one isolate collapse into one row, a disagreement becomes a sticky X
        const Extraction ex = test_extraction({{1, 2, {"ACDA", "ACDA", "AEDA", "AEHA"}, {0, 0, 1, 1}}}, 1, 1, 2);
        Candidate c;
        c.left = 1; c.right = 2; c.coverage = 2; c.raw_len = 4;
        c.rows = {0, 1, 2, 3};
        const Alignment al = filter_groups(ex, {c}, 0.6f, 0, 0, SR6, 1);   // columns: C/E polymorphic, D/X
        if (al.partitions.size() != 1 || std::string(al.concat[0].begin(), al.concat[0].end()) != "C" ||
            std::string(al.concat[1].begin(), al.concat[1].end()) != "E")
            return "failed: group_sorting::consensus_rows_by_isolate";
    }
    return std::string();
}

}  // namespace kh
