// TEST INFRASTRUCTURE — the reference's own unit tests of the host logic, restated on the oracle's functions:
// /root/reference/src/group_extraction.rs:666-748, src/group_sorting.rs:438-659, src/group_filtering.rs:413-693.
// They pin the oracle where the end-to-end goldens are blind (ties, masking, missingness order, ...).
// orc_self_test (capi.cpp) returns the name of the first failing reference test, or an empty string.
#include "kamino_oracle.hpp"

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

namespace kmo {

namespace {

using Rows = std::vector<std::string>;

BlockObs obs(BlockArena& arena, uint16_t sid, const std::string& aa) {   // tests' obs(): protein_id = sid
    return BlockObs{sid, uint32_t(sid), arena.push(reinterpret_cast<const uint8_t*>(aa.data()), aa.size())};
}

SortedGroups sorted_with_k(const Rows& rows, size_t k) {                 // group_filtering.rs:397-413
    SortedGroups s;
    RawDirectCandidate c;
    c.key = {1, 2};
    c.coverage = rows.size();
    c.raw_len = rows[0].size();
    for (size_t sid = 0; sid < rows.size(); sid++) c.selected.push_back(obs(s.arena, uint16_t(sid), rows[sid]));
    s.raw_candidates.push_back(std::move(c));
    for (size_t sid = 0; sid < rows.size(); sid++) s.species_names.push_back("species_" + std::to_string(sid));
    s.protein_names.assign(rows.size(), "protein");
    s.k = k;
    s.min_needed = 1;
    s.n_species = rows.size();
    s.scheme = SR6;
    return s;
}

RawDirectCandidate raw_candidate(BlockArena& arena, AnchorPair key, size_t coverage, size_t raw_len, const Rows& rows) {   // group_sorting.rs:419-436
    RawDirectCandidate c;
    c.key = key;
    c.coverage = coverage;
    c.raw_len = raw_len;
    for (size_t sid = 0; sid < rows.size(); sid++) c.selected.push_back(obs(arena, uint16_t(sid), rows[sid]));
    return c;
}

std::vector<std::vector<uint8_t>> to_rows(const Rows& r) {
    std::vector<std::vector<uint8_t>> out;
    for (const std::string& s : r) out.emplace_back(s.begin(), s.end());
    return out;
}

std::vector<std::pair<uint64_t, uint64_t>> keys_of(const std::vector<RawDirectCandidate>& v) {
    std::vector<std::pair<uint64_t, uint64_t>> k;
    for (const auto& c : v) k.emplace_back(c.key.left, c.key.right);
    return k;
}

struct FilterCase { const char* name; Rows rows; size_t k; float min_freq; size_t constant, mask; Rows want; };

}  // namespace

std::string self_test() {
#define ORC_CHECK(name, ...) do { if (!(__VA_ARGS__)) return std::string(name); } while (0)
    using Keys = std::vector<std::pair<uint64_t, uint64_t>>;
    // ---- group_extraction.rs:666-708
    ORC_CHECK("group_extraction::truncate_protein_name_keeps_record_id_and_removes_organism_suffix",
              truncate_protein_name("WP_012345.1 hypothetical protein [Escherichia coli]") == "WP_012345.1 hypothetical protein" &&
              truncate_protein_name("WP_012345.1 hypothetical protein") == "WP_012345.1 hypothetical protein");
    // not a reference test: str::trim / trim_end follow Unicode White_Space (NBSP, NEL, EN QUAD, IDEOGRAPHIC SPACE ...)
    ORC_CHECK("oracle::unicode_white_space_in_str_trim",
              truncate_protein_name("\xc2\xa0id kinase\xe2\x80\x80 [Homo]\xe3\x80\x80") == "id kinase" &&
              rs_trim("\xc2\x85 x\xe2\x80\xa8") == "x" && rs_trim("caf\xc3\xa9") == "caf\xc3\xa9" &&
              rs_ws_len(std::string("\xe2\x81\x9f"), 0) == 3 && rs_ws_len(std::string("\xe2\x80\x8b"), 0) == 0);
    {
        BlockArena arena;
        const BlockAa a = arena.push(reinterpret_cast<const uint8_t*>("ACD"), 3), b = arena.push(reinterpret_cast<const uint8_t*>("EFGH"), 4);
        ORC_CHECK("group_extraction::block_arena_push_and_get_preserve_exact_bytes",
                  a.len == 3 && b.len == 4 && std::string(arena.get(a), arena.get(a) + 3) == "ACD" &&
                  std::string(arena.get(b), arena.get(b) + 4) == "EFGH");
    }
    {   // pending_species_merge_is_deterministic_when_results_arrive_out_of_order (:711-748): results are merged in
        // species order whatever order they arrive in
        auto extraction = [](size_t sid, const std::string& protein, const std::string& aa) {
            SpeciesExtraction e;
            e.sid = sid;
            e.groups[{7, 11}] = {BlockObs{uint16_t(sid), 0, e.arena.push(reinterpret_cast<const uint8_t*>(aa.data()), aa.size())}};
            e.protein_names = {protein};
            return e;
        };
        RawDirectGroups groups;
        BlockArena arena;
        std::vector<std::string> names;
        std::vector<SpeciesExtraction> arrivals;
        arrivals.push_back(extraction(1, "protein_1", "BBBB"));
        arrivals.push_back(extraction(0, "protein_0", "AAAA"));
        arrivals.push_back(extraction(2, "protein_2", "CCCC"));
        std::vector<std::pair<size_t, SpeciesExtraction>> pending;
        size_t next_sid = 0;
        for (SpeciesExtraction& a : arrivals) {
            pending.emplace_back(a.sid, std::move(a));
            for (bool merged = true; merged;) {
                merged = false;
                for (size_t i = 0; i < pending.size(); i++)
                    if (pending[i].first == next_sid) {
                        merge_species_extraction(std::move(pending[i].second), groups, arena, names);
                        pending.erase(pending.begin() + i);
                        next_sid++;
                        merged = true;
                        break;
                    }
            }
        }
        const std::vector<BlockObs>& o = groups[{7, 11}];
        ORC_CHECK("group_extraction::pending_species_merge_is_deterministic_when_results_arrive_out_of_order",
                  names == std::vector<std::string>{"protein_0", "protein_1", "protein_2"} && o.size() == 3 && o[0].sid == 0 && o[1].sid == 1 &&
                  o[2].sid == 2 && o[0].protein_id == 0 && o[1].protein_id == 1 && o[2].protein_id == 2 &&
                  std::string(arena.get(o[0].aa), arena.get(o[0].aa) + 4) == "AAAA" && std::string(arena.get(o[1].aa), arena.get(o[1].aa) + 4) == "BBBB" &&
                  std::string(arena.get(o[2].aa), arena.get(o[2].aa) + 4) == "CCCC");
    }
    // ---- group_sorting.rs:438-659
    ORC_CHECK("group_sorting::direct_overlap_k_uses_minimum_block_length_below_k_11",
              direct_overlap_k(1) == 3 && direct_overlap_k(9) == 19 && direct_overlap_k(10) == 21 && direct_overlap_k(20) == 21);
    {
        const std::string blk = "ACDEFGHIKLMNPQRSTVW";
        RawExtractedGroups raw;
        raw.groups[{10, 20}] = {obs(raw.arena, 0, blk), obs(raw.arena, 1, blk)};
        raw.groups[{30, 40}] = {obs(raw.arena, 0, blk), obs(raw.arena, 1, blk)};
        raw.species_names = {"species_0", "species_1"};
        raw.protein_names = {"protein_0", "protein_1"};
        raw.k = 9;
        raw.min_needed = 2;
        raw.n_species = 2;
        const SortedGroups s = sort_and_deduplicate_groups(std::move(raw), SR6);
        ORC_CHECK("group_sorting::sort_nonoverlap_uses_shorter_overlap_k_when_k_less_than_11", keys_of(s.raw_candidates) == Keys{{10, 20}});
    }
    {
        BlockArena arena;
        const std::vector<BlockObs> o = {obs(arena, 0, "ACD"), obs(arena, 0, "ACD"), obs(arena, 1, "ACE")};
        ORC_CHECK("group_sorting::consensus_rows_by_isolate_uses_arena_backed_observations",
                  consensus_rows_by_isolate(o, arena, 2, 3) == to_rows({"ACD", "ACE"}));
    }
    {
        BlockArena arena;
        const std::string shared = "ACDEFGHIKLMNPQRSTVWYA", xs(21, 'X');
        std::vector<RawDirectCandidate> c = {raw_candidate(arena, {20, 21}, 1, 21, {shared}), raw_candidate(arena, {30, 31}, 1, 21, {xs}),
                                             raw_candidate(arena, {10, 11}, 2, 21, {shared})};
        ORC_CHECK("group_sorting::nonoverlap_filter_drops_weaker_shared_kmer_and_retains_empty_kmer_candidates",
                  keys_of(retain_nonoverlapping_raw_candidates(std::move(c), arena, 21, SR6)) == Keys{{10, 11}, {30, 31}});
    }
    {
        BlockArena arena;
        std::vector<RawDirectCandidate> c = {raw_candidate(arena, {20, 21}, 3, 9, {"ZZACDEFZZ"}), raw_candidate(arena, {40, 41}, 1, 5, {"GGGGG"}),
                                             raw_candidate(arena, {30, 31}, 1, 5, {"XXXXX"}), raw_candidate(arena, {10, 11}, 4, 5, {"ACDEF"})};
        ORC_CHECK("group_sorting::streaming_nonoverlap_preserves_full_path_overlap_semantics",
                  keys_of(retain_nonoverlapping_raw_candidates(std::move(c), arena, 5, SR6)) == Keys{{10, 11}, {30, 31}, {40, 41}});
    }
    {
        BlockArena arena;
        RawDirectCandidate out;
        const std::vector<BlockObs> tie = {obs(arena, 0, "ACDEF"), obs(arena, 1, "ACDEF"), obs(arena, 2, "ACDEFG"), obs(arena, 3, "ACDEFG")};
        const std::vector<BlockObs> low = {obs(arena, 0, "ACDEF")};
        const std::vector<BlockObs> shrt = {obs(arena, 0, "ACD"), obs(arena, 1, "ACD")};
        ORC_CHECK("group_sorting::length_filter_rejects_ties_low_support_and_short_blocks",
                  !select_unique_best_supported_length({1, 2}, tie, 4, 2, 2, out) && !select_unique_best_supported_length({1, 2}, low, 4, 2, 2, out) &&
                  !select_unique_best_supported_length({1, 2}, shrt, 2, 2, 2, out));
    }
    {
        const Rows rows = {"*AC**FGHIKL*", "*AC**DGHIKL*"};
        RawExtractedGroups raw;
        raw.groups[{31, 37}] = {obs(raw.arena, 0, rows[0]), obs(raw.arena, 1, rows[1])};
        raw.species_names = {"species_0", "species_1"};
        raw.protein_names = {"protein", "protein"};
        raw.k = 2;
        raw.min_needed = 2;
        raw.n_species = 2;
        const SortedGroups s = sort_and_deduplicate_groups(std::move(raw), SR6);
        ORC_CHECK("group_sorting::sort_retains_original_raw_anchor_key_without_refinement",
                  s.raw_candidates.size() == 1 && s.raw_candidates[0].key == (AnchorPair{31, 37}) && s.raw_candidates[0].raw_len == 12);
    }
    {
        BlockArena arena;
        const std::string shared = "ACDEFGHIKLMNPQRSTVWYA", ys(21, 'Y');
        std::vector<RawDirectCandidate> c = {raw_candidate(arena, {50, 60}, 2, 21, {shared}), raw_candidate(arena, {10, 20}, 2, 21, {shared}),
                                             raw_candidate(arena, {30, 40}, 1, 21, {ys})};
        ORC_CHECK("group_sorting::sort_nonoverlap_keeps_strongest_overlapping_group_deterministically",
                  keys_of(retain_nonoverlapping_raw_candidates(std::move(c), arena, 21, SR6)) == Keys{{10, 20}, {30, 40}});
    }
    // ---- group_filtering.rs:413-693
    ORC_CHECK("group_filtering::majority_recoded_consensus_ignores_missing_invalid_and_ties_by_state_value",
              majority_recoded_consensus(to_rows({"AAX", "DDB", "D--"}), 3, SR6) == (std::vector<int>{1, 0, -1}));
    {
        struct MaskCase { const char* name; std::string row; std::vector<int> cons; size_t mask; std::string want; };
        const std::vector<MaskCase> mc = {
            {"group_filtering::same_recoded_state_differences_are_not_masked", "PPPPP", {0, 0, 0, 0, 0}, 3, "PPPPP"},
            {"group_filtering::fewer_than_mask_consecutive_recoded_differences_are_not_masked", "ADDAA", {0, 0, 0, 0, 0}, 3, "ADDAA"},
            {"group_filtering::exactly_mask_consecutive_recoded_differences_mask_entire_row", "ADDDA", {0, 0, 0, 0, 0}, 3, "XXXXX"},
            {"group_filtering::missing_ambiguous_and_invalid_positions_break_difference_runs", "DDAD-DXBD", {0, 0, 0, 0, 0, 0, 0, 0, 0}, 3, "DDAD-DXBD"},
            {"group_filtering::consensus_positions_without_valid_recoded_state_break_difference_runs", "DDDDA", {0, 0, -1, 0, 0}, 3, "DDDDA"},
            {"group_filtering::mask_zero_returns_without_changing_rows", "DDDDD", {0, 0, 0, 0, 0}, 0, "DDDDD"},
        };
        for (const MaskCase& c : mc) {
            auto rows = to_rows({c.row});
            mask_rows_with_long_divergence(rows, c.cons, c.mask, SR6);
            ORC_CHECK(c.name, rows == to_rows({c.want}));
        }
    }
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
        const AlignmentResult r = filter_groups(sorted_with_k(fc.rows, fc.k), fc.min_freq, fc.constant, fc.mask, 1);
        const std::string name = std::string("group_filtering::") + fc.name;
        if (fc.want.empty()) {
            ORC_CHECK(name, r.partitions.empty() && std::all_of(r.concat.begin(), r.concat.end(), [](const auto& v) { return v.empty(); }));
            continue;
        }
        const size_t len = fc.want[0].size();
        ORC_CHECK(name, r.partitions.size() == 1 && r.partitions[0].start == 0 && r.partitions[0].end == len - 1 && r.partitions[0].len == len &&
                            r.concat == to_rows(fc.want));
    }
    {   // final_middle_column_before_constants_is_polymorphic (:590-608)
        const AlignmentResult r = filter_groups(sorted_with_k({"AAAQFCPQRRR", "AAAQDCDQRRR"}, 3), 1.0f, 3, 0, 1);
        const size_t middle_len = r.partitions[0].len - 3;
        std::vector<std::vector<uint8_t>> trimmed;
        for (const auto& row : r.concat) trimmed.emplace_back(row.begin(), row.begin() + middle_len);
        ORC_CHECK("group_filtering::final_middle_column_before_constants_is_polymorphic",
                  is_polymorphic_column(trimmed, middle_len - 1, SR6) && r.concat[0][middle_len - 1] == 'P' && r.concat[1][middle_len - 1] == 'D');
    }
#undef ORC_CHECK
    return std::string();
}

}  // namespace kmo
