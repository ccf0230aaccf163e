// pipeline.cpp — CPU ORACLE (test infrastructure): output writers, run_with_args, CLI.
// Restates /root/reference/src/output.rs, src/lib.rs:206-298 and the clap surface of
// src/lib.rs:119-175.  `--genomes` (src/protein_prediction.rs) is NOT restated: out of scope
// (SURVEY.md §8f-4); the flag is rejected with an explicit error.
#include "kamino_oracle.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace kmo {

// PathBuf::push + set_extension for a file name without directory components.
static std::string build_path(const std::string& base, const char* suffix, const char* ext) {  // output.rs:12-22
    std::string fname = rs_file_name(base);
    std::string stem = rs_file_stem(fname);
    if (stem.empty()) stem = fname.empty() ? "output" : fname;
    std::string name = stem + suffix;
    bool hp = false;
    std::string parent = rs_parent(base, &hp);
    // set_extension: replace everything after the last '.' of the file name (if any)
    std::string st = rs_file_stem(name);
    std::string final_name = st + "." + ext;
    if (parent.empty()) return final_name;
    if (parent == "/") return "/" + final_name;
    return parent + "/" + final_name;
}

OutputPaths output_paths(const std::string& out_base) {  // output.rs:9-30
    return OutputPaths{build_path(out_base, "_alignment", "fas"), build_path(out_base, "_missing", "tsv"),
                       build_path(out_base, "_partitions", "tsv"), build_path(out_base, "_NJ", "tree")};
}

static FILE* create_file(const std::string& p) {
    FILE* f = std::fopen(p.c_str(), "wb");
    if (!f) throw std::runtime_error("create \"" + p + "\": cannot open for writing");
    return f;
}

void write_outputs(const std::string& out_base, const AlignmentResult& res, bool nj, size_t threads,
                   size_t* total_len_out, double* pct_missing_out) {  // output.rs:32-95
    OutputPaths paths = output_paths(out_base);
    const auto& species = res.species_names;
    const auto& concat = res.concat;
    FILE* w = create_file(paths.fas);
    size_t total_len = concat.empty() ? 0 : concat[0].size();
    size_t total_pos = total_len * species.size();
    size_t total_missing = 0;
    for (auto& s : concat)
        for (uint8_t b : s)
            if (b == '-' || b == 'X') total_missing++;
    double miss = total_pos == 0 ? 0.0 : 100.0 * double(total_missing) / double(total_pos);
    for (size_t sid = 0; sid < species.size(); sid++) {
        std::fprintf(w, ">%s\n", species[sid].c_str());
        const auto& row = concat[sid];
        for (size_t off = 0; off < row.size(); off += 60) {
            size_t n = std::min<size_t>(60, row.size() - off);
            std::fwrite(row.data() + off, 1, n, w);
            std::fputc('\n', w);
        }
    }
    std::fclose(w);
    FILE* tw = create_file(paths.tsv);
    std::fprintf(tw, "species\tpct_missing\n");
    for (size_t sid = 0; sid < species.size(); sid++) {
        size_t len = std::max<size_t>(concat[sid].size(), 1);
        size_t m = 0;
        for (uint8_t b : concat[sid])
            if (b == '-' || b == 'X') m++;
        std::fprintf(tw, "%s\t%.1f\n", species[sid].c_str(), 100.0 * double(m) / double(len));
    }
    std::fclose(tw);
    FILE* pw = create_file(paths.parts);
    std::fprintf(pw, "start_pos\tend_pos\tlength\tconsensus protein name\n");
    for (size_t i = 0; i < res.partitions.size() && i < res.partition_names.size(); i++)
        std::fprintf(pw, "%zu\t%zu\t%zu\t%s\n", res.partitions[i].start, res.partitions[i].end,
                     res.partitions[i].len, res.partition_names[i].c_str());
    std::fclose(pw);
    if (nj) {
        std::string tree = nj_tree_newick(species, concat, threads);
        FILE* t = create_file(paths.tree);
        std::fprintf(t, "%s\n", tree.c_str());
        std::fclose(t);
    }
    if (total_len_out) *total_len_out = total_len;
    if (pct_missing_out) *pct_missing_out = miss;
}

void run_with_args(const Args& args) {  // lib.rs:206-298
    size_t default_k = 13, max_k = 21;
    size_t k = args.k < 0 ? default_k : size_t(args.k);
    size_t constant = args.constant < 0 ? std::min<size_t>(3, k) : size_t(args.constant);
    std::fprintf(stderr, "kamino 2.0.0 (C++ oracle restatement)\n");
    if (!(args.min_freq >= 0.6f && args.min_freq <= 1.0f))
        throw std::runtime_error("min_freq must be between 0.6 and 1.0");
    if (args.threads == 0) throw std::runtime_error("threads must be >=1");
    if (k < 1 || k > max_k) throw std::runtime_error("invalid k");
    if (constant > k) throw std::runtime_error("constant <= k");
    std::vector<SpeciesInput> inputs;
    if (args.has_input_file) {
        auto v = collect_species_inputs_from_table(args.input_file);
        inputs.insert(inputs.end(), v.begin(), v.end());
    }
    if (args.has_input) {
        auto v = collect_species_inputs_from_dir(args.input);
        inputs.insert(inputs.end(), v.begin(), v.end());
    }
    std::fprintf(stderr, "# process input files\n");
    if (args.genomes)
        throw std::runtime_error("--genomes (protein prediction) is outside the oracle's scope");
    if (inputs.empty()) throw std::runtime_error("At least one input source with files must be provided.");
    RawExtractedGroups raw = extract_groups(inputs, k, args.min_freq, args.length_middle, args.recode,
                                            args.threads);
    std::fprintf(stderr, "# analyse variant groups\n . raw variant groups: %zu\n", raw.groups.size());
    SortedGroups sorted = sort_and_deduplicate_groups(std::move(raw), args.recode);
    std::fprintf(stderr, " . sorted variant groups: %zu\n", sorted.raw_candidates.size());
    AlignmentResult res = filter_groups(std::move(sorted), args.min_freq, constant, args.mask, args.threads);
    std::fprintf(stderr, " . filtered variant groups: %zu\n", res.partitions.size());
    size_t alen = 0;
    double amiss = 0;
    write_outputs(args.output, res, args.nj, args.threads, &alen, &amiss);
    std::fprintf(stderr, "# output files\n . alignment: length=%zu missing=%.1f%%\n", alen, amiss);
    if (args.nj) std::fprintf(stderr, " . NJ tree\n");
}

static bool parse_scheme(const std::string& v, RecodeScheme& out) {
    if (v == "dayhoff6") { out = Dayhoff6; return true; }
    if (v == "sr6") { out = SR6; return true; }
    if (v == "kgb6") { out = KGB6; return true; }
    return false;
}

int cli_main(int argc, char** argv) {  // main.rs:5-9 + clap derive (lib.rs:119-175)
    Args a;
    try {
        for (int i = 1; i < argc; i++) {
            std::string arg = argv[i], val;
            bool has_val = false;
            if (arg.rfind("--", 0) == 0) {
                size_t eq = arg.find('=');
                if (eq != std::string::npos) { val = arg.substr(eq + 1); arg = arg.substr(0, eq); has_val = true; }
            }
            auto need = [&]() -> std::string {
                if (has_val) return val;
                if (i + 1 >= argc) throw std::runtime_error("missing value for " + arg);
                return argv[++i];
            };
            if (arg == "-i" || arg == "--input-directory") { a.input = need(); a.has_input = true; }
            else if (arg == "-I" || arg == "--input-file") { a.input_file = need(); a.has_input_file = true; }
            else if (arg == "--genomes") a.genomes = true;
            else if (arg == "-k" || arg == "--k") a.k = std::stol(need());
            else if (arg == "-f" || arg == "--min-freq") a.min_freq = std::strtof(need().c_str(), nullptr);
            else if (arg == "-o" || arg == "--output") a.output = need();
            else if (arg == "-c" || arg == "--constant") a.constant = std::stol(need());
            else if (arg == "-l" || arg == "--length-middle") a.length_middle = size_t(std::stoul(need()));
            else if (arg == "-m" || arg == "--mask") a.mask = size_t(std::stoul(need()));
            else if (arg == "-t" || arg == "--threads") a.threads = size_t(std::stoul(need()));
            else if (arg == "-r" || arg == "--recode") {
                if (!parse_scheme(need(), a.recode)) throw std::runtime_error("invalid value for --recode");
            } else if (arg == "--nj" || arg == "--NJ") a.nj = true;
            else if (arg == "-v" || arg == "-V" || arg == "--version") { std::printf("kamino 2.0.0\n"); return 0; }
            else throw std::runtime_error("unexpected argument '" + arg + "'");
        }
        if (!a.has_input && !a.has_input_file)
            throw std::runtime_error("one of --input-directory / --input-file is required");
        run_with_args(a);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}

}  // namespace kmo
