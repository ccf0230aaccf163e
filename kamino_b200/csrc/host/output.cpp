// output.cpp — output writers, NJ tree and the CLI driver of the B200 pipeline.
// Equivalent to /root/reference/src/output.rs, src/phylo.rs (distance counts come from the GPU:
// kmn_pair_counts; the f64 correction and the O(n^3) NJ reduction stay on the host so every f64
// operation happens in the reference's order — compile with -ffp-contract=off) and src/lib.rs:206-298.
#include "../../../include/kamino_b200.h"
#include "kamino_host.hpp"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <limits>
#include <ctime>
#include <set>

namespace fs = std::filesystem;

namespace kh {

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
}
StageTimer::StageTimer(const char* w) : what(w), t0(now_s()) {}
StageTimer::~StageTimer() {
    if (std::getenv("KAMINO_TIMING")) std::fprintf(stderr, "[time] %s: %.3f s\n", what, now_s() - t0);
}

namespace {

// <parent>/<file_stem(prefix)><suffix>.<ext> with Rust's file_stem / set_extension rules (output.rs:9-30)
std::string derive_path(const std::string& base, const char* suffix, const char* ext) {
    const fs::path b(base);
    std::string stem = b.stem().string();
    if (stem.empty()) stem = b.filename().string();
    if (stem.empty()) stem = "output";
    fs::path name(stem + suffix);
    name.replace_extension(ext);          // "a.b_alignment" -> "a.fas", exactly like PathBuf::set_extension
    return (b.parent_path() / name).string();
}

FILE* create(const std::string& p) {
    FILE* f = std::fopen(p.c_str(), "wb");
    if (!f) throw std::runtime_error("create \"" + p + "\"");
    return f;
}

const double LG_PI[20] = {   // phylo.rs:27-31
    0.079066, 0.055941, 0.041977, 0.053052, 0.012937, 0.040767, 0.071586, 0.057337, 0.022355, 0.062157,
    0.099081, 0.064600, 0.022951, 0.042302, 0.044040, 0.061197, 0.053287, 0.012777, 0.027843, 0.070200};

double f81_c() {   // phylo.rs:34-43
    double s = 0.0;
    for (double pi : LG_PI) s += pi * pi;
    return 1.0 - s;
}

struct TreeNode { int a = -1, b = -1; double la = 0, lb = 0; };

std::string quoted(const std::string& n) {   // phylo.rs:235-245
    if (n.find_first_of(" :(),;") == std::string::npos) return n;
    std::string q = "'";
    for (char c : n) q += c == '\'' ? std::string("''") : std::string(1, c);
    return q + "'";
}

void emit(const std::vector<TreeNode>& t, const std::vector<std::string>& names, size_t id, std::string& out) {   // phylo.rs:253-265
    if (id < names.size()) { out += quoted(names[id]); return; }
    char buf[64];
    out += "(";
    emit(t, names, size_t(t[id].a), out);
    std::snprintf(buf, sizeof buf, ":%.6f,", t[id].la);
    out += buf;
    emit(t, names, size_t(t[id].b), out);
    std::snprintf(buf, sizeof buf, ":%.6f)", t[id].lb);
    out += buf;
}

// phylo.rs:274-301: counts on the GPU(s) of the run, distances + classic NJ on the host
std::string nj_newick(const std::vector<std::string>& names, const std::vector<std::vector<uint8_t>>& seqs, GpuSet& gpus, size_t threads) {
    const size_t n = names.size();
    if (n != seqs.size()) throw std::runtime_error("names and sequences length mismatch");
    if (n < 2) throw std::runtime_error("need at least 2 sequences");
    std::set<size_t> lens;
    for (auto& s : seqs) lens.insert(s.size());
    if (lens.size() != 1) throw std::runtime_error("all sequences must have identical lengths");
    const size_t L = *lens.begin();
    static const char order[] = "ARNDCQEGHILKMFPSTWYV";   // phylo.rs:47-79
    uint8_t map[256];
    std::memset(map, 20, sizeof map);
    for (int i = 0; i < 20; i++) map[uint8_t(order[i])] = uint8_t(i);
    // map_sequence (phylo.rs:75-79) on the host threads, straight into page-locked memory
    if (gpus.size() == 0) throw std::runtime_error("GPU front end: no context for the pair counts");
    struct PinnedBytes {
        void* p = nullptr;
        ~PinnedBytes() { kmn_host_free(p); }
    } pin;
    if (kmn_host_alloc(&pin.p, n * L) != 0) throw std::runtime_error(std::string("GPU front end: ") + kmn_last_error());
    uint8_t* mapped = static_cast<uint8_t*>(pin.p);
    {
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t i; (i = next.fetch_add(1)) < n;)
                for (size_t c = 0; c < L; c++) mapped[i * L + c] = map[seqs[i][c]];
        };
        std::vector<std::thread> pool;
        for (size_t t = 0; t < std::max<size_t>(1, std::min(threads, n)); t++) pool.emplace_back(work);
        for (auto& t : pool) t.join();
    }
    std::vector<uint32_t> valid(n * n, 0), mism(n * n, 0);
    // the rayon loop over rows i of compute_distance_matrix (phylo.rs:124-138), sharded over the run's GPUs as
    // contiguous row blocks aligned to the kernel's 64-row tiles and balanced by upper-triangle tiles
    const size_t G = gpus.size();
    const size_t n_tiles = (n + 63) / 64, total = n_tiles * (n_tiles + 1) / 2;
    std::vector<std::pair<uint32_t, uint32_t>> blocks;
    {
        size_t t = 0, done = 0;
        for (size_t r = 0; r < G; r++) {
            const size_t target = total * (r + 1) / G, t0 = t;
            auto dist = [](size_t a, size_t b) { return a > b ? a - b : b - a; };
            while (t < n_tiles && (r == G - 1 || dist(done + (n_tiles - t), target) <= dist(done, target))) {
                done += n_tiles - t;
                t++;
            }
            blocks.emplace_back(uint32_t(std::min(t0 * 64, n)), uint32_t(std::min(t * 64, n)));
        }
    }
    std::vector<std::string> errs(G);
    {
        std::vector<std::thread> pool;
        for (size_t r = 0; r < G; r++)
            pool.emplace_back([&, r] {
                const uint32_t b0 = blocks[r].first, b1 = blocks[r].second;
                if (b1 <= b0) return;
                if (kmn_pair_counts_rows(gpus.ctx[r], mapped, uint32_t(n), L, b0, b1, valid.data() + size_t(b0) * n,
                                         mism.data() + size_t(b0) * n) != 0)
                    errs[r] = kmn_last_error();
            });
        for (auto& t : pool) t.join();
    }
    for (auto& e : errs) if (!e.empty()) throw std::runtime_error("GPU front end: " + e);

    const size_t dim = 2 * n - 1;
    std::vector<double> d(dim * dim, 0.0);
    const double C = f81_c();
    for (size_t i = 0; i < n; i++)
        for (size_t j = i + 1; j < n; j++) {   // phylo.rs:96-105
            const uint32_t v = valid[i * n + j], m = mism[i * n + j];
            double dist = 0.0;
            if (v != 0) {
                const double p = double(m) / double(v);
                double pc = p / C;
                if (pc >= 1.0) pc = 0.999999999;
                dist = -std::log(1.0 - pc);
            }
            d[i * dim + j] = dist;
            d[j * dim + i] = dist;
        }
    // phylo.rs:145-232
    std::vector<TreeNode> tree(n);
    std::vector<size_t> active(n);
    for (size_t i = 0; i < n; i++) active[i] = i;
    std::vector<double> r(dim);
    while (active.size() > 2) {
        const size_t m = active.size();
        std::fill(r.begin(), r.end(), 0.0);
        for (size_t i : active) {
            double sum = 0.0;
            for (size_t j : active) if (i != j) sum += d[i * dim + j];
            r[i] = sum;
        }
        size_t bi = active[0], bj = active[1];
        double best = std::numeric_limits<double>::infinity();
        for (size_t a = 0; a < m; a++)
            for (size_t b = a + 1; b < m; b++) {
                const size_t i = active[a], j = active[b];
                const double q = double(m - 2) * d[i * dim + j] - r[i] - r[j];
                if (q < best) { best = q; bi = i; bj = j; }
            }
        const double dij = d[bi * dim + bj];
        double li = 0.5 * dij + (r[bi] - r[bj]) / (2.0 * double(m - 2));
        double lj = dij - li;
        if (li < 0.0) li = 0.0;
        if (lj < 0.0) lj = 0.0;
        const size_t u = tree.size();
        TreeNode node;
        node.a = int(bi); node.b = int(bj); node.la = li; node.lb = lj;
        tree.push_back(node);
        for (size_t x : active) {
            if (x == bi || x == bj) continue;
            const double du = 0.5 * (d[bi * dim + x] + d[bj * dim + x] - dij);
            d[u * dim + x] = du;
            d[x * dim + u] = du;
        }
        std::vector<size_t> rest;
        for (size_t x : active) if (x != bi && x != bj) rest.push_back(x);
        rest.push_back(u);
        active.swap(rest);
    }
    double len = d[active[0] * dim + active[1]] * 0.5;
    if (len < 0.0) len = 0.0;
    TreeNode root;
    root.a = int(active[0]); root.b = int(active[1]); root.la = len; root.lb = len;
    tree.push_back(root);
    std::string out;
    emit(tree, names, tree.size() - 1, out);
    return out + ";";
}

}  // namespace

void write_outputs(const std::string& out_base, const Alignment& al, bool nj, GpuSet& gpus, size_t threads, size_t* total_len,
                   double* pct_missing) {   // output.rs:32-95
    const auto& sp = al.species_names;
    const size_t len = al.concat.empty() ? 0 : al.concat[0].size();
    size_t missing = 0;
    for (auto& row : al.concat) for (uint8_t b : row) missing += (b == '-' || b == 'X');
    const size_t cells = len * sp.size();
    if (total_len) *total_len = len;
    if (pct_missing) *pct_missing = cells == 0 ? 0.0 : 100.0 * double(missing) / double(cells);
    FILE* f = create(derive_path(out_base, "_alignment", "fas"));
    for (size_t s = 0; s < sp.size(); s++) {
        std::fprintf(f, ">%s\n", sp[s].c_str());
        for (size_t off = 0; off < al.concat[s].size(); off += 60) {   // 60-column FASTA
            std::fwrite(al.concat[s].data() + off, 1, std::min<size_t>(60, al.concat[s].size() - off), f);
            std::fputc('\n', f);
        }
    }
    std::fclose(f);
    f = create(derive_path(out_base, "_missing", "tsv"));
    std::fprintf(f, "species\tpct_missing\n");
    for (size_t s = 0; s < sp.size(); s++) {
        size_t m = 0;
        for (uint8_t b : al.concat[s]) m += (b == '-' || b == 'X');
        std::fprintf(f, "%s\t%.1f\n", sp[s].c_str(), 100.0 * double(m) / double(std::max<size_t>(al.concat[s].size(), 1)));
    }
    std::fclose(f);
    f = create(derive_path(out_base, "_partitions", "tsv"));
    std::fprintf(f, "start_pos\tend_pos\tlength\tconsensus protein name\n");
    for (size_t i = 0; i < al.partitions.size(); i++)
        std::fprintf(f, "%zu\t%zu\t%zu\t%s\n", al.partitions[i].start, al.partitions[i].end, al.partitions[i].len,
                     al.partition_names[i].c_str());
    std::fclose(f);
    if (nj) {
        const std::string tree = nj_newick(sp, al.concat, gpus, threads);
        f = create(derive_path(out_base, "_NJ", "tree"));
        std::fprintf(f, "%s\n", tree.c_str());
        std::fclose(f);
    }
}

void run(const Options& o) {   // lib.rs:206-298
    const size_t k = o.k < 0 ? 13 : size_t(o.k);
    const size_t constant = o.constant < 0 ? std::min<size_t>(3, k) : size_t(o.constant);
    if (!o.quiet) std::fprintf(stderr, "kamino-b200 (GPU front end; kamino 2.0.0 semantics)\n");
    if (!(o.min_freq >= 0.6f && o.min_freq <= 1.0f)) throw std::runtime_error("min_freq must be between 0.6 and 1.0");
    if (o.threads == 0) throw std::runtime_error("threads must be >=1");
    if (k < 1 || k > 21) throw std::runtime_error("invalid k");
    if (constant > k) throw std::runtime_error("constant <= k");
    std::vector<SpeciesFile> inputs;   // table entries first, then directory entries (lib.rs:221-227)
    if (o.has_file) { auto v = inputs_from_table(o.input_file); inputs.insert(inputs.end(), v.begin(), v.end()); }
    if (o.has_dir) { auto v = inputs_from_directory(o.input_dir); inputs.insert(inputs.end(), v.begin(), v.end()); }
    if (!o.quiet) std::fprintf(stderr, "# process input files\n");
    // Genome mode (lib.rs:235-248): predicted proteomes go to a temporary directory next to the output prefix,
    // removed when the run ends (TempDir's drop, lib.rs:296)
    struct TempDir {
        std::string path;
        ~TempDir() {
            if (path.empty()) return;
            std::error_code ec;
            std::filesystem::remove_all(path, ec);
        }
    } tmp;
    if (o.genomes) {
        if (!o.quiet) std::fprintf(stderr, " . predict proteins from bacterial genomes\n");
        std::string dir;
        if (!o.predict_only.empty()) {
            std::filesystem::create_directories(o.predict_only);
            dir = o.predict_only;
        } else {
            std::string parent = std::filesystem::path(o.output).parent_path().string();
            if (parent.empty()) parent = ".";
            std::string tmpl = parent + "/tmp_kamino_XXXXXX";
            if (!mkdtemp(tmpl.data())) throw std::runtime_error("create temporary directory in " + parent);
            tmp.path = tmpl;
            dir = tmpl;
        }
        inputs = predict_proteomes(inputs, dir, o.threads);
        if (!o.predict_only.empty()) return;
    }
    if (inputs.empty()) throw std::runtime_error("At least one input source with files must be provided.");
    StageTimer t_all("total");
    Options opt = o;
    if (opt.devices.empty()) opt.devices.push_back(opt.device);
    if (opt.devices.size() != 1 && opt.devices.size() != 2 && opt.devices.size() != 4 && opt.devices.size() != 8 && opt.devices.size() != 16)
        throw std::runtime_error("--devices takes 1, 2, 4, 8 or 16 devices (the table splits evenly over them)");
    GpuSet gpus;   // lives until the outputs are written: --nj reuses the contexts of the passes
    Extraction ex = extract_groups(inputs, k, o.min_freq, o.length_middle, o.recode, opt, gpus);
    if (!o.quiet) std::fprintf(stderr, "# analyse variant groups\n . raw variant groups: %zu\n", ex.n_groups);
    std::vector<Candidate> cands;
    { StageTimer t("sort_and_deduplicate_groups"); cands = sort_and_deduplicate_groups(ex, o.recode); }
    if (!o.quiet) std::fprintf(stderr, " . sorted variant groups: %zu\n", cands.size());
    Alignment al;
    { StageTimer t("filter_groups"); al = filter_groups(ex, cands, o.min_freq, constant, o.mask, o.recode, o.threads); }
    if (!o.quiet) std::fprintf(stderr, " . filtered variant groups: %zu\n", al.partitions.size());
    size_t alen = 0;
    double amiss = 0;
    { StageTimer t("write_outputs (+NJ)"); write_outputs(o.output, al, o.nj, gpus, o.threads, &alen, &amiss); }
    if (!o.quiet) std::fprintf(stderr, "# output files\n . alignment: length=%zu missing=%.1f%%\n", alen, amiss);
    if (o.nj && !o.quiet) std::fprintf(stderr, " . NJ tree\n");
}

static std::vector<int> parse_devices(const std::string& v) {   // "0,1,2,3"
    std::vector<int> out;
    size_t pos = 0;
    while (pos <= v.size()) {
        const size_t comma = std::min(v.find(',', pos), v.size());
        const std::string tok = v.substr(pos, comma - pos);
        if (tok.empty() || tok.find_first_not_of("0123456789") != std::string::npos)
            throw std::runtime_error("invalid value '" + v + "' for '--devices' (comma-separated CUDA ordinals)");
        out.push_back(std::stoi(tok));
        pos = comma + 1;
    }
    return out;
}

int cli(int argc, char** argv) {   // main.rs:5-9 + clap surface of lib.rs:119-175
    Options o;
    if (const char* d = std::getenv("KAMINO_DEVICE")) o.device = std::atoi(d);
    try {
        if (const char* d = std::getenv("KAMINO_DEVICES")) o.devices = parse_devices(d);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    try {
        for (int i = 1; i < argc; i++) {
            std::string a = argv[i], val;
            bool inl = false;
            if (a.rfind("--", 0) == 0) {
                const size_t eq = a.find('=');
                if (eq != std::string::npos) { val = a.substr(eq + 1); a.resize(eq); inl = true; }
            }
            auto need = [&]() -> std::string {
                if (inl) return val;
                if (i + 1 >= argc) throw std::runtime_error("a value is required for '" + a + "'");
                return argv[++i];
            };
            if (a == "-i" || a == "--input-directory") { o.input_dir = need(); o.has_dir = true; }
            else if (a == "-I" || a == "--input-file") { o.input_file = need(); o.has_file = true; }
            else if (a == "--genomes") o.genomes = true;
            else if (a == "-k" || a == "--k") o.k = std::stol(need());
            else if (a == "-f" || a == "--min-freq") o.min_freq = std::strtof(need().c_str(), nullptr);
            else if (a == "-o" || a == "--output") o.output = need();
            else if (a == "-c" || a == "--constant") o.constant = std::stol(need());
            else if (a == "-l" || a == "--length-middle") o.length_middle = std::stoul(need());
            else if (a == "-m" || a == "--mask") o.mask = std::stoul(need());
            else if (a == "-t" || a == "--threads") o.threads = std::stoul(need());
            else if (a == "-r" || a == "--recode") {
                const std::string v = need();
                if (v == "dayhoff6") o.recode = Dayhoff6; else if (v == "sr6") o.recode = SR6; else if (v == "kgb6") o.recode = KGB6;
                else throw std::runtime_error("invalid value '" + v + "' for '--recode'");
            }
            else if (a == "--nj" || a == "--NJ") o.nj = true;           // README spells it --NJ, clap accepts --nj
            else if (a == "--device") o.device = std::stoi(need());
            else if (a == "--devices") o.devices = parse_devices(need());
            else if (a == "--batch-bytes") o.batch_bytes = std::stoull(need());
            else if (a == "--predict-only") o.predict_only = need();
            else if (a == "--self-test") {   // hidden: host-logic unit tests that need no GPU
                const std::string r = prediction_self_test();
                if (!r.empty()) throw std::runtime_error("protein prediction self-test " + r);
                const std::string x = extraction_self_test();
                if (!x.empty()) throw std::runtime_error("extraction self-test " + x);
                const std::string h = host_logic_self_test();
                if (!h.empty()) throw std::runtime_error("host logic self-test " + h);
                const std::string f = fasta_index_self_test();
                if (!f.empty()) throw std::runtime_error("FASTA record table self-test " + f);
                std::printf("self-test ok\n");
                return 0;
            }
            else if (a == "-v" || a == "-V" || a == "--version") { std::printf("kamino 2.0.0 (b200)\n"); return 0; }
            else throw std::runtime_error("unexpected argument '" + a + "'");
        }
        if (!o.has_dir && !o.has_file) throw std::runtime_error("one of --input-directory / --input-file is required");
        run(o);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}

}  // namespace kh
