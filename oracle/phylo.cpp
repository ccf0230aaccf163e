// phylo.cpp — CPU ORACLE (test infrastructure).
// Restates /root/reference/src/phylo.rs.  Compile with -ffp-contract=off: Rust never fuses
// a*b+c, and the NJ criterion / branch lengths are f64 tie-break sensitive.
#include "kamino_oracle.hpp"
#include <atomic>
#include <cmath>
#include <cstdio>
#include <limits>
#include <set>
#include <thread>

namespace kmo {

const double LG_PI[20] = {  // phylo.rs:27-31
    0.079066, 0.055941, 0.041977, 0.053052, 0.012937, 0.040767, 0.071586, 0.057337,
    0.022355, 0.062157, 0.099081, 0.064600, 0.022951, 0.042302, 0.044040, 0.061197,
    0.053287, 0.012777, 0.027843, 0.070200,
};

double f81_constant() {  // phylo.rs:34-43 (sequential f64 sum)
    double s = 0.0;
    for (int i = 0; i < 20; i++) {
        double pi = LG_PI[i];
        s += pi * pi;
    }
    return 1.0 - s;
}

uint8_t aa_index_or_20(uint8_t b) {  // phylo.rs:47-79 (upper-case only; anything else -> 20)
    static const char order[] = "ARNDCQEGHILKMFPSTWYV";
    for (int i = 0; i < 20; i++)
        if (b == uint8_t(order[i])) return uint8_t(i);
    return 20;
}

void pair_counts(const uint8_t* a, const uint8_t* b, size_t len, uint64_t& valid, uint64_t& mism) {  // phylo.rs:84-95
    uint64_t v = 0, m = 0;
    for (size_t i = 0; i < len; i++) {
        if (a[i] < 20 && b[i] < 20) {
            v++;
            if (a[i] != b[i]) m++;
        }
    }
    valid = v;
    mism = m;
}

double lg_f81_from_counts(uint64_t valid_sites, uint64_t mismatches) {  // phylo.rs:96-105
    if (valid_sites == 0) return 0.0;
    double p = double(mismatches) / double(valid_sites);
    double pc = p / f81_constant();
    if (pc >= 1.0) pc = 0.999999999;
    return -std::log(1.0 - pc);
}

std::vector<double> compute_distance_matrix(const std::vector<std::vector<uint8_t>>& seqs,
                                            size_t threads) {  // phylo.rs:114-141
    size_t n = seqs.size();
    size_t dim = 2 * n - 1;
    std::vector<double> dist(dim * dim, 0.0);
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) break;
            for (size_t j = i + 1; j < n; j++) {
                uint64_t v, m;
                pair_counts(seqs[i].data(), seqs[j].data(), std::min(seqs[i].size(), seqs[j].size()), v, m);
                double d = lg_f81_from_counts(v, m);
                dist[i * dim + j] = d;
                dist[j * dim + i] = d;
            }
        }
    };
    size_t nt = std::max<size_t>(1, std::min(threads, n));
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nt; t++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    return dist;
}

struct Edge { size_t child; double len; };
struct Node { bool leaf; std::string name; std::vector<Edge> children; };

static std::string escape_name(const std::string& name) {  // phylo.rs:235-245
    bool needs = false;
    for (char c : name)
        if (c == ' ' || c == ':' || c == '(' || c == ')' || c == ',' || c == ';') needs = true;
    if (!needs) return name;
    std::string e = "'";
    for (char c : name) {
        if (c == '\'') e += "''";
        else e += c;
    }
    return e + "'";
}
static std::string format_len(double len) {  // phylo.rs:248-250 ({:.6})
    char buf[512];
    std::snprintf(buf, sizeof buf, "%.6f", len);
    return buf;
}
static std::string format_subtree(const std::vector<Node>& nodes, size_t id) {  // phylo.rs:253-265
    const Node& node = nodes[id];
    if (node.children.empty()) return escape_name(node.name);
    std::string s = "(";
    for (size_t i = 0; i < node.children.size(); i++) {
        if (i) s += ",";
        s += format_subtree(nodes, node.children[i].child) + ":" + format_len(node.children[i].len);
    }
    return s + ")";
}

std::string nj_newick_from_matrix(const std::vector<std::string>& names, std::vector<double>& dist) {  // phylo.rs:145-272
    size_t n = names.size();
    size_t dim = 2 * n - 1;
    auto idx = [dim](size_t i, size_t j) { return i * dim + j; };
    std::vector<Node> nodes;
    nodes.reserve(dim);
    for (auto& nm : names) nodes.push_back(Node{true, nm, {}});
    std::vector<size_t> active(n);
    for (size_t i = 0; i < n; i++) active[i] = i;
    while (active.size() > 2) {
        size_t m = active.size();
        std::vector<double> r(dim, 0.0);
        for (size_t i : active) {
            double sum = 0.0;
            for (size_t j : active)
                if (i != j) sum += dist[idx(i, j)];
            r[i] = sum;
        }
        size_t bi = active[0], bj = active[1];
        double best_q = std::numeric_limits<double>::infinity();
        for (size_t pi = 0; pi < active.size(); pi++) {
            size_t i = active[pi];
            for (size_t pj = pi + 1; pj < active.size(); pj++) {
                size_t j = active[pj];
                double q = double(m - 2) * dist[idx(i, j)] - r[i] - r[j];
                if (q < best_q) { best_q = q; bi = i; bj = j; }
            }
        }
        size_t i = bi, j = bj;
        double dij = dist[idx(i, j)];
        double denom = 2.0 * double(m - 2);
        double li = 0.5 * dij + (r[i] - r[j]) / denom;
        double lj = dij - li;
        if (li < 0.0) li = 0.0;
        if (lj < 0.0) lj = 0.0;
        size_t u = nodes.size();
        nodes.push_back(Node{false, "", {Edge{i, li}, Edge{j, lj}}});
        for (size_t kk : active) {
            if (kk == i || kk == j) continue;
            double duk = 0.5 * (dist[idx(i, kk)] + dist[idx(j, kk)] - dij);
            dist[idx(u, kk)] = duk;
            dist[idx(kk, u)] = duk;
        }
        std::vector<size_t> na;
        for (size_t x : active)
            if (x != i && x != j) na.push_back(x);
        na.push_back(u);
        active.swap(na);
    }
    size_t a = active[0], b = active[1];
    double len = dist[idx(a, b)] * 0.5;
    if (len < 0.0) len = 0.0;
    size_t root = nodes.size();
    nodes.push_back(Node{false, "", {Edge{a, len}, Edge{b, len}}});
    return format_subtree(nodes, root) + ";";
}

std::string nj_tree_newick(const std::vector<std::string>& names,
                           const std::vector<std::vector<uint8_t>>& seqs, size_t threads) {  // phylo.rs:274-301
    if (names.size() != seqs.size()) throw std::runtime_error("names and sequences length mismatch");
    if (names.size() < 2) throw std::runtime_error("need at least 2 sequences");
    std::set<size_t> lengths;
    for (auto& s : seqs) lengths.insert(s.size());
    if (lengths.size() != 1) throw std::runtime_error("all sequences must have identical lengths");
    std::vector<std::vector<uint8_t>> mapped;
    for (auto& s : seqs) {
        std::vector<uint8_t> m(s.size());
        for (size_t i = 0; i < s.size(); i++) m[i] = aa_index_or_20(s[i]);
        mapped.push_back(std::move(m));
    }
    std::vector<double> dist = compute_distance_matrix(mapped, threads);
    return nj_newick_from_matrix(names, dist);
}

}  // namespace kmo
