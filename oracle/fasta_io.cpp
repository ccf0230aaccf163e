// fasta_io.cpp — CPU ORACLE (test infrastructure): input discovery + FASTA reading.
// Restates /root/reference/src/io.rs and the seq_io 0.3.x fasta::Reader behaviour the
// reference relies on (seq_io = "0.3.4" in Cargo.toml:16; crate source is NOT in
// /root/reference — behaviour restated from its documented contract, pinned only by the
// single-line fixtures of tests/data/test_3diff; multi-line / CRLF cases: parity unpinned).
#include "kamino_oracle.hpp"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <set>
#include <sys/stat.h>
#include <zlib.h>

namespace kmo {

// ----------------------------------------------------- Rust std::path helpers
static std::string strip_trailing_slashes(const std::string& p) {
    size_t e = p.size();
    while (e > 1 && p[e - 1] == '/') e--;
    return p.substr(0, e);
}
std::string rs_file_name(const std::string& p0) {
    std::string p = strip_trailing_slashes(p0);
    size_t s = p.rfind('/');
    std::string n = (s == std::string::npos) ? p : p.substr(s + 1);
    if (n == ".." || n == "/" ) return "";
    return n;
}
std::string rs_parent(const std::string& p0, bool* has_parent) {
    std::string p = strip_trailing_slashes(p0);
    if (p.empty() || p == "/") { if (has_parent) *has_parent = false; return ""; }
    size_t s = p.rfind('/');
    if (has_parent) *has_parent = true;
    if (s == std::string::npos) return "";
    if (s == 0) return "/";
    return strip_trailing_slashes(p.substr(0, s));
}
// Path::extension / file_stem: split at the LAST '.', but a leading '.' is not a separator.
std::string rs_extension(const std::string& n) {
    if (n.empty() || n == "..") return "";
    size_t d = n.rfind('.');
    if (d == std::string::npos || d == 0) return "";
    return n.substr(d + 1);
}
std::string rs_file_stem(const std::string& n) {
    if (n.empty() || n == "..") return n;
    size_t d = n.rfind('.');
    if (d == std::string::npos || d == 0) return n;
    return n.substr(0, d);
}

static bool eq_ignore_ascii_case(const std::string& a, const char* b) {
    size_t n = std::strlen(b);
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; i++) {
        char x = a[i], y = b[i];
        if (x >= 'A' && x <= 'Z') x = char(x + 32);
        if (y >= 'A' && y <= 'Z') y = char(y + 32);
        if (x != y) return false;
    }
    return true;
}
static bool matches_fasta_ext(const std::string& ext) {  // io.rs:34-37
    static const char* set[] = {"fa", "fas", "fasta", "faa", "fna"};
    for (const char* e : set)
        if (eq_ignore_ascii_case(ext, e)) return true;
    return false;
}
static bool is_supported_ext(const std::string& file_name) {  // io.rs:19-32
    std::string ext = rs_extension(file_name);
    if (eq_ignore_ascii_case(ext, "gz")) {
        std::string stem = rs_file_stem(file_name);
        return matches_fasta_ext(rs_extension(stem));
    }
    return matches_fasta_ext(ext);
}

std::vector<std::string> collect_fasta_files(const std::string& dir) {  // io.rs:16-49
    DIR* d = opendir(dir.c_str());
    if (!d) throw std::runtime_error("read_dir " + dir + ": cannot open directory");
    std::vector<std::string> names;
    while (struct dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        std::string full = dir + (dir.empty() || dir.back() == '/' ? "" : "/") + n;
        struct stat st;
        // DirEntry::file_type does not follow symlinks (lstat); only regular files pass.
        if (lstat(full.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
        if (!is_supported_ext(n)) continue;
        names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());  // io.rs:46 (byte-wise file_name order)
    std::vector<std::string> out;
    for (auto& n : names) out.push_back(dir + (dir.empty() || dir.back() == '/' ? "" : "/") + n);
    return out;
}

static std::string ascii_lower(std::string s) {
    for (auto& c : s)
        if (c >= 'A' && c <= 'Z') c = char(c + 32);
    return s;
}
static bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

std::string species_name_from_path(const std::string& p) {  // io.rs:164-183
    std::string stem = rs_file_name(p);
    if (stem.empty()) stem = "unnamed";
    if (ends_with(stem, ".gz")) stem.resize(stem.size() - 3);
    for (const char* ext : {".fa", ".fasta", ".fas", ".faa", ".fna"}) {
        if (ends_with(ascii_lower(stem), ext)) {
            stem.resize(stem.size() - std::strlen(ext));
            break;
        }
    }
    return stem;
}

static void sort_inputs(std::vector<SpeciesInput>& v) {  // io.rs:77,160
    std::sort(v.begin(), v.end(), [](const SpeciesInput& a, const SpeciesInput& b) {
        if (a.name != b.name) return a.name < b.name;
        return a.path < b.path;
    });
}

std::vector<SpeciesInput> collect_species_inputs_from_dir(const std::string& dir) {  // io.rs:61-79
    std::vector<SpeciesInput> inputs;
    std::set<std::string> seen;
    for (auto& path : collect_fasta_files(dir)) {
        std::string name = species_name_from_path(path);
        if (!seen.insert(name).second)
            throw std::runtime_error("isolate name \"" + name +
                                     "\" appears more than once in input directory.");
        inputs.push_back({name, path});
    }
    sort_inputs(inputs);
    return inputs;
}

// str::trim / trim_end (Unicode White_Space, kamino_oracle.hpp)
static std::string trim_end(const std::string& s) { return rs_trim_end(s); }
static std::string trim(const std::string& s) { return rs_trim(s); }

std::vector<SpeciesInput> collect_species_inputs_from_table(const std::string& table) {  // io.rs:82-162
    std::ifstream f(table, std::ios::binary);
    if (!f) throw std::runtime_error("open \"" + table + "\"");
    bool hp = false;
    std::string base = rs_parent(table, &hp);
    if (!hp) base = ".";
    std::vector<SpeciesInput> inputs;
    std::set<std::string> seen;
    std::string line;
    size_t line_idx = 0;
    while (std::getline(f, line)) {
        line_idx++;
        if (!line.empty() && line.back() == '\r') line.pop_back();  // BufRead::lines strips \r\n
        std::string t = trim_end(line);
        if (t.empty()) continue;
        std::vector<std::string> parts;
        size_t s = 0;
        for (;;) {
            size_t tab = t.find('\t', s);
            if (tab == std::string::npos) { parts.push_back(t.substr(s)); break; }
            parts.push_back(t.substr(s, tab - s));
            s = tab + 1;
        }
        std::string where = "Line " + std::to_string(line_idx) + " in \"" + table + "\"";
        if (parts.size() < 2) throw std::runtime_error(where + " is missing a proteome path.");
        if (parts.size() > 2)
            throw std::runtime_error(where + " has more than two tab-delimited columns.");
        std::string name = trim(parts[0]), path_str = trim(parts[1]);
        if (name.empty() || path_str.empty())
            throw std::runtime_error(where + " must contain non-empty isolate name and path.");
        if (!seen.insert(name).second)
            throw std::runtime_error("isolate name \"" + name + "\" appears more than once in \"" +
                                     table + "\".");
        std::string path = path_str;
        if (path[0] != '/') path = base.empty() ? path : (base == "/" ? "/" + path : base + "/" + path);
        struct stat st;
        if (stat(path.c_str(), &st) != 0) throw std::runtime_error("read metadata for \"" + path + "\"");
        if (!S_ISREG(st.st_mode))
            throw std::runtime_error("Proteome path \"" + path + "\" (line " + std::to_string(line_idx) +
                                     " in \"" + table + "\") is not a file.");
        inputs.push_back({name, path});
    }
    sort_inputs(inputs);
    return inputs;
}

// ------------------------------------------------------------- open_fasta
std::string read_all_maybe_gz(const std::string& path) {  // io.rs:185-200
    std::string out;
    if (ends_with(path, ".gz")) {  // case-sensitive suffix test, io.rs:191
        gzFile g = gzopen(path.c_str(), "rb");
        if (!g) throw std::runtime_error("open \"" + path + "\"");
        gzbuffer(g, 512 * 1024);
        if (gzdirect(g)) {  // flate2::MultiGzDecoder rejects non-gzip input
            gzclose(g);
            throw std::runtime_error("invalid gzip header in \"" + path + "\"");
        }
        std::vector<char> buf(1 << 20);
        for (;;) {
            int n = gzread(g, buf.data(), (unsigned)buf.size());
            if (n < 0) {
                int err = 0;
                std::string msg = gzerror(g, &err);
                gzclose(g);
                throw std::runtime_error("gzip read error in \"" + path + "\": " + msg);
            }
            if (n == 0) break;
            out.append(buf.data(), size_t(n));
        }
        gzclose(g);
    } else {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("open \"" + path + "\"");
        std::vector<char> buf(1 << 20);
        size_t n;
        while ((n = std::fread(buf.data(), 1, buf.size(), f)) > 0) out.append(buf.data(), n);
        std::fclose(f);
    }
    return out;
}

// seq_io::fasta::Reader semantics: blank lines before the first record are skipped, any
// other first byte than '>' is an error; a record runs to the next LINE starting with '>'.
std::vector<FastaRecord> parse_fasta_bytes(const std::string& b, const std::string& what) {
    std::vector<FastaRecord> recs;
    size_t n = b.size(), pos = 0;
    // skip leading empty lines (optionally CR-terminated)
    while (pos < n) {
        if (b[pos] == '\n') { pos++; continue; }
        if (b[pos] == '\r' && pos + 1 < n && b[pos + 1] == '\n') { pos += 2; continue; }
        break;
    }
    if (pos >= n) return recs;
    if (b[pos] != '>')
        throw std::runtime_error("FASTA parse error in " + what + ": expected '>' at record start");
    while (pos < n) {
        // pos is at '>'
        size_t eol = b.find('\n', pos);
        size_t head_end = (eol == std::string::npos) ? n : eol;
        FastaRecord r;
        size_t he = head_end;
        if (he > pos + 1 && b[he - 1] == '\r') he--;
        r.head = b.substr(pos + 1, he - (pos + 1));
        size_t seq_start = (eol == std::string::npos) ? n : eol + 1;
        // find next line beginning with '>'
        size_t q = seq_start, next = n;
        if (q < n && b[q] == '>') next = q;
        else {
            for (;;) {
                size_t nl = b.find('\n', q);
                if (nl == std::string::npos || nl + 1 >= n) { next = n; break; }
                if (b[nl + 1] == '>') { next = nl + 1; break; }
                q = nl + 1;
            }
        }
        r.seq = b.substr(seq_start, next - seq_start);
        recs.push_back(std::move(r));
        pos = next;
    }
    return recs;
}

std::vector<FastaRecord> read_fasta(const std::string& path) {
    return parse_fasta_bytes(read_all_maybe_gz(path), "\"" + path + "\"");
}

}  // namespace kmo
