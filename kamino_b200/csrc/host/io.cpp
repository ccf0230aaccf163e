// io.cpp — input discovery and FASTA framing for the B200 pipeline (host side).
// Equivalent to /root/reference/src/io.rs and the seq_io 0.3 fasta::Reader contract it relies on.
#include "kamino_host.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <set>
#include <zlib.h>

namespace fs = std::filesystem;

namespace kh {

static std::string lower(std::string s) {
    for (char& c : s) if (c >= 'A' && c <= 'Z') c = char(c - 'A' + 'a');
    return s;
}
static bool has_suffix(const std::string& s, const char* suf) {
    const size_t n = std::strlen(suf);
    return s.size() >= n && std::memcmp(s.data() + s.size() - n, suf, n) == 0;
}
static bool fasta_ext(std::string ext) {          // io.rs:34-37 (extension given with its leading dot)
    ext = lower(ext);
    for (const char* e : {".fa", ".fas", ".fasta", ".faa", ".fna"}) if (ext == e) return true;
    return false;
}

std::string species_name_of(const std::string& path) {   // io.rs:164-183
    std::string stem = fs::path(path).filename().string();
    if (stem.empty()) stem = "unnamed";
    if (has_suffix(stem, ".gz")) stem.resize(stem.size() - 3);
    const std::string low = lower(stem);
    for (const char* ext : {".fa", ".fasta", ".fas", ".faa", ".fna"})
        if (has_suffix(low, ext)) { stem.resize(stem.size() - std::strlen(ext)); break; }
    return stem;
}

static void order_inputs(std::vector<SpeciesFile>& v) {   // io.rs:77,160
    std::sort(v.begin(), v.end(), [](const SpeciesFile& a, const SpeciesFile& b) {
        return a.name != b.name ? a.name < b.name : a.path < b.path;
    });
}

std::vector<SpeciesFile> inputs_from_directory(const std::string& dir) {   // io.rs:16-79
    std::vector<std::string> names;
    std::error_code ec;
    fs::directory_iterator it(dir, ec);
    if (ec) throw std::runtime_error("read_dir " + dir + ": " + ec.message());
    for (const auto& e : it) {
        if (!fs::is_regular_file(e.symlink_status())) continue;       // DirEntry::file_type: no symlink following
        const fs::path p = e.path().filename();
        std::string ext = p.extension().string();
        if (lower(ext) == ".gz") ext = p.stem().extension().string();  // inner extension of x.faa.gz
        if (!fasta_ext(ext)) continue;
        names.push_back(p.string());
    }
    std::sort(names.begin(), names.end());                             // io.rs:46
    std::vector<SpeciesFile> out;
    std::set<std::string> seen;
    const std::string base = (!dir.empty() && dir.back() == '/') ? dir : dir + "/";
    for (const auto& n : names) {
        SpeciesFile sf{species_name_of(n), base + n};
        if (!seen.insert(sf.name).second)
            throw std::runtime_error("isolate name \"" + sf.name + "\" appears more than once in input directory.");
        out.push_back(std::move(sf));
    }
    order_inputs(out);
    return out;
}

static std::string trimmed(const std::string& s) { return uws::trim(s); }   // str::trim (Unicode White_Space), io.rs:122-123

std::vector<SpeciesFile> inputs_from_table(const std::string& tsv) {   // io.rs:82-162
    std::ifstream f(tsv, std::ios::binary);
    if (!f) throw std::runtime_error("open \"" + tsv + "\"");
    fs::path base = fs::path(tsv).parent_path();
    std::vector<SpeciesFile> out;
    std::set<std::string> seen;
    std::string line;
    for (size_t no = 1; std::getline(f, line); no++) {
        line = uws::trim_end(line);   // strip CR + trim_end
        if (line.empty()) continue;
        std::vector<std::string> cols;
        size_t s = 0;
        for (;;) {
            const size_t t = line.find('\t', s);
            cols.push_back(line.substr(s, t == std::string::npos ? std::string::npos : t - s));
            if (t == std::string::npos) break;
            s = t + 1;
        }
        const std::string where = "Line " + std::to_string(no) + " in \"" + tsv + "\"";
        if (cols.size() < 2) throw std::runtime_error(where + " is missing a proteome path.");
        if (cols.size() > 2) throw std::runtime_error(where + " has more than two tab-delimited columns.");
        const std::string name = trimmed(cols[0]), rel = trimmed(cols[1]);
        if (name.empty() || rel.empty()) throw std::runtime_error(where + " must contain non-empty isolate name and path.");
        if (!seen.insert(name).second)
            throw std::runtime_error("isolate name \"" + name + "\" appears more than once in \"" + tsv + "\".");
        fs::path p(rel);
        if (p.is_relative()) p = base / p;
        std::error_code ec;
        const auto st = fs::status(p, ec);
        if (ec) throw std::runtime_error("read metadata for \"" + p.string() + "\"");
        if (!fs::is_regular_file(st))
            throw std::runtime_error("Proteome path \"" + p.string() + "\" (line " + std::to_string(no) + ") is not a file.");
        out.push_back({name, p.string()});
    }
    order_inputs(out);
    return out;
}

static std::vector<uint8_t> slurp(const std::string& path) {   // io.rs:185-200
    std::vector<uint8_t> data;
    if (has_suffix(path, ".gz")) {                              // case-sensitive suffix test (io.rs:191)
        gzFile g = gzopen(path.c_str(), "rb");
        if (!g) throw std::runtime_error("open \"" + path + "\"");
        gzbuffer(g, 1 << 20);
        if (gzdirect(g)) { gzclose(g); throw std::runtime_error("invalid gzip header in \"" + path + "\""); }
        size_t cap = 1 << 22;
        data.resize(cap);
        size_t n = 0;
        for (;;) {
            if (n == cap) { cap *= 2; data.resize(cap); }
            const int got = gzread(g, data.data() + n, unsigned(std::min<size_t>(cap - n, 1u << 30)));
            if (got < 0) { gzclose(g); throw std::runtime_error("gzip read error in \"" + path + "\""); }
            if (got == 0) break;
            n += size_t(got);
        }
        gzclose(g);
        data.resize(n);
    } else {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("open \"" + path + "\"");
        std::fseek(f, 0, SEEK_END);
        const long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        data.resize(sz > 0 ? size_t(sz) : 0);
        if (sz > 0 && std::fread(data.data(), 1, size_t(sz), f) != size_t(sz)) {
            std::fclose(f);
            throw std::runtime_error("read \"" + path + "\"");
        }
        std::fclose(f);
    }
    return data;
}

// Record = a line starting with '>' up to the next such line; seq() keeps its line terminators
// (the device skips them).  Blank lines before the first record are ignored; any other leading byte
// is a format error (seq_io fasta::Reader).
Proteome parse_proteome(const std::vector<uint8_t>& b, const std::string& path) {
    Proteome p;
    const size_t n = b.size();
    size_t pos = 0;
    while (pos < n && (b[pos] == '\n' || (b[pos] == '\r' && pos + 1 < n && b[pos + 1] == '\n'))) pos += b[pos] == '\n' ? 1 : 2;
    if (pos >= n) return p;
    if (b[pos] != '>') throw std::runtime_error("FASTA parse error in \"" + path + "\": expected '>' at record start");
    p.seq.reserve(n);
    while (pos < n) {
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(b.data() + pos, '\n', n - pos));
        size_t head_end = nl ? size_t(nl - b.data()) : n;
        size_t he = head_end;
        if (he > pos + 1 && b[he - 1] == '\r') he--;
        p.heads.emplace_back(reinterpret_cast<const char*>(b.data() + pos + 1), he - pos - 1);
        size_t q = nl ? head_end + 1 : n;
        const size_t seq_begin = q;
        size_t next = n;
        if (q < n && b[q] == '>') next = q;
        else {
            while (q < n) {
                const uint8_t* l = static_cast<const uint8_t*>(std::memchr(b.data() + q, '\n', n - q));
                if (!l || size_t(l - b.data()) + 1 >= n) break;
                q = size_t(l - b.data()) + 1;
                if (b[q] == '>') { next = q; break; }
            }
        }
        p.rec_begin.push_back(p.seq.size());
        p.seq.insert(p.seq.end(), b.begin() + seq_begin, b.begin() + next);
        p.rec_end.push_back(p.seq.size());
        pos = next;
    }
    return p;
}

Proteome load_proteome(const std::string& path) { return parse_proteome(slurp(path), path); }

std::vector<uint8_t> load_fasta_bytes(const std::string& path) {
    std::vector<uint8_t> b = slurp(path);
    if (!b.empty() && b.back() != '\n') b.push_back('\n');
    return b;
}

void index_records(Proteome& p, const uint64_t* hdr, size_t n_rec, uint64_t base) {
    const std::vector<uint8_t>& b = p.seq;
    const size_t n = b.size();
    p.heads.clear();
    p.rec_begin.assign(n_rec, 0);
    p.rec_end.assign(n_rec, 0);
    p.heads.reserve(n_rec);
    for (size_t r = 0; r < n_rec; r++) {
        const size_t pos = size_t(hdr[r] - base);
        if (pos >= n || b[pos] != '>') throw std::runtime_error("record table of the device does not match the file bytes");
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(b.data() + pos, '\n', n - pos));
        const size_t head_end = nl ? size_t(nl - b.data()) : n;
        size_t he = head_end;
        if (he > pos + 1 && b[he - 1] == '\r') he--;
        p.heads.emplace_back(reinterpret_cast<const char*>(b.data() + pos + 1), he - pos - 1);
        p.rec_begin[r] = nl ? head_end + 1 : n;
        p.rec_end[r] = r + 1 < n_rec ? size_t(hdr[r + 1] - base) : n;
    }
}

// Host-only check of the device-framed path's bookkeeping: for awkward FASTA texts, index_records() fed with
// the header offsets a line scan finds (what kmn_fetch_records returns) must describe the same records —
// headers and whitespace-stripped sequences — as the seq_io-style parser.
std::string fasta_index_self_test() {
    const std::vector<std::string> texts = {
        ">a desc\nACDE\nFGH\n>b\nKLMN\n",
        "\n\r\n>a\r\nAC\r\nDE\r\n>b x y\r\n\r\nFG\r\n",
        ">a\n>b\n\n>c d\nAC>DE\n \t\n>e",
        ">only header",
        "",
        "\n\n",
        ">x>y>z\nMKV\n>w\nLLL",
    };
    auto strip = [](const std::vector<uint8_t>& v, uint64_t b, uint64_t e) {
        std::string out;
        for (uint64_t i = b; i < e; i++) {
            const uint8_t c = v[i];
            if (c == 0x20 || c == 0x09 || c == 0x0A || c == 0x0C || c == 0x0D) continue;
            out.push_back(char(c));
        }
        return out;
    };
    for (size_t t = 0; t < texts.size(); t++) {
        const std::vector<uint8_t> raw(texts[t].begin(), texts[t].end());
        const Proteome want = parse_proteome(raw, "self-test");
        Proteome got;
        got.seq = raw;
        if (!got.seq.empty() && got.seq.back() != '\n') got.seq.push_back('\n');
        std::vector<uint64_t> hdr;
        const uint64_t base = 1000 + t;   // offsets are relative to the batch buffer
        for (size_t i = 0; i < got.seq.size(); i++)
            if (got.seq[i] == '>' && (i == 0 || got.seq[i - 1] == '\n')) hdr.push_back(base + i);
        index_records(got, hdr.data(), hdr.size(), base);
        if (got.n_rec() != want.n_rec() || got.heads != want.heads) return "failed: record table of text " + std::to_string(t);
        for (size_t r = 0; r < want.n_rec(); r++)
            if (strip(got.seq, got.rec_begin[r], got.rec_end[r]) != strip(want.seq, want.rec_begin[r], want.rec_end[r]))
                return "failed: sequence " + std::to_string(r) + " of text " + std::to_string(t);
    }
    return std::string();
}

}  // namespace kh
