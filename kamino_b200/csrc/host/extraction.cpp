// extraction.cpp — stage driver of the B200 pipeline: both passes run on the GPU through the C ABI,
// the host only pairs the boundary hits the device returns and books the resulting blocks.
// Equivalent to /root/reference/src/group_extraction.rs:146-242 (anchor selection / pairing),
// :397-404 (protein names) and :527-660 (stage order, limits, deterministic sid-ordered merge).
#include "../../../include/kamino_b200.h"
#include "kamino_host.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

namespace kh {

namespace {

struct Ctx {   // RAII over kmn_ctx
    kmn_ctx* p = nullptr;
    Ctx(int device, int k, int scheme) {
        if (kmn_create(&p, device, k, scheme) != 0) throw std::runtime_error(std::string("GPU front end: ") + kmn_last_error());
    }
    ~Ctx() { kmn_destroy(p); }
    void ok(int rc) const { if (rc != 0) throw std::runtime_error(std::string("GPU front end: ") + kmn_last_error()); }
};

bool valid_utf8(const std::string& s) {   // rec.id()? / rec.desc() fail on invalid UTF-8 (group_extraction.rs:429-430)
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s.data());
    size_t i = 0, n = s.size();
    while (i < n) {
        const unsigned c = p[i];
        if (c < 0x80) { i++; continue; }
        int need;
        unsigned cp;
        if (c >= 0xC2 && c <= 0xDF) { need = 1; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { need = 2; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { need = 3; cp = c & 0x07; }
        else return false;
        if (i + size_t(need) >= n) return false;
        for (int j = 1; j <= need; j++) {
            const unsigned d = p[i + j];
            if ((d & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (d & 0x3F);
        }
        if (need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
        if (need == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
        i += size_t(need) + 1;
    }
    return true;
}

bool is_ws(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
std::string trim(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && is_ws((unsigned char)s[b])) b++;
    while (e > b && is_ws((unsigned char)s[e - 1])) e--;
    return s.substr(b, e - b);
}

// "id desc" with both parts trimmed, cut at the first " [" (group_extraction.rs:397-404, :429-436)
std::string protein_name_of(const std::string& head) {
    const size_t sp = head.find(' ');
    const std::string id = trim(sp == std::string::npos ? head : head.substr(0, sp));
    const std::string desc = sp == std::string::npos ? std::string() : trim(head.substr(sp + 1));
    std::string name = trim(desc.empty() ? id : id + " " + desc);
    const size_t cut = name.find(" [");
    if (cut != std::string::npos) {
        name.resize(cut);
        while (!name.empty() && is_ws((unsigned char)name.back())) name.pop_back();
    }
    return name;
}

}  // namespace

// group_extraction.rs:666-677 (truncate_protein_name_keeps_record_id_and_removes_organism_suffix) on protein_name_of,
// plus the UTF-8 gate of the header fields (:429-430)
std::string extraction_self_test() {
    if (protein_name_of("WP_012345.1 hypothetical protein [Escherichia coli]") != "WP_012345.1 hypothetical protein")
        return "failed: group_extraction::truncate_protein_name (organism suffix)";
    if (protein_name_of("WP_012345.1 hypothetical protein") != "WP_012345.1 hypothetical protein")
        return "failed: group_extraction::truncate_protein_name (no suffix)";
    if (protein_name_of("id_only") != "id_only" || protein_name_of("id   spaced  desc [x] [y]") != "id spaced  desc")
        return "failed: group_extraction::protein names (id only / inner spaces)";
    if (!valid_utf8("caf\xc3\xa9") || valid_utf8("bad\xff") || valid_utf8("\xc3") || valid_utf8("\xed\xa0\x80"))
        return "failed: group_extraction::header UTF-8 validation";
    return std::string();
}

Extraction extract_groups(const std::vector<SpeciesFile>& inputs, size_t k, float min_freq, size_t length_middle,
                          Scheme scheme, const Options& opt) {
    const size_t n = inputs.size();
    if (n > 0xFFFF)   // group_extraction.rs:538-543
        throw std::runtime_error("too many species (" + std::to_string(n) + ") for compact direct block observations; maximum is 65535");
    const float nf = float(n < 1 ? size_t(1) : n);
    volatile float prod = min_freq * nf;                       // f32 arithmetic, group_extraction.rs:544
    const uint32_t min_needed = uint32_t(std::ceil(float(prod)));
    // the CUDA context and the 512 MiB table come up on a side thread while the host threads load the proteomes
    std::unique_ptr<Ctx> gpu_holder;
    std::string gpu_err;
    std::thread gpu_init([&] {
        StageTimer t("kmn_create (overlaps the load)");
        try { gpu_holder = std::make_unique<Ctx>(opt.device, int(k), int(scheme)); }
        catch (const std::exception& e) { gpu_err = *e.what() ? e.what() : "GPU front end: error"; }
    });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } gpu_joiner{gpu_init};

    // ---- load every proteome (host threads; gunzip, and FASTA framing only with KAMINO_HOST_FASTA=1: by
    // default the file bytes go to the GPU as they are and kmn_count_fasta finds the records)
    const bool device_fasta = std::getenv("KAMINO_HOST_FASTA") == nullptr;
    if (!opt.quiet) std::fprintf(stderr, " . pass 1: estimate k-mer occupancies\n");
    std::vector<Proteome> prot(n);
    std::vector<std::string> errs(n);
    {
        StageTimer t("load proteomes");
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t i; (i = next.fetch_add(1)) < n;) {
                try {
                    if (device_fasta) prot[i].seq = load_fasta_bytes(inputs[i].path);   // records are found on the device
                    else prot[i] = load_proteome(inputs[i].path);
                }
                catch (const std::exception& e) { errs[i] = *e.what() ? e.what() : "error"; }
            }
        };
        std::vector<std::thread> pool;
        for (size_t t = 0; t < std::max<size_t>(1, std::min(opt.threads, n)); t++) pool.emplace_back(work);
        for (auto& t : pool) t.join();
    }
    gpu_init.join();
    if (!gpu_err.empty()) throw std::runtime_error(gpu_err);
    for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);   // first error in species order (:569-571)
    Ctx& gpu = *gpu_holder;

    // ---- pass 1: batches of whole species through kmn_count_batch
    auto t_p1 = std::make_unique<StageTimer>("pass 1 (stage + count + finalize)");
    struct BatchRef { uint32_t id; size_t sp0, sp1; };
    std::vector<BatchRef> batches;
    // the batch buffer is page-locked (kmn_host_alloc): its copies then overlap the kernels of kmn_count_batch
    struct Pinned {
        uint8_t* p = nullptr;
        size_t cap = 0;
        ~Pinned() { kmn_host_free(p); }
        void ensure(size_t n, const Ctx& gpu) {
            if (n <= cap) return;
            kmn_host_free(p);
            p = nullptr;
            cap = 0;
            void* q = nullptr;
            gpu.ok(kmn_host_alloc(&q, n));
            p = static_cast<uint8_t*>(q);
            cap = n;
        }
    } buf;
    std::vector<uint64_t> rec_off, sp_rec_off, hdr_off;
    for (size_t s0 = 0; s0 < n;) {
        size_t s1 = s0, bytes = 0, recs = 0;
        while (s1 < n && (s1 == s0 || bytes + prot[s1].seq.size() <= opt.batch_bytes)) {
            bytes += prot[s1].seq.size();
            recs += prot[s1].n_rec();
            s1++;
        }
        if (bytes + recs >= (uint64_t(1) << 31)) throw std::runtime_error("species " + inputs[s0].name + " is too large for one GPU batch");
        buf.ensure(bytes, gpu);
        rec_off.assign(1, 0);
        sp_rec_off.assign(1, 0);
        size_t at = 0;
        std::vector<uint64_t> sp_at(s1 - s0 + 1);
        for (size_t s = s0; s < s1; s++) {
            sp_at[s - s0] = at;
            for (size_t r = 0; r < prot[s].n_rec(); r++) rec_off.push_back(at + prot[s].rec_end[r]);
            at += prot[s].seq.size();
            sp_rec_off.push_back(rec_off.size() - 1);
        }
        sp_at[s1 - s0] = at;
        auto on_threads = [&](auto&& per_species) {
            std::atomic<size_t> next{s0};
            auto work = [&] { for (size_t s; (s = next.fetch_add(1)) < s1;) per_species(s); };
            std::vector<std::thread> pool;
            for (size_t t = 0; t < std::max<size_t>(1, std::min(opt.threads, s1 - s0)); t++) pool.emplace_back(work);
            for (auto& t : pool) t.join();
        };
        // gather the species into the pinned buffer on the host threads
        on_threads([&](size_t s) {
            if (!prot[s].seq.empty()) std::memcpy(buf.p + sp_at[s - s0], prot[s].seq.data(), prot[s].seq.size());
        });
        uint32_t id = 0;
        if (device_fasta) {
            uint64_t n_rec = 0;
            if (kmn_count_fasta(gpu.p, buf.p, sp_at.data(), uint32_t(s1 - s0), nullptr, &id, &n_rec) != 0) {
                const std::string gpu_msg = kmn_last_error();
                // a format error: the host reader words it like the reference does, first bad file in species order
                for (size_t s = s0; s < s1; s++) parse_proteome(prot[s].seq, inputs[s].path);
                throw std::runtime_error("GPU front end: " + gpu_msg);
            }
            hdr_off.resize(n_rec);
            sp_rec_off.assign(s1 - s0 + 1, 0);
            uint64_t got = 0;
            gpu.ok(kmn_fetch_records(gpu.p, id, hdr_off.data(), n_rec, &got, sp_rec_off.data()));
            std::vector<std::string> idx_err(s1 - s0);
            on_threads([&](size_t s) {   // headers and sequence ranges of the records the device found
                try {
                    index_records(prot[s], hdr_off.data() + sp_rec_off[s - s0], size_t(sp_rec_off[s - s0 + 1] - sp_rec_off[s - s0]),
                                  sp_at[s - s0]);
                } catch (const std::exception& e) { idx_err[s - s0] = e.what(); }
            });
            for (auto& e : idx_err) if (!e.empty()) throw std::runtime_error(e);
        } else {
            gpu.ok(kmn_count_batch(gpu.p, buf.p, rec_off.data(), rec_off.size() - 1, sp_rec_off.data(),
                                   uint32_t(s1 - s0), nullptr, &id));
        }
        batches.push_back({id, s0, s1});
        s0 = s1;
    }
    gpu.ok(kmn_finalize_table(gpu.p));   // snapshot (:576)
    t_p1.reset();
    auto t_p2 = std::make_unique<StageTimer>("pass 2 (scan + pairing)");

    // ---- pass 2: occupancy scan, boundary rule, anchor selection and bubble pairing on the device
    // (kmn_scan_batch + kmn_pair_batch); the host slices the amino-acid blocks per species on its threads
    // and merges them in sid order (:611-636)
    if (!opt.quiet) std::fprintf(stderr, " . pass 2: extract sequences\n");
    Extraction ex;
    ex.k = k;
    ex.min_needed = min_needed;
    ex.n_species = n;
    for (auto& in : inputs) ex.species_names.push_back(in.name);
    struct SpeciesOut {
        std::vector<kmn_pair> pairs;
        std::vector<Observation> obs;   // protein = record index, aa_off local to `arena`
        std::vector<uint8_t> arena;
        std::vector<std::string> names;  // protein names of the species' records
        std::string err;
    };
    std::vector<SpeciesOut> outs(n);
    auto t_dev = std::make_unique<StageTimer>("  pass 2: device scan + pairing + fetch");
    for (const BatchRef& b : batches) {
        uint64_t ts = 0, te = 0, np = 0;
        gpu.ok(kmn_scan_batch(gpu.p, b.id, min_needed, &ts, &te));
        gpu.ok(kmn_pair_batch(gpu.p, b.id, uint32_t(std::min<size_t>(length_middle, 0xFFFFFFFFu)), &np));
        for (size_t sid = b.sp0; sid < b.sp1; sid++) {   // calls on one context are serialised
            uint64_t cnt = 0;
            gpu.ok(kmn_fetch_pairs(gpu.p, b.id, uint32_t(sid - b.sp0), nullptr, 0, &cnt));
            outs[sid].pairs.resize(cnt);
            if (cnt) gpu.ok(kmn_fetch_pairs(gpu.p, b.id, uint32_t(sid - b.sp0), outs[sid].pairs.data(), cnt, &cnt));
        }
    }
    t_dev.reset();
    {
        StageTimer t_par("  pass 2: host block slicing (threads)");
        std::atomic<size_t> next{0};
        auto work = [&] {
            std::vector<uint8_t> aa;     // whitespace-stripped, upper-cased residues of the current record
            for (size_t sid; (sid = next.fetch_add(1)) < n;) {
                const Proteome& pr = prot[sid];
                SpeciesOut& o = outs[sid];
                for (const std::string& h : pr.heads) {   // id()/desc() are validated for every record (:429-430)
                    const size_t sp = h.find(' ');
                    if (!valid_utf8(sp == std::string::npos ? h : h.substr(0, sp)) ||
                        (sp != std::string::npos && !valid_utf8(h.substr(sp + 1)))) {
                        o.err = "invalid UTF-8 in FASTA header of species " + inputs[sid].name;
                        break;
                    }
                }
                if (!o.err.empty()) continue;
                o.names.reserve(pr.heads.size());
                for (const std::string& h : pr.heads) o.names.push_back(protein_name_of(h));
                uint32_t aa_rec = 0xFFFFFFFFu;
                o.obs.reserve(o.pairs.size());
                for (const kmn_pair& p : o.pairs) {
                    if (p.len > 0xFFFF) {   // :92-94
                        o.err = "amino-acid block length " + std::to_string(p.len) + " exceeds u16::MAX";
                        break;
                    }
                    if (aa_rec != p.rec) {   // strip + upper-case the record once (group_extraction.rs:455-482)
                        aa.clear();
                        for (uint64_t i = pr.rec_begin[p.rec]; i < pr.rec_end[p.rec]; i++) {
                            const uint8_t c = pr.seq[i];
                            if (c == 0x20 || c == 0x09 || c == 0x0A || c == 0x0C || c == 0x0D) continue;
                            aa.push_back((c >= 'a' && c <= 'z') ? uint8_t(c - 32) : c);
                        }
                        aa_rec = p.rec;
                    }
                    const size_t lo = size_t(p.seg_start) + p.left_start;
                    Observation ob;
                    ob.left = p.left;
                    ob.right = p.right;
                    ob.sid = uint32_t(sid);
                    ob.protein = p.rec;
                    ob.aa_off = o.arena.size();
                    ob.len = p.len;
                    ob.order = 0;
                    o.arena.insert(o.arena.end(), aa.begin() + lo, aa.begin() + lo + p.len);
                    o.obs.push_back(ob);
                }
                std::vector<kmn_pair>().swap(o.pairs);
            }
        };
        std::vector<std::thread> pool;
        for (size_t t = 0; t < std::max<size_t>(1, std::min(opt.threads, n)); t++) pool.emplace_back(work);
        for (auto& t : pool) t.join();
    }
    // ordered merge (:611-636): offsets by prefix sums in species order, copies on the host threads
    for (size_t sid = 0; sid < n; sid++)
        if (!outs[sid].err.empty()) throw std::runtime_error(outs[sid].err);   // first error in species order
    {
        StageTimer t_merge("  pass 2: ordered merge");
        std::vector<uint64_t> obs0(n + 1, 0), arena0(n + 1, 0), prot0(n + 1, 0);
        for (size_t sid = 0; sid < n; sid++) {
            obs0[sid + 1] = obs0[sid] + outs[sid].obs.size();
            arena0[sid + 1] = arena0[sid] + outs[sid].arena.size();
            prot0[sid + 1] = prot0[sid] + prot[sid].n_rec();
            if (prot0[sid + 1] > 0xFFFFFFFFull)   // :342-344
                throw std::runtime_error("too many protein records across all species to index with u32");
        }
        ex.obs.resize(obs0[n]);
        ex.arena.resize(arena0[n]);
        ex.protein_names.resize(prot0[n]);
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t sid; (sid = next.fetch_add(1)) < n;) {
                SpeciesOut& o = outs[sid];
                if (!o.arena.empty()) std::memcpy(ex.arena.data() + arena0[sid], o.arena.data(), o.arena.size());
                for (size_t i = 0; i < o.obs.size(); i++) {
                    Observation ob = o.obs[i];
                    ob.protein = uint32_t(prot0[sid] + ob.protein);
                    ob.aa_off += arena0[sid];
                    ob.order = obs0[sid] + i;   // emission order == the reference's Vec order
                    ex.obs[obs0[sid] + i] = ob;
                }
                for (size_t r = 0; r < o.names.size(); r++) ex.protein_names[prot0[sid] + r] = std::move(o.names[r]);
                SpeciesOut().obs.swap(o.obs);
                std::vector<uint8_t>().swap(o.arena);
                std::vector<std::string>().swap(o.names);
            }
        };
        std::vector<std::thread> pool;
        for (size_t t = 0; t < std::max<size_t>(1, std::min(opt.threads, n)); t++) pool.emplace_back(work);
        for (auto& t : pool) t.join();
    }
    t_p2.reset();
    StageTimer t_sort("group observations");
    // group by anchor pair on the device (stable radix sort of the observation indices by (left, right): inside
    // a group the emission order == the reference's Vec order survives) + per-group length selection
    const size_t n_obs = ex.obs.size();
    if (n_obs) {
        std::vector<uint64_t> left(n_obs), right(n_obs);
        std::vector<uint32_t> sid(n_obs), len(n_obs), perm(n_obs);
        for (size_t i = 0; i < n_obs; i++) {
            left[i] = ex.obs[i].left;
            right[i] = ex.obs[i].right;
            sid[i] = ex.obs[i].sid;
            len[i] = ex.obs[i].len;
        }
        ex.groups.resize(n_obs);
        uint64_t ng = 0;
        gpu.ok(kmn_group_observations(gpu.p, n_obs, left.data(), right.data(), sid.data(), len.data(), min_needed,
                                      perm.data(), ex.groups.data(), ex.groups.size(), &ng));
        ex.groups.resize(ng);
        ex.n_groups = ng;
        std::vector<Observation> sorted(n_obs);
        std::atomic<size_t> next{0};
        auto work = [&] {
            const size_t step = size_t(1) << 16;
            for (size_t i0; (i0 = next.fetch_add(step)) < n_obs;)
                for (size_t i = i0; i < std::min(n_obs, i0 + step); i++) sorted[i] = ex.obs[perm[i]];
        };
        std::vector<std::thread> pool;
        for (size_t t = 0; t < std::max<size_t>(1, opt.threads); t++) pool.emplace_back(work);
        for (auto& t : pool) t.join();
        ex.obs.swap(sorted);
    }
    return ex;
}

}  // namespace kh
