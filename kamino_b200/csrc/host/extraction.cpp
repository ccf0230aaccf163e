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

void gpu_ok(int rc) {
    if (rc != 0) throw std::runtime_error(std::string("GPU front end: ") + kmn_last_error());
}

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

// "id desc" with both parts trimmed (Unicode White_Space, like str::trim), cut at the first " ["
// (group_extraction.rs:397-404, :429-436)
std::string protein_name_of(const std::string& head) {
    const size_t sp = head.find(' ');   // seq_io splits id / desc at the first SPACE byte
    const std::string id = uws::trim(sp == std::string::npos ? head : head.substr(0, sp));
    const std::string desc = sp == std::string::npos ? std::string() : uws::trim(head.substr(sp + 1));
    const std::string name = uws::trim(desc.empty() ? id : id + " " + desc);
    const size_t cut = name.find(" [");
    return cut == std::string::npos ? name : uws::trim_end(name.substr(0, cut));
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
    // str::trim follows Unicode White_Space: NEL, NBSP, EN QUAD, LINE SEPARATOR, IDEOGRAPHIC SPACE around the fields
    if (protein_name_of("id1 \xc2\xa0\xe2\x80\x80kinase\xe3\x80\x80 [Homo\xc2\x85]") != "id1 kinase" ||
        protein_name_of("\xc2\x85id2\xe2\x80\xa8 desc\xe2\x81\x9f") != "id2 desc" ||
        protein_name_of("id3 caf\xc3\xa9\xc2\xa0") != "id3 caf\xc3\xa9")
        return "failed: group_extraction::protein names (Unicode white space)";
    return std::string();
}

GpuSet::~GpuSet() {
    for (kmn_ctx* c : ctx) kmn_destroy(c);
}

// One context per device, created concurrently (each creation allocates and zeroes a 512 MiB table).
void GpuSet::create(const std::vector<int>& devs, int k, int scheme) {
    if (!ctx.empty()) return;
    if (devs.empty()) throw std::runtime_error("no GPU device given");
    for (size_t a = 0; a < devs.size(); a++)
        for (size_t b = 0; b < a; b++)
            if (devs[a] == devs[b]) throw std::runtime_error("--devices lists device " + std::to_string(devs[a]) + " twice");
    devices = devs;
    ctx.assign(devs.size(), nullptr);
    std::vector<std::string> errs(devs.size());
    std::vector<std::thread> pool;
    for (size_t r = 0; r < devs.size(); r++)
        pool.emplace_back([&, r] {
            if (kmn_create(&ctx[r], devs[r], k, scheme) != 0) errs[r] = std::string("GPU front end: ") + kmn_last_error();
        });
    for (auto& t : pool) t.join();
    for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
    if (devs.size() > 1) gpu_ok(kmn_exchange_setup_local(ctx.data(), int(devs.size())));
}

Extraction extract_groups(const std::vector<SpeciesFile>& inputs, size_t k, float min_freq, size_t length_middle,
                          Scheme scheme, const Options& opt, GpuSet& gpus) {
    const size_t n = inputs.size();
    if (n > 0xFFFF)   // group_extraction.rs:538-543
        throw std::runtime_error("too many species (" + std::to_string(n) + ") for compact direct block observations; maximum is 65535");
    const float nf = float(n < 1 ? size_t(1) : n);
    volatile float prod = min_freq * nf;                       // f32 arithmetic, group_extraction.rs:544
    const uint32_t min_needed = uint32_t(std::ceil(float(prod)));
    // the CUDA contexts and their 512 MiB tables come up on a side thread while the host threads load the proteomes
    std::string gpu_err;
    std::thread gpu_init([&] {
        StageTimer t("kmn_create (overlaps the load)");
        try { gpus.create(opt.devices, int(k), int(scheme)); }
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
    const size_t G = gpus.size();

    // ---- pass 1: species shard round-robin over the devices (sid mod G: a species never spans devices, so its
    // Bloom filter stays device-local; the reference's par_iter over species, :557-567); every device counts its
    // species in batches of whole species through kmn_count_fasta / kmn_count_batch on its own host thread, then
    // ONE table exchange (the sketch is linear) gives every device the table of the whole job
    auto t_p1 = std::make_unique<StageTimer>("pass 1 (stage + count + exchange)");
    struct BatchRef { uint32_t id; std::vector<size_t> sids; };
    std::vector<std::vector<BatchRef>> rank_batches(G);
    struct RankError { size_t sid = SIZE_MAX; std::string msg; };
    std::vector<RankError> rank_err(G);
    const size_t helpers = std::max<size_t>(1, opt.threads / G);   // host threads of one device's gather / record indexing
    auto on_rank_threads = [&](auto&& per_rank) {
        std::vector<std::thread> pool;
        for (size_t r = 0; r < G; r++)
            pool.emplace_back([&, r] {
                try { per_rank(r); }
                catch (const std::exception& e) {
                    if (rank_err[r].msg.empty()) rank_err[r] = {rank_batches[r].empty() ? r : rank_batches[r].back().sids.front(), e.what()};
                }
            });
        for (auto& t : pool) t.join();
        const RankError* first = nullptr;
        for (auto& e : rank_err)
            if (!e.msg.empty() && (!first || e.sid < first->sid)) first = &e;
        if (first) throw std::runtime_error(first->msg);   // first error in species order (:569-571)
    };
    on_rank_threads([&](size_t r) {
        kmn_ctx* gpu = gpus.ctx[r];
        std::vector<size_t> mine;
        for (size_t sid = r; sid < n; sid += G) mine.push_back(sid);
        // the batch buffer is page-locked (kmn_host_alloc): its copies then overlap the kernels of kmn_count_batch
        struct Pinned {
            uint8_t* p = nullptr;
            size_t cap = 0;
            ~Pinned() { kmn_host_free(p); }
            void ensure(size_t n) {
                if (n <= cap) return;
                kmn_host_free(p);
                p = nullptr;
                cap = 0;
                void* q = nullptr;
                gpu_ok(kmn_host_alloc(&q, n));
                p = static_cast<uint8_t*>(q);
                cap = n;
            }
        } buf;
        std::vector<uint64_t> rec_off, sp_rec_off, hdr_off;
        for (size_t i0 = 0; i0 < mine.size();) {
            size_t i1 = i0, bytes = 0, recs = 0;
            while (i1 < mine.size() && (i1 == i0 || bytes + prot[mine[i1]].seq.size() <= opt.batch_bytes)) {
                bytes += prot[mine[i1]].seq.size();
                recs += prot[mine[i1]].n_rec();
                i1++;
            }
            const size_t nb = i1 - i0;
            if (bytes + recs >= (uint64_t(1) << 31)) throw std::runtime_error("species " + inputs[mine[i0]].name + " is too large for one GPU batch");
            buf.ensure(bytes);
            rec_off.assign(1, 0);
            sp_rec_off.assign(1, 0);
            size_t at = 0;
            std::vector<uint64_t> sp_at(nb + 1);
            for (size_t i = i0; i < i1; i++) {
                const Proteome& pr = prot[mine[i]];
                sp_at[i - i0] = at;
                for (size_t q = 0; q < pr.n_rec(); q++) rec_off.push_back(at + pr.rec_end[q]);
                at += pr.seq.size();
                sp_rec_off.push_back(rec_off.size() - 1);
            }
            sp_at[nb] = at;
            auto on_threads = [&](auto&& per_species) {   // per_species(index inside the batch)
                std::atomic<size_t> next{0};
                auto work = [&] { for (size_t i; (i = next.fetch_add(1)) < nb;) per_species(i); };
                std::vector<std::thread> pool;
                for (size_t t = 0; t < std::max<size_t>(1, std::min(helpers, nb)); t++) pool.emplace_back(work);
                for (auto& t : pool) t.join();
            };
            // gather the species into the pinned buffer on the host threads
            on_threads([&](size_t i) {
                const Proteome& pr = prot[mine[i0 + i]];
                if (!pr.seq.empty()) std::memcpy(buf.p + sp_at[i], pr.seq.data(), pr.seq.size());
            });
            BatchRef ref;
            ref.sids.assign(mine.begin() + i0, mine.begin() + i1);
            uint32_t id = 0;
            if (device_fasta) {
                uint64_t n_rec = 0;
                if (kmn_count_fasta(gpu, buf.p, sp_at.data(), uint32_t(nb), nullptr, &id, &n_rec) != 0) {
                    const std::string gpu_msg = kmn_last_error();
                    // a format error: the host reader words it like the reference does, first bad file in species order
                    for (size_t i = i0; i < i1; i++) {
                        try { parse_proteome(prot[mine[i]].seq, inputs[mine[i]].path); }
                        catch (const std::exception& e) { rank_err[r] = {mine[i], e.what()}; return; }
                    }
                    rank_err[r] = {mine[i0], "GPU front end: " + gpu_msg};
                    return;
                }
                hdr_off.resize(n_rec);
                sp_rec_off.assign(nb + 1, 0);
                uint64_t got = 0;
                gpu_ok(kmn_fetch_records(gpu, id, hdr_off.data(), n_rec, &got, sp_rec_off.data()));
                std::vector<std::string> idx_err(nb);
                on_threads([&](size_t i) {   // headers and sequence ranges of the records the device found
                    try {
                        index_records(prot[mine[i0 + i]], hdr_off.data() + sp_rec_off[i], size_t(sp_rec_off[i + 1] - sp_rec_off[i]), sp_at[i]);
                    } catch (const std::exception& e) { idx_err[i] = e.what(); }
                });
                for (size_t i = 0; i < nb; i++)
                    if (!idx_err[i].empty()) { rank_err[r] = {mine[i0 + i], idx_err[i]}; return; }
            } else {
                gpu_ok(kmn_count_batch(gpu, buf.p, rec_off.data(), rec_off.size() - 1, sp_rec_off.data(), uint32_t(nb), nullptr, &id));
            }
            ref.id = id;
            rank_batches[r].push_back(std::move(ref));
            i0 = i1;
        }
    });
    if (G > 1) gpu_ok(kmn_exchange_reduce_all(gpus.ctx.data(), int(G)));   // one exchange per job: every device now holds the whole table
    else gpu_ok(kmn_finalize_table(gpus.ctx[0]));                          // snapshot (:576)
    t_p1.reset();
    auto t_p2 = std::make_unique<StageTimer>("pass 2 (scan + pairing)");

    // ---- pass 2: occupancy scan, boundary rule, anchor selection and bubble pairing on the device that holds
    // the species (kmn_scan_batch + kmn_pair_batch; :589-608); the host slices the amino-acid blocks per species
    // on its threads and merges them in sid order (:611-636)
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
    on_rank_threads([&](size_t r) {
        kmn_ctx* gpu = gpus.ctx[r];
        for (const BatchRef& b : rank_batches[r]) {
            uint64_t ts = 0, te = 0, np = 0;
            gpu_ok(kmn_scan_batch(gpu, b.id, min_needed, &ts, &te));
            gpu_ok(kmn_pair_batch(gpu, b.id, uint32_t(std::min<size_t>(length_middle, 0xFFFFFFFFu)), &np));
            for (size_t i = 0; i < b.sids.size(); i++) {   // calls on one context are serialised
                uint64_t cnt = 0;
                gpu_ok(kmn_fetch_pairs(gpu, b.id, uint32_t(i), nullptr, 0, &cnt));
                outs[b.sids[i]].pairs.resize(cnt);
                if (cnt) gpu_ok(kmn_fetch_pairs(gpu, b.id, uint32_t(i), outs[b.sids[i]].pairs.data(), cnt, &cnt));
            }
        }
    });
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
        if (G == 1) {
            // one device holds every batch's anchor pairs in emission order: the device groups them in place
            std::vector<uint32_t> ids, base;
            for (const BatchRef& b : rank_batches[0]) { ids.push_back(b.id); base.push_back(uint32_t(b.sids.front())); }
            uint64_t n_dev = 0;
            gpu_ok(kmn_group_batches(gpus.ctx[0], ids.data(), uint32_t(ids.size()), base.data(), 1, min_needed, perm.data(), perm.size(),
                                     ex.groups.data(), ex.groups.size(), &n_dev, &ng));
            if (n_dev != n_obs) throw std::runtime_error("GPU front end: the device grouped " + std::to_string(n_dev) +
                                                         " observations, the host holds " + std::to_string(n_obs));
        } else {
            // the pairs live on several devices: the grouping device receives the observation keys from the host
            gpu_ok(kmn_group_observations(gpus.ctx[0], n_obs, left.data(), right.data(), sid.data(), len.data(), min_needed,
                                          perm.data(), ex.groups.data(), ex.groups.size(), &ng));
        }
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
