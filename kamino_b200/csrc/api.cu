// api.cu — context + extern "C" surface of libkamino_b200.so (see include/kamino_b200.h).
//
// Host-side orchestration of the sm_100a kernels.  There is deliberately no CPU code path: every
// entry point either runs the CUDA kernels or fails with an error message.
#include "../../include/kamino_b200.h"
#include "cms_kernels.cuh"
#include "count_kernels.cuh"
#include "fasta_kernels.cuh"
#include "frame_kernels.cuh"
#include "group_kernels.cuh"
#include "pair_kernels.cuh"
#include "scan_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

using namespace kmn;

namespace {

thread_local std::string g_err;

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define KMN_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +      \
                            __FILE__ + ":" + std::to_string(__LINE__) + ")");                       \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        KMN_CUDA(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count) { if (count > n) alloc(count); }
};

// recode.rs:40-80 as a 256-entry table (product-side copy; the oracle has its own).
RecodeLut make_lut(int scheme) {
    static const char* classes[3][6] = {
        {"C", "AGPST", "DENQ", "HKR", "ILMV", "FWY"},      // Dayhoff6  recode.rs:40-51
        {"APST", "DENG", "QKR", "MIVL", "WC", "FYH"},      // SR6       recode.rs:54-65
        {"AGPS", "DENQHKRT", "MIL", "W", "FY", "CV"},      // KGB6      recode.rs:68-80
    };
    RecodeLut lut;
    std::memset(lut.v, CODE_INVALID, sizeof lut.v);
    for (int c = 0; c < 6; c++)
        for (const char* p = classes[scheme][c]; *p; p++) {
            lut.v[uint8_t(*p)] = uint8_t(c);
            lut.v[uint8_t(*p) + 32] = uint8_t(c);  // lower case (to_ascii_uppercase, group_extraction.rs:379)
        }
    for (uint8_t ws : {0x20, 0x09, 0x0A, 0x0C, 0x0D}) lut.v[ws] = CODE_SKIP;  // u8::is_ascii_whitespace
    return lut;
}

struct Batch {
    uint64_t n_raw = 0;
    uint32_t n_rec = 0, n_sp = 0;
    size_t raw_pad = 0;      // n_raw rounded up to FR_TILE
    size_t codes_cap = 0;    // allocated codes bytes (multiple of WARP_TILE)
    uint32_t n_codes = 0;    // stripped residues + n_rec separators (known after framing)
    bool framed = false;
    DevBuf<uint8_t> raw, codes;
    DevBuf<uint64_t> rec_off, sp_rec_off;
    DevBuf<uint64_t> hdr_off;    // kmn_count_fasta: offset of each record's '>' in the caller's bytes
    bool from_fasta = false;
    DevBuf<uint32_t> rec_cstart, sp_cstart;
    std::vector<uint32_t> h_sp_cstart;   // host copy (after framing)
    std::vector<uint64_t> h_sp_rec_off;
    std::vector<uint64_t> h_sp_raw_off;  // raw byte offset of each species (plans the copy / compute ranges)
    // pass 2 results
    bool scanned = false;
    bool occ_narrow = false;                 // occ[] came from the narrow table: exact where >= min_needed, 0 below
    DevBuf<uint16_t> occ;
    DevBuf<uint32_t> inv_pos;
    uint32_t n_inv = 0;
    DevBuf<uint2> starts, ends;              // compact hits: (start inside the segment, occupancy)
    DevBuf<kmn_hit> full_starts, full_ends;  // kmn_hit form, expanded on the first kmn_fetch_hits
    bool hits_expanded = false;
    uint32_t tot_starts = 0, tot_ends = 0;
    DevBuf<uint32_t> seg_s_off, seg_e_off;   // per-segment offsets into starts / ends (n_inv + 1)
    std::vector<uint32_t> h_sp_starts_off, h_sp_ends_off;
    // anchor pairs (kmn_pair_batch)
    bool paired = false;
    DevBuf<kmn_pair> pairs;
    std::vector<uint32_t> h_sp_pairs_off;
    // kmn_stage_batch_async: the batch's host-to-device copies were queued on the copy stream; the first consumer waits
    cudaEvent_t staged_ev = nullptr;
    bool staged_pending = false;
    const uint64_t* pending_rec_off = nullptr;   // the caller's rec_off, not yet checked for order (check_deferred_offsets)
    bool bad_offsets = false;
    ~Batch() { if (staged_ev) cudaEventDestroy(staged_ev); }
};

}  // namespace

constexpr uint32_t kMaxH2dChunks = 32;
constexpr float kKernelGBs = 41.f;     // raw bytes per second the frame / Bloom / partition kernels of a (small) range consume

struct kmn_ctx {
    int device = 0;
    int k = 13;
    int scheme = KMN_SR6;
    uint64_t k_mask = 0;
    RecodeLut lut;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    DevBuf<uint16_t> tab;                 // 4 x 2^26 saturating u16 cells (512 MiB): THE count-min table
    DevBuf<uint16_t> xtab;                // table after the fused multi-GPU exchange (allocated on setup)
    DevBuf<uint32_t> wide;                // u32 widening of the table for the NCCL variant (allocated on demand)
    bool widened = false;
    const uint16_t* final_tab = nullptr;  // what pass 2 reads once finalized (tab or xtab)
    // narrow copy of the finalized table for pass 2 (scan_kernels.cuh): valid for (tab8_src, tab8_base); tab_max_src: the
    // table whose largest cell is tab_max.  Both are forgotten whenever final_tab is (re)assigned.
    DevBuf<uint8_t> tab8;
    DevBuf<uint32_t> tab_max_dev;
    const uint16_t* tab8_src = nullptr;
    const uint16_t* tab_max_src = nullptr;
    uint32_t tab8_base = 0, tab_max = 0;
    bool scan_wide = false;               // test / tuning knob KMN_SCAN_WIDE: pass 2 always reads the u16 table (8 sweeps)
    bool tab_zero = false;                // kmn_reset_table: tab is LOGICALLY zero, its memory is stale (materialize_table)
    bool finalized = false;
    DevBuf<unsigned long long> new_counts;
    DevBuf<uint32_t> tile_cnt, tile_base, tile_rec, seg_s_cnt, seg_e_cnt, sp_s_off, sp_e_off;
    std::vector<std::unique_ptr<Batch>> batches;
    PeerTables peers{};                   // mapped peer buffers of the fused exchange
    PeerPlanes planes{};                  // the same peers, byte-plane views (default exchange)
    bool exchange_planes = true;          // KMN_EXCHANGE_PLANES=0: exchange whole u16 cells instead
    bool peers_ready = false;
    uint32_t exchange_epoch = 0;          // exchanges run so far (every rank counts alike)
    DevBuf<uint32_t> xsync;               // [0] go word, [1] finished-CTA counter, [2] error word of the exchange kernels
    uint32_t* xabort_host = nullptr;      // pinned word (== 1): source of the copy that raises the device abort word xsync[3]
    uint32_t exchange_timeout_ms = 30000; // KMN_EXCHANGE_TIMEOUT_MS: a rank waits this long for a peer before it gives up
    int exchange_grid = 0;
    std::vector<void*> ipc_opened;
    std::vector<std::unique_ptr<Batch>> spare;   // released batches whose buffers get reused
    cudaStream_t copy_stream = nullptr;   // host -> device copies of kmn_count_batch, overlapped with the kernels
    cudaStream_t frame_stream = nullptr;  // K0 of range j+1 runs here, beside the Bloom / count-min kernels of range j
    cudaEvent_t ev_frame[1 + kMaxH2dChunks] = {};
    cudaEvent_t ev[12] = {};
    cudaEvent_t ev_copy[1 + kMaxH2dChunks] = {};
    bool h2d_chunks_forced = false;       // KMN_H2D_CHUNKS given: no automatic schedule for long batches
    uint32_t h2d_chunks = 5;              // species ranges per kmn_count_batch (KMN_H2D_CHUNKS): every range costs ~50 us of launch gaps and tails
    float h2d_ratio = 0.f;                // KMN_H2D_RATIO: growth of consecutive ranges (0: from the measured link rate)
    cudaEvent_t ev_h2d[2] = {};           // first / last host-to-device range copy of the last kmn_count_batch
    float h2d_gbs = 0.f;                  // measured link rate of that call (0: not measured yet)
    int h2d_mode = 0;                     // KMN_H2D_MODE: 0 adaptive (geometric ranges), 1 doubling ranges, 2 equal ranges
    DevBuf<uint32_t> frame_ticket, tile_blo, n_tiles_dev, scan_sums, scan_offs;
    DevBuf<unsigned long long> tile_state;
    // scratch of kmn_group_observations
    DevBuf<uint64_t> g_left, g_right;
    DevBuf<uint32_t> g_sid, g_len, g_perm[2], g_hist, g_base, g_head, g_scan, g_first, g_spoff;
    DevBuf<kmn_group> g_groups;
    DevBuf<uint64_t> fa_file_off;          // scratch of kmn_count_fasta
    DevBuf<uint8_t> fa_tile_sum, fa_tile_in;
    DevBuf<uint32_t> fa_tile_hdr, fa_tile_base, fa_err;
    DevBuf<kmn_pair> pair_scratch;        // kmn_pair_batch: pairs before compaction (one slot per start candidate)
    DevBuf<uint32_t> nj_valid, nj_mism;   // counts of the last pair_counts call (kmn_pair_counts_device)
    DevBuf<uint8_t> nj_mapped;            // scratch of the pair counts, kept between calls (no cudaMalloc per call)
    DevBuf<uint32_t> nj_planes;
    uint32_t nj_rows = 0, nj_n = 0;
    float stage_ms[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t launches = 0;
    // Bloom stage scratch (count_kernels.cuh): probe runs per (source tile, bucket), tile tables, flags
    DevBuf<uint32_t> bl_ent, bl_tile_sp, bl_tile0, bl_slow, bl_lost2, bl_scratch;
    DevBuf<uint16_t> bl_cnt;
    DevBuf<unsigned long long> bl_ovf;
    DevBuf<uint32_t> bl_ovf_n;
    uint32_t bloom_run_cap = RUN_CAP;     // test knob KMN_BLOOM_RUN_CAP: small slots force the slow exact path
    DevBuf<uint16_t> cms_ent;             // (bucket, tile) slots of the CMS partition (scratch of count_staged)
    DevBuf<uint32_t> cms_running, cms_spill, cms_state, cms_tile_sp, cms_tile0;
    int cms_force = 0;                    // test knob: 1 = exact (u32) redo in every bucket, 2 = direct fallback
    uint32_t cms_cap_override = 0;        // test knob: spill list capacity (forces the spill-full fallback)
    uint32_t cms_slot_limit = PART_SLOT_CAP;  // test knob: staged entries per bucket and tile (forces spills)
};

namespace {

struct DeviceGuard {
    explicit DeviceGuard(int dev) { KMN_CUDA(cudaSetDevice(dev)); }
};

void check_launch(kmn_ctx* c) {
    c->launches++;
    KMN_CUDA(cudaGetLastError());
}

int grid_for(kmn_ctx* c, int ctas_per_sm) { return c->sm_count * ctas_per_sm; }

// TMA descriptor of the partition's slot array ent[CMS_BUCKETS][tiles_cap][PART_SLOT] (u16): one box = the slots
// of 256 consecutive buckets for one tile.  cuTensorMapEncodeTiled is resolved through the runtime, so the
// library does not link against libcuda.
CUtensorMap make_slots_map(uint16_t* ent, uint32_t tiles_cap) {
    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        KMN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) throw std::runtime_error("cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    CUtensorMap m;
    const cuuint64_t dims[3] = {PART_SLOT, tiles_cap, CMS_BUCKETS};
    const cuuint64_t strides[2] = {PART_SLOT * sizeof(uint16_t), uint64_t(tiles_cap) * PART_SLOT * sizeof(uint16_t)};
    const cuuint32_t box[3] = {PART_SLOT, 1, PART_BOX_BUCKETS};
    const cuuint32_t elem[3] = {1, 1, 1};
    const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, ent, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
    return m;
}

// out[0..n] = exclusive scan of in[0..n) (out[n] = total) on the context's stream: one CTA for short arrays,
// per-tile sums + their scan + per-tile scans for long ones.
void device_exclusive_scan(kmn_ctx* c, const uint32_t* in, uint32_t* out, uint32_t n) {
    if (n <= 4 * SCAN_TILE) {
        const ScanJob job{in, out, n, nullptr};
        exclusive_scan_kernel<<<1, 1024, 0, c->stream>>>(job, job);
        check_launch(c);
        return;
    }
    const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    c->scan_sums.ensure(tiles);
    c->scan_offs.ensure(size_t(tiles) + 1);
    scan_tile_sums_kernel<<<tiles, 1024, 0, c->stream>>>(in, n, c->scan_sums.p);
    check_launch(c);
    const ScanJob job{c->scan_sums.p, c->scan_offs.p, tiles, nullptr};
    exclusive_scan_kernel<<<1, 1024, 0, c->stream>>>(job, job);
    check_launch(c);
    scan_tiles_kernel<<<tiles, 1024, 0, c->stream>>>(in, n, c->scan_offs.p, out);
    check_launch(c);
}

Batch& get_batch(kmn_ctx* c, uint32_t id) {
    if (id >= c->batches.size() || !c->batches[id]) throw std::runtime_error("unknown batch id");
    return *c->batches[id];
}

// Validates the offsets, takes a (recycled) Batch and sizes its buffers.  No device work.
std::unique_ptr<Batch> take_batch(kmn_ctx* c) {
    std::unique_ptr<Batch> b;
    if (!c->spare.empty()) {
        b = std::move(c->spare.back());
        c->spare.pop_back();
        b->framed = b->scanned = b->paired = b->hits_expanded = b->from_fasta = b->occ_narrow = false;
        b->n_codes = b->n_inv = 0;
    } else {
        b = std::make_unique<Batch>();
    }
    return b;
}

// rec_off must be non-decreasing: one pass over 8 bytes per record (~0.15 ms per 100 proteomes from host memory).
// Branch-free so that it vectorises.
bool rec_off_sorted(const uint64_t* rec_off, uint64_t n_rec) {
    uint64_t bad = 0;
    for (uint64_t r = 0; r < n_rec; r++) bad |= uint64_t(rec_off[r + 1] < rec_off[r]);
    return bad == 0;
}

// kmn_stage_batch_async defers that pass: it would run while the GPU idles between two counts.  It runs instead
// inside the next count call of the context, after that call's kernels have been queued (run_pass1), or right
// before the batch itself is used, whichever comes first; a batch that fails it refuses to be used.
void check_deferred_offsets(Batch& b) {
    if (!b.pending_rec_off) return;
    b.bad_offsets = !rec_off_sorted(b.pending_rec_off, b.n_rec);
    b.pending_rec_off = nullptr;
}

std::unique_ptr<Batch> new_batch(kmn_ctx* c, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                                 const uint64_t* sp_rec_off, uint32_t n_sp, bool defer_offset_check = false) {
    if (!rec_off || !sp_rec_off) throw std::runtime_error("rec_off / sp_rec_off must not be NULL");
    if (n_rec >= (1ull << 31)) throw std::runtime_error("too many records in one batch");
    if (n_sp > 65535) throw std::runtime_error("too many species in one batch (maximum 65535)");  // group_extraction.rs:538
    if (rec_off[0] != 0) throw std::runtime_error("rec_off[0] must be 0");
    const uint64_t n_raw = rec_off[n_rec];
    if (n_raw + n_rec >= (1ull << 31)) throw std::runtime_error("batch too large: residues + records must stay below 2^31");
    if (n_raw && !residues) throw std::runtime_error("residues must not be NULL");
    if (!defer_offset_check && !rec_off_sorted(rec_off, n_rec)) throw std::runtime_error("rec_off must be non-decreasing");
    if (sp_rec_off[0] != 0 || sp_rec_off[n_sp] != n_rec) throw std::runtime_error("sp_rec_off must span [0, n_rec]");
    for (uint32_t s = 0; s < n_sp; s++)
        if (sp_rec_off[s + 1] < sp_rec_off[s]) throw std::runtime_error("sp_rec_off must be non-decreasing");

    // staged batches recycle the device buffers of released ones (cudaMalloc/cudaFree of ~1 GB per
    // batch would otherwise dominate the host-to-table path)
    std::unique_ptr<Batch> b = take_batch(c);
    b->n_raw = n_raw;
    b->n_rec = uint32_t(n_rec);
    b->n_sp = n_sp;
    b->pending_rec_off = defer_offset_check ? rec_off : nullptr;
    b->bad_offsets = false;
    b->raw_pad = (size_t(n_raw) / FR_TILE + 1) * FR_TILE;   // at least one pad byte: offset n_raw has a chunk
    b->raw.ensure(b->raw_pad);
    b->rec_off.ensure(n_rec + 1);
    b->sp_rec_off.ensure(size_t(n_sp) + 1);
    b->h_sp_rec_off.assign(sp_rec_off, sp_rec_off + n_sp + 1);
    b->h_sp_raw_off.resize(size_t(n_sp) + 1);
    for (uint32_t s = 0; s <= n_sp; s++) b->h_sp_raw_off[s] = rec_off[sp_rec_off[s]];
    // worst case: every raw byte is a residue, plus one separator per record
    b->codes_cap = (size_t(n_raw) + n_rec + WARP_TILE - 1) / WARP_TILE * WARP_TILE + WARP_TILE;
    b->codes.ensure(b->codes_cap);
    b->rec_cstart.ensure(n_rec + 1);
    b->sp_cstart.ensure(size_t(n_sp) + 1);
    return b;
}

void stage_batch(kmn_ctx* c, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                 const uint64_t* sp_rec_off, uint32_t n_sp, uint32_t* batch_id) {
    std::unique_ptr<Batch> b = new_batch(c, residues, rec_off, n_rec, sp_rec_off, n_sp);
    if (b->n_raw) KMN_CUDA(cudaMemcpyAsync(b->raw.p, residues, b->n_raw, cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaMemsetAsync(b->raw.p + b->n_raw, 0x20, b->raw_pad - b->n_raw, c->stream));  // spaces: skipped like any whitespace
    KMN_CUDA(cudaMemcpyAsync(b->rec_off.p, rec_off, (n_rec + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaMemcpyAsync(b->sp_rec_off.p, sp_rec_off, (size_t(n_sp) + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));  // the caller may free its buffers on return
    c->batches.push_back(std::move(b));
    if (batch_id) *batch_id = uint32_t(c->batches.size() - 1);
}

// kmn_stage_batch_async: the same copies on the copy stream, no wait.  The compute stream waits for them when the
// batch is first used (run_pass1), so the copies of batch i+1 overlap the kernels of batch i in a multi-batch job.
void stage_batch_async(kmn_ctx* c, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                       const uint64_t* sp_rec_off, uint32_t n_sp, uint32_t* batch_id) {
    std::unique_ptr<Batch> b = new_batch(c, residues, rec_off, n_rec, sp_rec_off, n_sp, /*defer_offset_check=*/true);
    // a recycled batch's buffers may still be read by work queued on the compute stream
    KMN_CUDA(cudaEventRecord(c->ev_copy[0], c->stream));
    KMN_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_copy[0], 0));
    KMN_CUDA(cudaMemcpyAsync(b->rec_off.p, rec_off, (n_rec + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->copy_stream));
    KMN_CUDA(cudaMemcpyAsync(b->sp_rec_off.p, sp_rec_off, (size_t(n_sp) + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->copy_stream));
    if (b->n_raw) KMN_CUDA(cudaMemcpyAsync(b->raw.p, residues, b->n_raw, cudaMemcpyHostToDevice, c->copy_stream));
    KMN_CUDA(cudaMemsetAsync(b->raw.p + b->n_raw, 0x20, b->raw_pad - b->n_raw, c->copy_stream));
    if (!b->staged_ev) KMN_CUDA(cudaEventCreateWithFlags(&b->staged_ev, cudaEventDisableTiming));
    KMN_CUDA(cudaEventRecord(b->staged_ev, c->copy_stream));
    b->staged_pending = true;
    c->batches.push_back(std::move(b));
    if (batch_id) *batch_id = uint32_t(c->batches.size() - 1);
}

// kmn_count_fasta, front half: whole FASTA files -> (header-blanked raw bytes, rec_off, sp_rec_off) on the
// device (fasta_kernels.cuh).  Two small host round trips: the record count sizes the record tables, the
// per-file record ranges plan pass 1.
std::unique_ptr<Batch> frame_fasta(kmn_ctx* c, const uint8_t* bytes, const uint64_t* file_off, uint32_t n_files) {
    if (!file_off) throw std::runtime_error("file_off must not be NULL");
    if (n_files > 65535) throw std::runtime_error("too many species in one batch (maximum 65535)");  // group_extraction.rs:538
    if (file_off[0] != 0) throw std::runtime_error("file_off[0] must be 0");
    for (uint32_t f = 0; f < n_files; f++)
        if (file_off[f + 1] < file_off[f]) throw std::runtime_error("file_off must be non-decreasing");
    const uint64_t n_raw = file_off[n_files];
    if (n_raw >= (1ull << 31)) throw std::runtime_error("batch too large: bytes + records must stay below 2^31");
    if (n_raw && !bytes) throw std::runtime_error("bytes must not be NULL");

    std::unique_ptr<Batch> b = take_batch(c);
    b->from_fasta = true;
    b->n_raw = n_raw;
    b->n_sp = n_files;
    b->raw_pad = (size_t(n_raw) / FR_TILE + 1) * FR_TILE;
    b->raw.ensure(b->raw_pad);
    const uint32_t n_tiles = uint32_t(b->raw_pad / FRAME_TILE);
    c->fa_file_off.ensure(size_t(n_files) + 1);
    c->fa_tile_sum.ensure(n_tiles);
    c->fa_tile_in.ensure(n_tiles);
    c->fa_tile_hdr.ensure(n_tiles);
    c->fa_tile_base.ensure(size_t(n_tiles) + 1);
    c->fa_err.ensure(1);
    if (n_raw) KMN_CUDA(cudaMemcpyAsync(b->raw.p, bytes, n_raw, cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaMemsetAsync(b->raw.p + n_raw, 0x20, b->raw_pad - n_raw, c->stream));
    KMN_CUDA(cudaMemcpyAsync(c->fa_file_off.p, file_off, (size_t(n_files) + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaMemsetAsync(c->fa_err.p, 0xFF, sizeof(uint32_t), c->stream));
    if (n_files) {
        fasta_check_kernel<<<(n_files + 127) / 128, 128, 0, c->stream>>>(b->raw.p, c->fa_file_off.p, n_files, c->fa_err.p);
        check_launch(c);
    }
    fasta_summary_kernel<<<n_tiles, FRAME_THREADS, 0, c->stream>>>(b->raw.p, c->fa_tile_sum.p, c->fa_tile_hdr.p);
    check_launch(c);
    device_exclusive_scan(c, c->fa_tile_hdr.p, c->fa_tile_base.p, n_tiles);
    fasta_tile_state_kernel<<<1, 1024, 0, c->stream>>>(c->fa_tile_sum.p, n_tiles, c->fa_tile_in.p);
    check_launch(c);
    uint32_t n_rec = 0, err = 0;
    KMN_CUDA(cudaMemcpyAsync(&n_rec, c->fa_tile_base.p + n_tiles, sizeof n_rec, cudaMemcpyDeviceToHost, c->stream));
    KMN_CUDA(cudaMemcpyAsync(&err, c->fa_err.p, sizeof err, cudaMemcpyDeviceToHost, c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    if (err != 0xFFFFFFFFu) {
        c->spare.push_back(std::move(b));
        const uint32_t f = (err & 0x7FFFFFFFu) - 1;
        if (err & 0x80000000u)
            throw std::runtime_error("file " + std::to_string(f) + " does not end with a line feed (append one before kmn_count_fasta)");
        throw std::runtime_error("FASTA parse error in file " + std::to_string(f) + ": expected '>' at record start");
    }
    if (n_raw + n_rec >= (1ull << 31)) {
        c->spare.push_back(std::move(b));
        throw std::runtime_error("batch too large: bytes + records must stay below 2^31");
    }
    b->n_rec = n_rec;
    b->rec_off.ensure(size_t(n_rec) + 1);
    b->hdr_off.ensure(size_t(n_rec) + 1);
    b->sp_rec_off.ensure(size_t(n_files) + 1);
    fasta_blank_kernel<<<n_tiles, FRAME_THREADS, 0, c->stream>>>(b->raw.p, c->fa_tile_in.p, c->fa_tile_base.p, b->hdr_off.p, b->rec_off.p);
    check_launch(c);
    fasta_species_kernel<<<(n_files + 1 + 127) / 128, 128, 0, c->stream>>>(b->hdr_off.p, n_rec, c->fa_file_off.p, n_files, n_raw,
                                                                         b->sp_rec_off.p, b->rec_off.p);
    check_launch(c);
    b->h_sp_rec_off.resize(size_t(n_files) + 1);
    KMN_CUDA(cudaMemcpyAsync(b->h_sp_rec_off.data(), b->sp_rec_off.p, (size_t(n_files) + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                             c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    b->h_sp_raw_off.assign(file_off, file_off + n_files + 1);
    b->codes_cap = (size_t(n_raw) + n_rec + WARP_TILE - 1) / WARP_TILE * WARP_TILE + WARP_TILE;
    b->codes.ensure(b->codes_cap);
    b->rec_cstart.ensure(size_t(n_rec) + 1);
    b->sp_cstart.ensure(size_t(n_files) + 1);
    return b;
}

// ---- pass 1 ---------------------------------------------------------------------------------------
// One species range of a batch: its raw bytes arrive (or already sit) in HBM, then frame -> Bloom ->
// count-min partition run on it.  Nothing here needs a host round trip: tile tables are built on the device
// and grids are sized from host-side upper bounds, so with host residues the copy of range j+1 (copy stream)
// overlaps the kernels of range j (compute stream).
struct SpeciesRange {
    uint32_t s0, s1;                // species [s0, s1)
    uint64_t copy_begin, copy_end;  // raw bytes that must be resident before this range is framed
    uint32_t tile_begin, tile_end;  // frame tiles (FR_TILE raw bytes) processed with this range
    uint64_t codes_bound;           // upper bound of the codes[] positions of the range
};

std::vector<SpeciesRange> plan_ranges(const Batch& b, uint32_t want, double ratio, bool long_schedule) {
    std::vector<SpeciesRange> out;
    const uint32_t n_frame_tiles = uint32_t(b.raw_pad / FR_TILE);
    if (b.n_sp == 0) {   // records without species cannot exist (sp_rec_off spans them); frame only
        out.push_back({0, 0, 0, b.n_raw, 0, n_frame_tiles, 0});
        return out;
    }
    want = std::max<uint32_t>(1, std::min<uint32_t>(want, b.n_sp));
    // Geometric ranges, each `ratio` times the one before.  Only the first range's copy is exposed, so it should be
    // small; after that the copy of range j+1 has to land while the kernels of range j run, or the GPU idles: the
    // ranges may grow by at most (link rate / rate at which the kernels consume raw bytes).  Measured (KMN_TRACE_RANGES,
    // 55 GB/s link): doubling ranges left the GPU idle for 0.5 ms of a 4.8 ms call waiting for the last two copies;
    // 5 ranges growing by 1.5 take 4.69 ms, by 2.0 4.86 ms, equal ones 4.86 ms, 12 ranges > 5.0 ms
    // (profiles/r02_e2e_range_schedules.jsonl).  Ratio 1 (equal ranges) is right when the link is the slower side
    // (several ranks sharing the host's PCIe / memory path): then the kernels of the LAST range are what is left
    // after the copies.
    // Long batches (> 160 MB: a 1000-proteome job in one call) keep that ramp — 9 MB, then `ratio` times more each —
    // up to 48 MB and continue with ranges of that size: with 5 ranges the last one alone would be 38 % of the
    // batch, and its kernels would only start when the whole batch has been copied.
    std::vector<double> w;
    if (long_schedule) {
        double size = 9e6, left = double(b.n_raw);
        while (left > 0 && w.size() < kMaxH2dChunks) {
            w.push_back(size);
            left -= size;
            size = std::min(size * std::max(ratio, 1.0), 48e6);
        }
        if (left > 0) w.back() += left;   // (more than kMaxH2dChunks ranges: the last one takes the rest)
        want = std::max<uint32_t>(1, std::min<uint32_t>(uint32_t(w.size()), b.n_sp));
    } else {
        w.resize(want);
        for (uint32_t j = 0; j < want; j++) w[j] = j == 0 ? 1.0 : w[j - 1] * ratio;
    }
    double wsum = 0;
    for (double x : w) wsum += x;
    auto range_bytes = [&](size_t j) { return uint64_t(double(b.n_raw) * w[std::min<size_t>(j, w.size() - 1)] / wsum) + 1; };
    uint32_t s0 = 0;
    uint64_t copied = 0;
    uint32_t tiled = 0;
    while (s0 < b.n_sp) {
        uint32_t s1 = s0 + 1;
        if (out.size() + 1 == want) s1 = b.n_sp;
        else {
            const uint64_t per = range_bytes(out.size());
            while (s1 < b.n_sp && b.h_sp_raw_off[s1] - b.h_sp_raw_off[s0] < per) s1++;
        }
        SpeciesRange r;
        r.s0 = s0;
        r.s1 = s1;
        // the frame tile that holds the range's closing record boundary must be complete
        const uint64_t end_byte = b.h_sp_raw_off[s1];
        const bool last = s1 == b.n_sp;
        r.copy_begin = copied;
        r.copy_end = last ? b.n_raw : std::min<uint64_t>(b.n_raw, (end_byte / FR_TILE + 1) * FR_TILE);
        r.copy_end = std::max(r.copy_end, r.copy_begin);
        r.tile_begin = tiled;
        r.tile_end = last ? n_frame_tiles : std::max<uint32_t>(tiled, uint32_t(r.copy_end / FR_TILE));
        if (r.copy_end == b.n_raw) r.tile_end = n_frame_tiles;   // the pad tile carries the boundaries at n_raw
        const uint64_t recs = b.h_sp_rec_off[s1] - b.h_sp_rec_off[s0];
        r.codes_bound = (b.h_sp_raw_off[s1] - b.h_sp_raw_off[s0]) + recs;
        copied = r.copy_end;
        tiled = r.tile_end;
        out.push_back(r);
        s0 = s1;
    }
    return out;
}

// host_residues / host_rec_off / host_sp_rec_off != NULL: the batch's inputs still live on the host
// (kmn_count_batch); NULL: staged before (kmn_count_staged).  frame_only: K0 alone (pass 2 on a loaded table).
void run_pass1(kmn_ctx* c, Batch& b, const uint8_t* host_residues, const uint64_t* host_rec_off,
               const uint64_t* host_sp_rec_off, bool frame_only, uint64_t* new_kmers_per_species) {
    if (!frame_only && c->finalized) throw std::runtime_error("table is finalized; call kmn_reset_table before counting again");
    const bool from_host = host_rec_off != nullptr;
    if (!from_host) {   // (from host buffers: checked below, under the first range's copy)
        check_deferred_offsets(b);
        if (b.bad_offsets) throw std::runtime_error("rec_off must be non-decreasing");
    }
    if (b.staged_pending) {   // kmn_stage_batch_async: the copies run on the copy stream
        KMN_CUDA(cudaStreamWaitEvent(c->stream, b.staged_ev, 0));
        b.staged_pending = false;
    }
    // range schedule: doubling sizes while the link outruns the kernels (~27 GB/s of raw bytes), equal sizes and
    // more of them when the previous call measured a slower link
    // range schedule: growth ratio from the link rate the previous call measured (55 GB/s when none has been)
    double ratio = c->h2d_ratio > 0.f ? c->h2d_ratio : std::min(2.0, std::max(1.0, double(c->h2d_gbs > 0.f ? c->h2d_gbs : 55.f) / kKernelGBs));
    if (c->h2d_mode == 1) ratio = 2.0;
    if (c->h2d_mode == 2) ratio = 1.0;
    const uint32_t chunks = c->h2d_chunks;
    const uint32_t want = from_host ? std::max<uint32_t>(1, std::min<uint32_t>(chunks, uint32_t(b.n_raw >> 22))) : 1;
    const bool long_schedule = from_host && !c->h2d_chunks_forced && c->h2d_mode == 0 && b.n_raw > (uint64_t(160) << 20);
    const std::vector<SpeciesRange> ranges = plan_ranges(b, want, ratio, long_schedule);
    const bool timed = ranges.size() == 1;   // per-stage events only make sense without the pipeline
    const int grid = grid_for(c, 8);

    // KMN_TRACE_RANGES=1: device timeline of the copy / frame / compute streams of this call on stderr (tuning aid)
    static const bool trace = std::getenv("KMN_TRACE_RANGES") != nullptr;
    struct Mark { cudaEvent_t e; const char* what; size_t j; };
    std::vector<Mark> marks;
    auto mark = [&](cudaStream_t st, const char* what, size_t j) {
        if (!trace) return;
        cudaEvent_t e;
        KMN_CUDA(cudaEventCreate(&e));
        KMN_CUDA(cudaEventRecord(e, st));
        marks.push_back({e, what, j});
    };
    KMN_CUDA(cudaEventRecord(c->ev[0], c->stream));
    mark(c->stream, "start", 0);
    if (from_host) {
        // (rec_off — 8 bytes per record, 3 MB per 100 proteomes — travels in slices with the ranges' raw bytes)
        KMN_CUDA(cudaMemcpyAsync(b.sp_rec_off.p, host_sp_rec_off, (size_t(b.n_sp) + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        KMN_CUDA(cudaEventRecord(c->ev_copy[0], c->stream));
        KMN_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_copy[0], 0));   // the raw buffer's previous users are done
    }
    if (b.n_rec == 0) KMN_CUDA(cudaMemsetAsync(b.codes.p, CODE_INVALID, b.codes_cap, c->stream));   // else: frame_kernel + codes_tail_kernel write every byte
    const uint32_t n_frame_tiles = uint32_t(b.raw_pad / FR_TILE);
    c->tile_rec.ensure(n_frame_tiles + 2);
    c->tile_blo.ensure(n_frame_tiles + 2);
    c->tile_state.ensure(n_frame_tiles + 2);
    // look-back states of the frame tiles (the chain of kept bytes runs through all ranges of the call) and one
    // ticket counter per range
    KMN_CUDA(cudaMemsetAsync(c->tile_state.p, 0, (size_t(n_frame_tiles) + 2) * sizeof(unsigned long long), c->stream));
    KMN_CUDA(cudaMemsetAsync(c->frame_ticket.p, 0, (kMaxH2dChunks + 1) * sizeof(uint32_t), c->stream));
    if (b.n_rec == 0) KMN_CUDA(cudaMemsetAsync(b.rec_cstart.p, 0, sizeof(uint32_t), c->stream));

    BloomRuns R{};
    CmsTiles T{};
    CmsSlots L{};
    CUtensorMap slots_map{};
    if (!frame_only) {
        c->new_counts.ensure(size_t(b.n_sp) + 1);
        KMN_CUDA(cudaMemsetAsync(c->new_counts.p, 0, (size_t(b.n_sp) + 1) * sizeof(unsigned long long), c->stream));
        uint64_t max_codes = 0;
        uint32_t max_sp = 0;
        for (const SpeciesRange& r : ranges) {
            max_codes = std::max(max_codes, r.codes_bound);
            max_sp = std::max(max_sp, r.s1 - r.s0);
        }
        // source tiles of a range: codes / tile size, plus two partial tiles per species
        const uint32_t bl_tiles = uint32_t(max_codes / BP_TILE) + 2 * max_sp + 2;
        const uint32_t cms_tiles = uint32_t(max_codes / PART_TILE) + 2 * max_sp + 2;
        R.cap = c->bloom_run_cap;
        c->bl_tile_sp.ensure(bl_tiles);
        c->bl_tile0.ensure(size_t(b.n_sp) + 1);
        c->bl_slow.ensure(size_t(b.n_sp) + 1);
        c->bl_ent.ensure(size_t(bl_tiles) * BB_BUCKETS * R.cap);
        c->bl_cnt.ensure(size_t(bl_tiles) * BB_BUCKETS);
        c->bl_ovf.ensure(size_t(bl_tiles) * OVF_CAP);
        c->bl_ovf_n.ensure(bl_tiles);
        c->bl_lost2.ensure(b.codes_cap / CHUNK);
        c->cms_tile_sp.ensure(cms_tiles);
        c->cms_tile0.ensure(size_t(b.n_sp) + 1);
        KMN_CUDA(cudaMemsetAsync(c->bl_slow.p, 0, (size_t(b.n_sp) + 1) * sizeof(uint32_t), c->stream));
        KMN_CUDA(cudaMemsetAsync(c->bl_lost2.p, 0, (b.codes_cap / CHUNK) * sizeof(uint32_t), c->stream));
        R.ent = c->bl_ent.p;
        R.cnt = c->bl_cnt.p;
        R.ovf = c->bl_ovf.p;
        R.ovf_n = c->bl_ovf_n.p;
        R.tile_sp = c->bl_tile_sp.p;
        R.sp_tile0 = c->bl_tile0.p;
        R.n_tiles = c->n_tiles_dev.p;
        R.slow = c->bl_slow.p;
        R.lost2 = c->bl_lost2.p;
        T.tile_sp = c->cms_tile_sp.p;
        T.sp_tile0 = c->cms_tile0.p;
        T.n_tiles = c->n_tiles_dev.p + 1;
        // one fixed slot per (bucket, tile): the tiles of all ranges of the batch, each range bounded like its grid.
        // A slot that overflows spills single entries; a full spill list is not an error: the batch then takes
        // the exact direct path (cms_direct_kernel).
        const uint64_t all_codes = b.n_raw + b.n_rec;
        uint64_t tiles_cap = 0;
        for (const SpeciesRange& r : ranges) tiles_cap += r.codes_bound / PART_TILE + 2 * uint64_t(r.s1 - r.s0) + 2;
        L.tiles_cap = uint32_t(tiles_cap);
        L.spill_cap = c->cms_cap_override ? c->cms_cap_override : uint32_t(all_codes / 64 + 65536);
        c->cms_ent.ensure(size_t(CMS_BUCKETS) * L.tiles_cap * PART_SLOT);
        c->cms_spill.ensure(L.spill_cap);
        L.ent = c->cms_ent.p;
        L.running = c->cms_running.p;
        L.spill = c->cms_spill.p;
        L.state = c->cms_state.p;
        slots_map = make_slots_map(L.ent, L.tiles_cap);
        KMN_CUDA(cudaMemsetAsync(c->cms_running.p, 0, 2 * sizeof(uint32_t), c->stream));
        KMN_CUDA(cudaMemsetAsync(L.state, 0, 2 * sizeof(uint32_t), c->stream));
    }

    if (from_host)   // pad with spaces: skipped like any whitespace (disjoint from the range copies)
        KMN_CUDA(cudaMemsetAsync(b.raw.p + b.n_raw, 0x20, b.raw_pad - b.n_raw, c->copy_stream));
    // With several ranges the frame kernels (K0: a chain of small launches around two streaming kernels) run on
    // their own stream: range j+1 is framed while the Bloom / count-min kernels of range j keep the SMs busy, so
    // the serial tail of the small launches disappears behind them.  fs == c->stream when there is one range.
    cudaStream_t fs = ranges.size() > 1 ? c->frame_stream : c->stream;
    if (fs != c->stream) {   // everything queued so far (offset copies, memsets of codes / carry / scratch) precedes the first frame
        KMN_CUDA(cudaEventRecord(c->ev_frame[0], c->stream));
        KMN_CUDA(cudaStreamWaitEvent(fs, c->ev_frame[0], 0));
    }
    uint64_t rec_copied = 0;   // rec_off[0, rec_copied) is on its way
    for (size_t j = 0; j < ranges.size(); j++) {
        const SpeciesRange& r = ranges[j];
        // the range's frame kernels read rec_off up to the first boundary behind the range's last raw byte
        uint32_t rec_hi = b.n_rec;
        if (from_host) {   // copy stream: this range's record offsets and raw bytes, then the frame stream may frame them
            if (j + 1 < ranges.size())
                rec_hi = uint32_t(std::min<uint64_t>(b.n_rec, uint64_t(std::upper_bound(host_rec_off, host_rec_off + b.n_rec + 1, r.copy_end) - host_rec_off)));
            if (uint64_t(rec_hi) + 1 > rec_copied) {
                KMN_CUDA(cudaMemcpyAsync(b.rec_off.p + rec_copied, host_rec_off + rec_copied, (uint64_t(rec_hi) + 1 - rec_copied) * sizeof(uint64_t),
                                         cudaMemcpyHostToDevice, c->copy_stream));
                rec_copied = uint64_t(rec_hi) + 1;
            }
            if (j == 0) KMN_CUDA(cudaEventRecord(c->ev_h2d[0], c->copy_stream));
            if (r.copy_end > r.copy_begin)
                KMN_CUDA(cudaMemcpyAsync(b.raw.p + r.copy_begin, host_residues + r.copy_begin, r.copy_end - r.copy_begin,
                                         cudaMemcpyHostToDevice, c->copy_stream));
            if (j + 1 == ranges.size()) KMN_CUDA(cudaEventRecord(c->ev_h2d[1], c->copy_stream));
            cudaEvent_t e = c->ev_copy[1 + j % (kMaxH2dChunks)];
            KMN_CUDA(cudaEventRecord(e, c->copy_stream));
            KMN_CUDA(cudaStreamWaitEvent(fs, e, 0));
            mark(c->copy_stream, "copy done", j);
            if (j == 0) {   // the order check of rec_off runs while the first copy is in flight, before any kernel reads an offset
                check_deferred_offsets(b);
                if (b.bad_offsets) {
                    KMN_CUDA(cudaStreamSynchronize(c->copy_stream));
                    throw std::runtime_error("rec_off must be non-decreasing");
                }
            }
        }
        mark(fs, "frame begins", j);
        // ---- K0 frame on the range's tiles (running total of kept bytes carried on the device)
        const uint32_t nt = r.tile_end - r.tile_begin;
        if (b.n_rec > 0 && nt > 0) {
            frame_tile_records_kernel<<<(nt + 1 + 255) / 256, 256, 0, fs>>>(b.rec_off.p, rec_hi, r.tile_begin, nt + 1, c->tile_rec.p,
                                                                            c->tile_blo.p);
            check_launch(c);
            frame_kernel<<<nt, FRAME_THREADS, 0, fs>>>(b.raw.p, b.n_raw, c->lut, b.rec_off.p, b.n_rec, b.codes.p, b.rec_cstart.p,
                                                    r.tile_begin, c->tile_rec.p, c->tile_blo.p, c->tile_state.p,
                                                    c->frame_ticket.p + j % (kMaxH2dChunks + 1));
            check_launch(c);
        }
        if (b.n_rec > 0 && j + 1 == ranges.size()) {   // rec_cstart[n_rec] = number of codes
            codes_tail_kernel<<<grid_for(c, 2), 256, 0, fs>>>(b.codes.p, b.rec_cstart.p + b.n_rec, b.codes_cap);
            check_launch(c);
        }
        {
            const uint32_t s_begin = j == 0 ? 0 : r.s0 + 1, s_end = r.s1;   // boundary s0 was written by the previous range
            if (s_end >= s_begin) {
                species_offsets_kernel<<<(s_end - s_begin + 1 + 255) / 256, 256, 0, fs>>>(b.rec_cstart.p, b.sp_rec_off.p,
                                                                                       s_begin, s_end, b.sp_cstart.p);
                check_launch(c);
            }
        }
        mark(fs, "frame done", j);
        if (fs != c->stream) {   // the Bloom / count-min kernels of this range follow its frame
            cudaEvent_t e = c->ev_frame[1 + j % kMaxH2dChunks];
            KMN_CUDA(cudaEventRecord(e, fs));
            KMN_CUDA(cudaStreamWaitEvent(c->stream, e, 0));
        }
        mark(c->stream, "bloom begins", j);
        if (timed) KMN_CUDA(cudaEventRecord(c->ev[1], c->stream));
        if (frame_only || r.s1 == r.s0) continue;
        // ---- K2 Bloom: probe partition, per-bucket shared-memory bitsets, slow exact path
        const uint32_t n_sp_r = r.s1 - r.s0;
        const uint32_t bl_bound = uint32_t(r.codes_bound / BP_TILE) + 2 * n_sp_r + 2;
        species_tiles_kernel<<<2, 1024, 0, c->stream>>>(b.sp_cstart.p, r.s0, r.s1,
                                                      TilesJob{BP_TILE_WT, c->bl_tile0.p, c->bl_tile_sp.p, c->n_tiles_dev.p, nullptr},
                                                      TilesJob{PART_WTILES, c->cms_tile0.p, c->cms_tile_sp.p, c->n_tiles_dev.p + 1,
                                                               c->cms_running.p});
        check_launch(c);
        {
            const uint32_t bp_grid = std::min<uint32_t>(bl_bound, 2u * uint32_t(c->sm_count));
            if (R.cap == RUN_CAP)
                bloom_partition_kernel<RUN_CAP><<<bp_grid, BP_THREADS, bp_smem_for(R.cap), c->stream>>>(b.codes.p, b.sp_cstart.p, c->k, c->k_mask, R);
            else   // test knob KMN_BLOOM_RUN_CAP
                bloom_partition_kernel<0><<<bp_grid, BP_THREADS, bp_smem_for(R.cap), c->stream>>>(b.codes.p, b.sp_cstart.p, c->k, c->k_mask, R);
        }
        check_launch(c);
        if (timed) KMN_CUDA(cudaEventRecord(c->ev[8], c->stream));
        // one warp per (species, bucket), persistent CTAs of BW_WARPS warps
        const uint32_t bb_items = n_sp_r * BB_BUCKETS;
        bloom_bucket_kernel<<<std::min<uint32_t>(uint32_t(c->sm_count), (bb_items + BW_WARPS - 1) / BW_WARPS), BW_THREADS, BW_SMEM, c->stream>>>(
            b.sp_cstart.p, r.s0, bb_items, R);
        check_launch(c);
        if (timed) KMN_CUDA(cudaEventRecord(c->ev[9], c->stream));
        bloom_slow_kernel<<<std::min<uint32_t>(n_sp_r, BS_MAX_CTAS), BS_THREADS, BS_SMEM, c->stream>>>(
            b.codes.p, b.sp_cstart.p, r.s0, r.s1, c->k, c->k_mask, R, c->bl_scratch.p);
        check_launch(c);
        if (timed) KMN_CUDA(cudaEventRecord(c->ev[2], c->stream));
        // ---- K3a count-min partition (also counts the passing occurrences per species)
        const uint32_t cms_bound = uint32_t(r.codes_bound / PART_TILE) + 2 * n_sp_r + 2;
        if (timed) KMN_CUDA(cudaEventRecord(c->ev[6], c->stream));
        cms_partition_kernel<<<std::min<uint32_t>(cms_bound, uint32_t(c->sm_count)), PART_THREADS, PART_SMEM, c->stream>>>(
            slots_map, b.codes.p, b.sp_cstart.p, T, c->bl_lost2.p, c->k, c->k_mask, L, c->cms_slot_limit, c->new_counts.p);
        check_launch(c);
        if (timed) KMN_CUDA(cudaEventRecord(c->ev[7], c->stream));
        mark(c->stream, "partition done", j);
    }
    const bool counted = !frame_only && b.n_sp > 0;
    if (counted) {
        // ---- K3b: per-bucket shared-memory histograms merged into the saturating table, once per batch
        if (c->cms_force == 2) {
            static const uint32_t one = 1;
            KMN_CUDA(cudaMemcpyAsync(L.state, &one, sizeof one, cudaMemcpyHostToDevice, c->stream));
        }
        // the first batch after kmn_reset_table writes every cell instead of adding to it
        cms_apply_kernel<<<CMS_BUCKETS, APPLY_THREADS, APPLY_SMEM, c->stream>>>(L, c->tab.p, c->cms_force == 1, c->tab_zero);
        check_launch(c);
        c->tab_zero = false;
        cms_spill_kernel<<<grid_for(c, 2), 256, 0, c->stream>>>(L, c->tab.p);
        check_launch(c);
        cms_direct_kernel<<<grid, 256, 0, c->stream>>>(b.codes.p, c->bl_lost2.p, uint32_t(b.codes_cap / WARP_TILE) - 1, c->k, c->k_mask,
                                                       c->tab.p, L.state);
        check_launch(c);
    }
    KMN_CUDA(cudaEventRecord(c->ev[3], c->stream));
    mark(c->stream, "apply done", 0);
    // host work that fits under the kernels just queued (before the device-to-host copies below: into pageable memory
    // they block until the stream has drained)
    for (auto& other : c->batches)
        if (other) check_deferred_offsets(*other);
    b.h_sp_cstart.resize(size_t(b.n_sp) + 1);
    KMN_CUDA(cudaMemcpyAsync(b.h_sp_cstart.data(), b.sp_cstart.p, (size_t(b.n_sp) + 1) * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, c->stream));
    std::vector<unsigned long long> h(b.n_sp);
    if (counted && new_kmers_per_species)
        KMN_CUDA(cudaMemcpyAsync(h.data(), c->new_counts.p, size_t(b.n_sp) * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    if (from_host && b.n_raw >= (uint64_t(8) << 20)) {   // link rate of this call: steers the next call's range schedule
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_h2d[0], c->ev_h2d[1]) == cudaSuccess) {
            if (ms > 0.f) c->h2d_gbs = float(b.n_raw) / (ms * 1e6f);
        } else {
            (void)cudaGetLastError();   // a failed measurement must not surface as the next launch's error
        }
    }
    if (trace && !marks.empty()) {
        for (const Mark& m : marks) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, marks[0].e, m.e) != cudaSuccess) (void)cudaGetLastError();
            std::fprintf(stderr, "[kmn trace] %8.3f ms  range %zu  %s\n", ms, m.j, m.what);
            if (&m != &marks[0]) cudaEventDestroy(m.e);
        }
        cudaEventDestroy(marks[0].e);
    }
    b.n_codes = b.h_sp_cstart[b.n_sp];
    b.framed = true;
    b.scanned = false;
    if (!frame_only && new_kmers_per_species)
        for (uint32_t s = 0; s < b.n_sp; s++) new_kmers_per_species[s] = counted ? h[s] : 0;
    c->stage_ms[0] = c->stage_ms[1] = c->stage_ms[2] = c->stage_ms[6] = c->stage_ms[7] = c->stage_ms[10] = c->stage_ms[11] = 0.f;
    if (timed && !frame_only) {
        KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[0], c->ev[0], c->ev[1]));
        if (counted) {
            KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[1], c->ev[1], c->ev[2]));
            KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[2], c->ev[2], c->ev[3]));
            KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[6], c->ev[6], c->ev[7]));
            KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[7], c->ev[8], c->ev[9]));
            KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[10], c->ev[1], c->ev[8]));
            KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[11], c->ev[7], c->ev[3]));
        }
    }
}

// A reset table that no batch has been counted into yet is zeroed here, on first use.
void materialize_table(kmn_ctx* c) {
    if (!c->tab_zero) return;
    KMN_CUDA(cudaMemsetAsync(c->tab.p, 0, CMS_CELLS * sizeof(uint16_t), c->stream));
    c->tab_zero = false;
}

void finalize_table(kmn_ctx* c) {
    // The table is the saturating u16 snapshot at all times (cms_kernels.cuh); only the NCCL variant,
    // which summed a widened copy across ranks, has something to clamp back.
    KMN_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (!c->widened) materialize_table(c);
    if (c->widened) {
        cms_narrow_kernel<<<grid_for(c, 8), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(c->wide.p),
                                                               reinterpret_cast<uint2*>(c->tab.p), CMS_CELLS / 4);
        check_launch(c);
        c->widened = false;
    }
    KMN_CUDA(cudaEventRecord(c->ev[1], c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[3], c->ev[0], c->ev[1]));
    c->final_tab = c->tab.p; c->tab8_src = c->tab_max_src = nullptr;
    c->finalized = true;
}

// 8 L2-sized sweeps over the u16 table: (row, half-row)
void wide_occupancy_sweeps(kmn_ctx* c, Batch& b, uint32_t n_tiles) {
    const int grid = grid_for(c, 8);
    for (uint32_t r = 0; r < CMS_DEPTH; r++)
        for (uint32_t half = 0; half < 2; half++) {
            occupancy_sweep_kernel<<<grid, SCAN_THREADS, 0, c->stream>>>(
                b.codes.p, n_tiles, c->k, c->k_mask, c->final_tab + (size_t(r) << CMS_WIDTH_LOG2), cms_seed(int(r)), half,
                r == 0 && half == 0, b.occ.p);
            check_launch(c);
        }
}

// Pass 2 may read the narrow (one byte per cell) copy of the finalized table when every cell fits the byte window
// above min_needed - 1 (scan_kernels.cuh).  Builds / reuses the copy; false: use the u16 table.
bool narrow_table_for(kmn_ctx* c, uint32_t min_needed) {
    if (c->scan_wide || min_needed < 1) return false;
    if (c->tab_max_src != c->final_tab) {
        c->tab_max_dev.ensure(1);
        KMN_CUDA(cudaMemsetAsync(c->tab_max_dev.p, 0, sizeof(uint32_t), c->stream));
        table_max_kernel<<<grid_for(c, 8), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(c->final_tab), CMS_CELLS / 8, c->tab_max_dev.p);
        check_launch(c);
        KMN_CUDA(cudaMemcpyAsync(&c->tab_max, c->tab_max_dev.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        c->tab_max_src = c->final_tab;
    }
    const uint32_t base = min_needed - 1;
    if (c->tab_max > base + 255u) return false;
    if (c->tab8_src != c->final_tab || c->tab8_base != base) {
        c->tab8.ensure(CMS_CELLS);
        narrow_table_kernel<<<grid_for(c, 8), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(c->final_tab),
                                                                  reinterpret_cast<uint2*>(c->tab8.p), CMS_CELLS / 8, base);
        check_launch(c);
        c->tab8_src = c->final_tab;
        c->tab8_base = base;
    }
    return true;
}

void scan_batch(kmn_ctx* c, Batch& b, uint32_t min_needed, uint64_t* n_starts, uint64_t* n_ends) {
    if (!c->finalized) throw std::runtime_error("kmn_scan_batch needs a finalized table (kmn_finalize_table / kmn_table_load)");
    if (!b.framed) run_pass1(c, b, nullptr, nullptr, nullptr, true, nullptr);
    KMN_CUDA(cudaEventRecord(c->ev[0], c->stream));
    const uint32_t n_tiles = uint32_t((size_t(b.n_codes) + WARP_TILE - 1) / WARP_TILE);
    b.occ.ensure(size_t(n_tiles + 1) * WARP_TILE);
    const int grid = grid_for(c, 8);
    b.occ_narrow = n_tiles && narrow_table_for(c, min_needed);
    if (b.occ_narrow) {   // 4 L2-sized sweeps over the narrow table: one per row
        for (uint32_t r = 0; r < CMS_DEPTH; r++) {
            occupancy_sweep8_kernel<<<grid, SCAN_THREADS, 0, c->stream>>>(
                b.codes.p, n_tiles, c->k, c->k_mask, c->tab8.p + (size_t(r) << CMS_WIDTH_LOG2), cms_seed(int(r)), r == 0,
                r + 1 == CMS_DEPTH, c->tab8_base, b.occ.p);
            check_launch(c);
        }
    } else if (n_tiles) {
        wide_occupancy_sweeps(c, b, n_tiles);
    }
    // reset positions -> segments
    const uint32_t n_ctiles = (b.n_codes + 4095) / 4096;
    c->tile_cnt.ensure(n_ctiles + 1);
    c->tile_base.ensure(n_ctiles + 2);
    uint32_t n_inv = 0;
    if (n_ctiles) {
        invalid_count_kernel<<<n_ctiles, 256, 0, c->stream>>>(b.codes.p, b.n_codes, c->tile_cnt.p);
        check_launch(c);
        const ScanJob job{c->tile_cnt.p, c->tile_base.p, n_ctiles, nullptr};
        exclusive_scan_kernel<<<1, 1024, 0, c->stream>>>(job, job);
        check_launch(c);
        KMN_CUDA(cudaMemcpyAsync(&n_inv, c->tile_base.p + n_ctiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        b.inv_pos.ensure(n_inv);
        invalid_scatter_kernel<<<n_ctiles, 256, 0, c->stream>>>(b.codes.p, b.n_codes, c->tile_base.p, b.inv_pos.p);
        check_launch(c);
    }
    b.n_inv = n_inv;
    c->seg_s_cnt.ensure(size_t(n_inv) + 1);
    c->seg_e_cnt.ensure(size_t(n_inv) + 1);
    b.seg_s_off.ensure(size_t(n_inv) + 2);
    b.seg_e_off.ensure(size_t(n_inv) + 2);
    c->sp_s_off.ensure(size_t(b.n_sp) + 1);
    c->sp_e_off.ensure(size_t(b.n_sp) + 1);
    uint32_t tot_s = 0, tot_e = 0;
    if (n_inv) {
        segment_boundary_kernel<false><<<grid, SCAN_THREADS, 0, c->stream>>>(
            b.occ.p, b.inv_pos.p, n_inv, c->k, min_needed, c->seg_s_cnt.p, c->seg_e_cnt.p, nullptr, nullptr, nullptr, nullptr);
        check_launch(c);
        device_exclusive_scan(c, c->seg_s_cnt.p, b.seg_s_off.p, n_inv);
        device_exclusive_scan(c, c->seg_e_cnt.p, b.seg_e_off.p, n_inv);
        KMN_CUDA(cudaMemcpyAsync(&tot_s, b.seg_s_off.p + n_inv, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaMemcpyAsync(&tot_e, b.seg_e_off.p + n_inv, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        b.starts.ensure(tot_s);
        b.ends.ensure(tot_e);
        segment_boundary_kernel<true><<<grid, SCAN_THREADS, 0, c->stream>>>(
            b.occ.p, b.inv_pos.p, n_inv, c->k, min_needed, nullptr, nullptr, b.seg_s_off.p, b.seg_e_off.p, b.starts.p, b.ends.p);
        check_launch(c);
        species_hit_offsets_kernel<<<(b.n_sp + 1 + 255) / 256, 256, 0, c->stream>>>(
            b.inv_pos.p, n_inv, b.sp_cstart.p, b.n_sp, b.seg_s_off.p, b.seg_e_off.p, c->sp_s_off.p, c->sp_e_off.p);
        check_launch(c);
        b.h_sp_starts_off.resize(size_t(b.n_sp) + 1);
        b.h_sp_ends_off.resize(size_t(b.n_sp) + 1);
        KMN_CUDA(cudaMemcpyAsync(b.h_sp_starts_off.data(), c->sp_s_off.p, (size_t(b.n_sp) + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaMemcpyAsync(b.h_sp_ends_off.data(), c->sp_e_off.p, (size_t(b.n_sp) + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    } else {
        b.h_sp_starts_off.assign(size_t(b.n_sp) + 1, 0);
        b.h_sp_ends_off.assign(size_t(b.n_sp) + 1, 0);
    }
    KMN_CUDA(cudaEventRecord(c->ev[1], c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[4], c->ev[0], c->ev[1]));
    b.scanned = true;
    b.paired = false;
    b.hits_expanded = false;
    b.tot_starts = tot_s;
    b.tot_ends = tot_e;
    if (n_starts) *n_starts = tot_s;
    if (n_ends) *n_ends = tot_e;
}

void pair_batch(kmn_ctx* c, Batch& b, uint32_t length_middle, uint64_t* n_pairs) {
    if (!b.scanned) throw std::runtime_error("batch has not been scanned (kmn_scan_batch)");
    KMN_CUDA(cudaEventRecord(c->ev[0], c->stream));
    const uint32_t n_seg = b.n_inv;
    uint32_t total = 0;
    b.h_sp_pairs_off.assign(size_t(b.n_sp) + 1, 0);
    if (n_seg) {
        c->seg_s_cnt.ensure(size_t(n_seg) + 1);   // scratch: pairs per segment, then their offsets in seg_e_cnt
        c->seg_e_cnt.ensure(size_t(n_seg) + 2);
        const uint32_t blocks = (n_seg + 255) / 256;
        // one pass: pairs at the segments' start-hit offsets of a scratch array (pairs <= start candidates), then compaction
        c->pair_scratch.ensure(std::max<uint32_t>(b.tot_starts, 1));
        pair_segments_kernel<<<blocks, 256, 0, c->stream>>>(b.starts.p, b.ends.p, b.seg_s_off.p, b.seg_e_off.p, n_seg,
                                                                      uint32_t(c->k), length_middle, c->seg_s_cnt.p,
                                                                      c->pair_scratch.p, b.codes.p, b.inv_pos.p, b.rec_cstart.p, b.n_rec,
                                                                      b.sp_cstart.p, b.sp_rec_off.p, b.n_sp);
        check_launch(c);
        device_exclusive_scan(c, c->seg_s_cnt.p, c->seg_e_cnt.p, n_seg);
        KMN_CUDA(cudaMemcpyAsync(&total, c->seg_e_cnt.p + n_seg, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        b.pairs.ensure(total);
        if (total) {
            pair_compact_kernel<<<grid_for(c, 8), 256, 0, c->stream>>>(c->pair_scratch.p, b.seg_s_off.p, c->seg_e_cnt.p, n_seg, b.pairs.p);
            check_launch(c);
        }
        species_hit_offsets_kernel<<<(b.n_sp + 1 + 255) / 256, 256, 0, c->stream>>>(
            b.inv_pos.p, n_seg, b.sp_cstart.p, b.n_sp, c->seg_e_cnt.p, c->seg_e_cnt.p, c->sp_s_off.p, c->sp_e_off.p);
        check_launch(c);
        KMN_CUDA(cudaMemcpyAsync(b.h_sp_pairs_off.data(), c->sp_s_off.p, (size_t(b.n_sp) + 1) * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, c->stream));
    }
    KMN_CUDA(cudaEventRecord(c->ev[1], c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[8], c->ev[0], c->ev[1]));
    b.paired = true;
    if (n_pairs) *n_pairs = total;
}

// scratch of the grouping, sized for n observations (lives in the context: cudaMalloc / cudaFree per call would
// dominate the few ms of kernels)
void group_scratch(kmn_ctx* c, uint32_t n) {
    const uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
    c->g_left.ensure(n); c->g_right.ensure(n); c->g_sid.ensure(n); c->g_len.ensure(n);
    c->g_perm[0].ensure(n); c->g_perm[1].ensure(n);
    c->g_hist.ensure(size_t(RS_BINS) * n_blocks + 1); c->g_base.ensure(size_t(RS_BINS) * n_blocks + 2);
    c->g_head.ensure(n); c->g_scan.ensure(size_t(n) + 1); c->g_first.ensure(size_t(n) + 1);
}

// Sort + run detection + length selection on the observation arrays already in g_left / g_right / g_sid / g_len.
void group_core(kmn_ctx* c, uint32_t n, uint32_t min_needed, uint32_t* perm, kmn_group* groups, uint64_t cap_groups,
                uint64_t* n_groups) {
    DevBuf<uint64_t>&d_left = c->g_left, &d_right = c->g_right;
    DevBuf<uint32_t>&d_sid = c->g_sid, &d_len = c->g_len, &d_hist = c->g_hist, &d_base = c->g_base, &d_head = c->g_head,
                    &d_scan = c->g_scan, &d_first = c->g_first;
    DevBuf<uint32_t>* d_perm = c->g_perm;
    DevBuf<kmn_group>& d_groups = c->g_groups;
    const uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
    KMN_CUDA(cudaEventRecord(c->ev[0], c->stream));
    const uint32_t eb = (n + 255) / 256;
    iota_kernel<<<eb, 256, 0, c->stream>>>(d_perm[0].p, n);
    check_launch(c);
    // stable LSD radix sort by (left, right): the digits of right first, then those of left
    const uint32_t bits = uint32_t(3 * c->k);   // packed k-mers use 3k bits (group_extraction.rs:545-549)
    int cur = 0;
    for (int word = 0; word < 2; word++) {
        const uint64_t* key = word == 0 ? d_right.p : d_left.p;
        for (uint32_t shift = 0; shift < bits; shift += RS_BITS) {
            radix_hist_kernel<<<n_blocks, RS_THREADS, 0, c->stream>>>(key, d_perm[cur].p, n, shift, d_hist.p);
            check_launch(c);
            device_exclusive_scan(c, d_hist.p, d_base.p, RS_BINS * n_blocks);
            radix_scatter_kernel<<<n_blocks, RS_THREADS, 0, c->stream>>>(key, d_perm[cur].p, n, shift, d_base.p, d_perm[cur ^ 1].p);
            check_launch(c);
            cur ^= 1;
        }
    }
    // runs of equal keys -> groups
    group_heads_kernel<<<eb, 256, 0, c->stream>>>(d_left.p, d_right.p, d_perm[cur].p, n, d_head.p);
    check_launch(c);
    device_exclusive_scan(c, d_head.p, d_scan.p, n);
    uint32_t ng = 0;
    KMN_CUDA(cudaMemcpyAsync(&ng, d_scan.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    if (ng > cap_groups) throw std::runtime_error("group buffer too small");
    group_firsts_kernel<<<eb, 256, 0, c->stream>>>(d_head.p, d_scan.p, n, d_first.p);
    check_launch(c);
    d_groups.ensure(ng);
    group_select_kernel<<<(ng + 127) / 128, 128, 0, c->stream>>>(d_perm[cur].p, d_sid.p, d_len.p, d_first.p, ng, n, min_needed,
                                                             uint32_t(2 * c->k + 1), d_groups.p);
    check_launch(c);
    KMN_CUDA(cudaEventRecord(c->ev[1], c->stream));
    KMN_CUDA(cudaMemcpyAsync(perm, d_perm[cur].p, size_t(n) * 4, cudaMemcpyDeviceToHost, c->stream));
    KMN_CUDA(cudaMemcpyAsync(groups, d_groups.p, size_t(ng) * sizeof(kmn_group), cudaMemcpyDeviceToHost, c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[9], c->ev[0], c->ev[1]));
    if (n_groups) *n_groups = ng;
}

void group_observations(kmn_ctx* c, uint64_t n64, const uint64_t* left, const uint64_t* right, const uint32_t* sid,
                        const uint32_t* len, uint32_t min_needed, uint32_t* perm, kmn_group* groups, uint64_t cap_groups,
                        uint64_t* n_groups) {
    if (n64 >= (1ull << 31)) throw std::runtime_error("too many observations for one grouping call");
    const uint32_t n = uint32_t(n64);
    if (n_groups) *n_groups = 0;
    if (n == 0) return;
    if (!left || !right || !sid || !len || !perm || !groups) throw std::runtime_error("NULL argument");
    group_scratch(c, n);
    KMN_CUDA(cudaMemcpyAsync(c->g_left.p, left, size_t(n) * 8, cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaMemcpyAsync(c->g_right.p, right, size_t(n) * 8, cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaMemcpyAsync(c->g_sid.p, sid, size_t(n) * 4, cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaMemcpyAsync(c->g_len.p, len, size_t(n) * 4, cudaMemcpyHostToDevice, c->stream));
    group_core(c, n, min_needed, perm, groups, cap_groups, n_groups);
}

// The same from the anchor pairs that kmn_pair_batch left on the device: nothing but a few offsets is uploaded.
void group_batches(kmn_ctx* c, const uint32_t* batch_ids, uint32_t n_batches, const uint32_t* sid_base, uint32_t sid_stride,
                   uint32_t min_needed, uint32_t* perm, uint64_t cap_perm, kmn_group* groups, uint64_t cap_groups,
                   uint64_t* n_obs, uint64_t* n_groups) {
    if (!batch_ids || !sid_base || !n_obs) throw std::runtime_error("NULL argument");
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_batches; i++) {
        Batch& b = get_batch(c, batch_ids[i]);
        if (!b.paired) throw std::runtime_error("batch has not been paired (kmn_pair_batch)");
        total += b.h_sp_pairs_off[b.n_sp];
    }
    *n_obs = total;
    if (n_groups) *n_groups = 0;
    if (!perm || total == 0) return;   // size query
    if (total >= (1ull << 31)) throw std::runtime_error("too many observations for one grouping call");
    if (!groups) throw std::runtime_error("NULL argument");
    if (cap_perm < total) throw std::runtime_error("perm buffer too small");
    const uint32_t n = uint32_t(total);
    group_scratch(c, n);
    uint32_t at = 0;
    for (uint32_t i = 0; i < n_batches; i++) {
        Batch& b = get_batch(c, batch_ids[i]);
        const uint32_t np = b.h_sp_pairs_off[b.n_sp];
        if (np == 0) continue;
        c->g_spoff.ensure(size_t(b.n_sp) + 1);
        // (pageable source: the copy is staged before the call returns, so reusing g_spoff for the next batch is safe)
        KMN_CUDA(cudaMemcpyAsync(c->g_spoff.p, b.h_sp_pairs_off.data(), (size_t(b.n_sp) + 1) * sizeof(uint32_t),
                                 cudaMemcpyHostToDevice, c->stream));
        obs_from_pairs_kernel<<<(np + 255) / 256, 256, 0, c->stream>>>(b.pairs.p, np, c->g_spoff.p, b.n_sp, sid_base[i], sid_stride, at,
                                                                     c->g_left.p, c->g_right.p, c->g_sid.p, c->g_len.p);
        check_launch(c);
        at += np;
    }
    group_core(c, n, min_needed, perm, groups, cap_groups, n_groups);
}

// rows [row0, row1) of the pair matrix (row0 a multiple of the 64-row tile; row1 a multiple or n): the
// row block one rank computes when the matrix is sharded; the whole matrix is rows [0, n).
void pair_counts(kmn_ctx* c, const uint8_t* mapped, uint32_t n, uint64_t L, uint32_t row0, uint32_t row1, uint32_t* valid,
                 uint32_t* mismatch) {
    if (row0 > row1 || row1 > n) throw std::runtime_error("pair_counts: invalid row range");
    if (row0 % PAIR_TILE != 0 || (row1 % PAIR_TILE != 0 && row1 != n))
        throw std::runtime_error("pair_counts: row range must be aligned to 64-row tiles");
    c->nj_rows = row1 - row0;
    c->nj_n = n;
    if (row0 == row1) return;
    if (L >= (1ull << 32)) throw std::runtime_error("alignment too long for u32 counts");
    const uint32_t n_tiles = (n + PAIR_TILE - 1) / PAIR_TILE;
    const uint32_t n_pad = n_tiles * PAIR_TILE;
    const uint32_t W = uint32_t((L + 31) / 32);
    const uint32_t t0 = row0 / PAIR_TILE, t1 = (row1 + PAIR_TILE - 1) / PAIR_TILE;
    const size_t out_cells = size_t(row1 - row0) * n;
    // scratch lives in the context; a row block only pairs rows >= row0 (columns j > i), so only those rows of the
    // alignment are uploaded and bit-sliced (a rank of a sharded --nj moves its share, not the whole alignment)
    DevBuf<uint8_t>& d_mapped = c->nj_mapped;
    DevBuf<uint32_t>& d_planes = c->nj_planes;
    DevBuf<uint32_t>&d_valid = c->nj_valid, &d_mism = c->nj_mism;
    d_mapped.ensure(std::max<size_t>(size_t(n) * L, 1));
    d_planes.ensure(size_t(n_pad) * std::max<uint32_t>(W, 1) * PAIR_PLANES);
    d_valid.ensure(out_cells);
    d_mism.ensure(out_cells);
    if (L) KMN_CUDA(cudaMemcpyAsync(d_mapped.p + size_t(row0) * L, mapped + size_t(row0) * L, size_t(n - row0) * L,
                                    cudaMemcpyHostToDevice, c->stream));
    KMN_CUDA(cudaMemsetAsync(d_valid.p, 0, out_cells * sizeof(uint32_t), c->stream));
    KMN_CUDA(cudaMemsetAsync(d_mism.p, 0, out_cells * sizeof(uint32_t), c->stream));
    KMN_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (W) {
        const uint64_t items = uint64_t(n_pad - row0) * W;   // rows row0 .. n_pad
        pair_pack_kernel<<<uint32_t((items + 255) / 256), 256, 0, c->stream>>>(d_mapped.p, n, L, n_pad, W, row0, d_planes.p);
        check_launch(c);
        // tile rows t0..t1-1 of the upper triangle: sum of (n_tiles - t)
        const uint64_t n_blocks = uint64_t(t1 - t0) * n_tiles - (uint64_t(t1) * (t1 - 1) / 2 - uint64_t(t0) * (t0 > 0 ? t0 - 1 : 0) / 2);
        pair_count_kernel<<<uint32_t(n_blocks), PAIR_THREADS, PAIR_COUNT_SMEM, c->stream>>>(d_planes.p, n, W, n_tiles, t0, d_valid.p,
                                                                                          d_mism.p);
        check_launch(c);
    }
    KMN_CUDA(cudaEventRecord(c->ev[1], c->stream));
    if (valid) KMN_CUDA(cudaMemcpyAsync(valid, d_valid.p, out_cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (mismatch) KMN_CUDA(cudaMemcpyAsync(mismatch, d_mism.p, out_cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[5], c->ev[0], c->ev[1]));
}

// One IPC-exported allocation per rank: the exchanged u16 table, the signal words of the in-kernel rank
// synchronisation, and the byte planes of the plane exchange (partial table: P_*, exchanged table: F_*).
constexpr size_t kXOffFlags = CMS_CELLS * sizeof(uint16_t);
constexpr size_t kXOffPLo = kXOffFlags + 256;
constexpr size_t kXOffPHi = kXOffPLo + CMS_CELLS;
constexpr size_t kXOffPBm = kXOffPHi + CMS_CELLS;
constexpr size_t kXOffFLo = kXOffPBm + PLANE_BM_WORDS * sizeof(uint32_t);
constexpr size_t kXOffFHi = kXOffFLo + CMS_CELLS;
constexpr size_t kXOffFBm = kXOffFHi + CMS_CELLS;
constexpr size_t kXBytes = kXOffFBm + PLANE_BM_WORDS * sizeof(uint32_t);
void alloc_exchange_buffers(kmn_ctx* c) {
    if (c->xtab.p) return;
    c->xtab.alloc(kXBytes / sizeof(uint16_t));
    KMN_CUDA(cudaMemset(reinterpret_cast<uint8_t*>(c->xtab.p) + kXOffFlags, 0, 256));
    c->xsync.alloc(4);
    KMN_CUDA(cudaMemset(c->xsync.p, 0, 4 * sizeof(uint32_t)));
    if (!c->xabort_host) {   // pinned source word of kmn_exchange_abort's copy
        KMN_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->xabort_host), sizeof(uint32_t), cudaHostAllocPortable));
        *c->xabort_host = 1;
    }
}
ExchangeGuard exchange_guard(kmn_ctx* c) {
    return ExchangeGuard{c->xsync.p + 2, c->xsync.p + 3, uint64_t(c->exchange_timeout_ms) * 1000000ull};
}
// after the exchange kernel has finished: turn a recorded timeout / abort into an error (and re-arm the word)
void check_exchange_error(kmn_ctx* c) {
    uint32_t err = 0;
    KMN_CUDA(cudaMemcpyAsync(&err, c->xsync.p + 2, sizeof err, cudaMemcpyDeviceToHost, c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    if (!err) return;
    KMN_CUDA(cudaMemsetAsync(c->xsync.p + 2, 0, sizeof(uint32_t), c->stream));
    KMN_CUDA(cudaStreamSynchronize(c->stream));
    const uint32_t who = (err & 0xFFu) - 1u;
    std::string msg = (err & 0x200u) ? "table exchange aborted by the host while waiting for rank "
                                     : "table exchange timed out after " + std::to_string(c->exchange_timeout_ms) + " ms waiting for rank ";
    msg += std::to_string(who) + ((err & 0x100u) ? " to finish" : " to arrive");
    msg += " (peer died or never launched?); call kmn_exchange_setup on every rank before exchanging again";
    throw std::runtime_error(msg);
}
void set_peer_planes(PeerPlanes& pp, int p, void* xbuf) {
    uint8_t* b = static_cast<uint8_t*>(xbuf);
    pp.p_lo[p] = reinterpret_cast<const uint4*>(b + kXOffPLo);
    pp.p_hi[p] = reinterpret_cast<const uint4*>(b + kXOffPHi);
    pp.p_bm[p] = reinterpret_cast<const uint32_t*>(b + kXOffPBm);
    pp.f_lo[p] = reinterpret_cast<uint4*>(b + kXOffFLo);
    pp.f_hi[p] = reinterpret_cast<uint4*>(b + kXOffFHi);
    pp.f_bm[p] = reinterpret_cast<uint32_t*>(b + kXOffFBm);
    pp.flags[p] = reinterpret_cast<uint32_t*>(b + kXOffFlags);
}

template <typename F>
int guarded(kmn_ctx* c, F&& f) {
    try {
        if (!c) throw std::runtime_error("NULL context");
        DeviceGuard dg(c->device);
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}

}  // namespace

extern "C" {

const char* kmn_last_error(void) { return g_err.c_str(); }
uint32_t kmn_abi_version(void) { return 7; }

int kmn_create(kmn_ctx** out, int device, int k, int scheme) {
    try {
        if (!out) throw std::runtime_error("out must not be NULL");
        *out = nullptr;
        if (k < 1 || k > 21) throw std::runtime_error("invalid k");  // lib.rs:218
        if (scheme < 0 || scheme > 2) throw std::runtime_error("invalid recode scheme");
        int n_dev = 0;
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        if (e != cudaSuccess || n_dev == 0)
            throw std::runtime_error(std::string("no CUDA device available (there is no CPU fallback): ") +
                                     cudaGetErrorString(e));
        if (device < 0 || device >= n_dev) throw std::runtime_error("invalid device ordinal");
        KMN_CUDA(cudaSetDevice(device));
        auto c = std::make_unique<kmn_ctx>();
        c->device = device;
        c->k = k;
        c->scheme = scheme;
        c->k_mask = (3 * k >= 64) ? ~0ull : ((1ull << (3 * k)) - 1);  // group_extraction.rs:545-549
        c->lut = make_lut(scheme);
        if (const char* e = std::getenv("KMN_CMS_FORCE")) {   // test knob: exercise the rare exact paths
            const long v = std::strtol(e, nullptr, 10);
            if (v < 0 || v > 2) throw std::runtime_error("KMN_CMS_FORCE must be 0, 1 or 2");
            c->cms_force = int(v);
        }
        if (const char* e = std::getenv("KMN_CMS_SLOT")) {    // test knob: small slots -> many spills
            const long v = std::strtol(e, nullptr, 10);
            if (v < 1 || v > PART_SLOT_CAP) throw std::runtime_error("KMN_CMS_SLOT must be in 1..31");
            c->cms_slot_limit = uint32_t(v);
        }
        if (const char* e = std::getenv("KMN_CMS_CAP")) {     // test knob: tiny spill list
            const long v = std::strtol(e, nullptr, 10);
            if (v < 1 || v > (1l << 30)) throw std::runtime_error("KMN_CMS_CAP must be in 1..2^30");
            c->cms_cap_override = uint32_t(v);
        }
        KMN_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
        KMN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        KMN_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        KMN_CUDA(cudaStreamCreateWithFlags(&c->frame_stream, cudaStreamNonBlocking));
        for (auto& ev : c->ev_frame) KMN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : c->ev) KMN_CUDA(cudaEventCreate(&ev));
        for (auto& ev : c->ev_copy) KMN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : c->ev_h2d) KMN_CUDA(cudaEventCreate(&ev));
        c->frame_ticket.alloc(kMaxH2dChunks + 1);
        c->n_tiles_dev.alloc(2);
        if (const char* e = std::getenv("KMN_EXCHANGE_PLANES")) c->exchange_planes = std::strtol(e, nullptr, 10) != 0;
        if (const char* e = std::getenv("KMN_EXCHANGE_TIMEOUT_MS")) {
            const long v = std::strtol(e, nullptr, 10);
            if (v < 1 || v > 3600000) throw std::runtime_error("KMN_EXCHANGE_TIMEOUT_MS must be in 1..3600000");
            c->exchange_timeout_ms = uint32_t(v);
        }
        if (const char* e = std::getenv("KMN_H2D_CHUNKS")) {   // tuning knob
            const long v = std::strtol(e, nullptr, 10);
            if (v < 1 || v > long(kMaxH2dChunks)) throw std::runtime_error("KMN_H2D_CHUNKS must be in 1.." + std::to_string(kMaxH2dChunks));
            c->h2d_chunks_forced = true;
            c->h2d_chunks = uint32_t(v);
        }
        if (const char* e = std::getenv("KMN_H2D_RATIO")) {    // tuning knob: growth of consecutive ranges
            const double v = std::strtod(e, nullptr);
            if (!(v >= 1.0 && v <= 4.0)) throw std::runtime_error("KMN_H2D_RATIO must be in 1..4");
            c->h2d_ratio = float(v);
        }
        if (const char* e = std::getenv("KMN_H2D_MODE")) {     // tuning knob: adaptive / doubling / equal
            const std::string m = e;
            if (m == "adaptive") c->h2d_mode = 0;
            else if (m == "doubling") c->h2d_mode = 1;
            else if (m == "equal") c->h2d_mode = 2;
            else throw std::runtime_error("KMN_H2D_MODE must be adaptive, doubling or equal");
        }
        c->scan_wide = std::getenv("KMN_SCAN_WIDE") != nullptr;
        if (const char* e = std::getenv("KMN_BLOOM_RUN_CAP")) {   // test knob
            const long v = std::strtol(e, nullptr, 10);
            if (v < 1 || v > long(RUN_CAP)) throw std::runtime_error("KMN_BLOOM_RUN_CAP must be in 1.." + std::to_string(RUN_CAP));
            c->bloom_run_cap = uint32_t(v);
        }
        c->bl_scratch.alloc(size_t(BS_MAX_CTAS) << (BLOOM_BITS_LOG2 - 5));
        KMN_CUDA(cudaFuncSetAttribute(pair_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PAIR_COUNT_SMEM)));
        KMN_CUDA(cudaFuncSetAttribute(bloom_partition_kernel<RUN_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bp_smem_for(RUN_CAP))));
        KMN_CUDA(cudaFuncSetAttribute(bloom_partition_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bp_smem_for(RUN_CAP))));
        KMN_CUDA(cudaFuncSetAttribute(bloom_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BW_SMEM)));
        KMN_CUDA(cudaFuncSetAttribute(bloom_slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BS_SMEM)));
        c->tab.alloc(CMS_CELLS);
        c->cms_running.alloc(2);
        c->cms_state.alloc(2);
        KMN_CUDA(cudaFuncSetAttribute(cms_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PART_SMEM)));
        KMN_CUDA(cudaFuncSetAttribute(cms_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(APPLY_SMEM)));
        KMN_CUDA(cudaMemsetAsync(c->tab.p, 0, CMS_CELLS * sizeof(uint16_t), c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        *out = c.release();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

void kmn_destroy(kmn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);   // staged copies nobody consumed
    c->batches.clear();
    c->spare.clear();
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->xabort_host) cudaFreeHost(c->xabort_host);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_copy) if (ev) cudaEventDestroy(ev);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto& ev : c->ev_frame) if (ev) cudaEventDestroy(ev);
    if (c->frame_stream) cudaStreamDestroy(c->frame_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int kmn_host_alloc(void** ptr, uint64_t bytes) {
    try {
        if (!ptr) throw std::runtime_error("ptr must not be NULL");
        *ptr = nullptr;
        KMN_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));   // pinned for every device of the process
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
void kmn_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

int kmn_stage_batch(kmn_ctx* c, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                    const uint64_t* sp_rec_off, uint32_t n_sp, uint32_t* batch_id) {
    return guarded(c, [&] { stage_batch(c, residues, rec_off, n_rec, sp_rec_off, n_sp, batch_id); });
}

int kmn_stage_batch_async(kmn_ctx* c, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                          const uint64_t* sp_rec_off, uint32_t n_sp, uint32_t* batch_id) {
    return guarded(c, [&] { stage_batch_async(c, residues, rec_off, n_rec, sp_rec_off, n_sp, batch_id); });
}

int kmn_count_staged(kmn_ctx* c, uint32_t batch_id, uint64_t* new_kmers_per_species) {
    return guarded(c, [&] { run_pass1(c, get_batch(c, batch_id), nullptr, nullptr, nullptr, false, new_kmers_per_species); });
}

int kmn_count_batch(kmn_ctx* c, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                    const uint64_t* sp_rec_off, uint32_t n_sp, uint64_t* new_kmers_per_species,
                    uint32_t* batch_id) {
    return guarded(c, [&] {
        c->batches.push_back(new_batch(c, residues, rec_off, n_rec, sp_rec_off, n_sp, /*defer_offset_check=*/true));
        const uint32_t id = uint32_t(c->batches.size() - 1);
        if (batch_id) *batch_id = id;
        // host -> device copies of the species ranges overlap the kernels of the previous range
        run_pass1(c, get_batch(c, id), residues, rec_off, sp_rec_off, false, new_kmers_per_species);
    });
}

int kmn_count_fasta(kmn_ctx* c, const uint8_t* bytes, const uint64_t* file_off, uint32_t n_files,
                    uint64_t* new_kmers_per_species, uint32_t* batch_id, uint64_t* n_records) {
    return guarded(c, [&] {
        if (c->finalized) throw std::runtime_error("table is finalized; call kmn_reset_table before counting again");
        c->batches.push_back(frame_fasta(c, bytes, file_off, n_files));
        const uint32_t id = uint32_t(c->batches.size() - 1);
        if (batch_id) *batch_id = id;
        if (n_records) *n_records = get_batch(c, id).n_rec;
        run_pass1(c, get_batch(c, id), nullptr, nullptr, nullptr, false, new_kmers_per_species);
    });
}

int kmn_fetch_records(kmn_ctx* c, uint32_t batch_id, uint64_t* hdr_off, uint64_t cap, uint64_t* n, uint64_t* sp_rec_off) {
    return guarded(c, [&] {
        Batch& b = get_batch(c, batch_id);
        if (!b.from_fasta) throw std::runtime_error("batch was not built by kmn_count_fasta");
        if (!n) throw std::runtime_error("n must not be NULL");
        *n = b.n_rec;
        if (sp_rec_off) std::copy(b.h_sp_rec_off.begin(), b.h_sp_rec_off.end(), sp_rec_off);
        if (!hdr_off) return;
        if (cap < b.n_rec) throw std::runtime_error("hdr_off capacity too small");
        if (b.n_rec) KMN_CUDA(cudaMemcpyAsync(hdr_off, b.hdr_off.p, size_t(b.n_rec) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int kmn_reset_table(kmn_ctx* c) {
    return guarded(c, [&] {
        c->tab_zero = true;   // zeroed lazily: by the first batch's apply pass, or by materialize_table
        c->finalized = false;
        c->widened = false;
        c->final_tab = nullptr; c->tab8_src = c->tab_max_src = nullptr;
    });
}

int kmn_release_batches(kmn_ctx* c) {
    return guarded(c, [&] {
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->copy_stream));   // staged copies nobody consumed
        for (auto& b : c->batches) {
            if (b) { b->staged_pending = false; b->pending_rec_off = nullptr; }
            if (b && c->spare.size() < 2) c->spare.push_back(std::move(b));
        }
        c->batches.clear();
    });
}

int kmn_release_batch(kmn_ctx* c, uint32_t batch_id) {
    return guarded(c, [&] {
        get_batch(c, batch_id);                                      // unknown ids are an error
        std::unique_ptr<Batch>& b = c->batches[batch_id];
        if (b->staged_pending) {                                     // never used: its copies may still read the caller's buffers
            KMN_CUDA(cudaEventSynchronize(b->staged_ev));
            b->staged_pending = false;
        }
        b->pending_rec_off = nullptr;
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        if (c->spare.size() < 2) c->spare.push_back(std::move(b));
        b.reset();                                                   // the id stays taken until kmn_release_batches
    });
}

int kmn_finalize_table(kmn_ctx* c) { return guarded(c, [&] { finalize_table(c); }); }

int kmn_table_snapshot(kmn_ctx* c, uint16_t* out) {
    return guarded(c, [&] {
        if (!out) throw std::runtime_error("out must not be NULL");
        if (!c->finalized) throw std::runtime_error("table not finalized");
        KMN_CUDA(cudaMemcpyAsync(out, c->final_tab, CMS_CELLS * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int kmn_table_load(kmn_ctx* c, const uint16_t* table) {
    return guarded(c, [&] {
        if (!table) throw std::runtime_error("table must not be NULL");
        KMN_CUDA(cudaMemcpyAsync(c->tab.p, table, CMS_CELLS * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        c->tab_zero = false;
        c->widened = false;
        c->final_tab = c->tab.p; c->tab8_src = c->tab_max_src = nullptr;
        c->finalized = true;
    });
}

int kmn_table_accum_device_ptr(kmn_ctx* c, void** ptr) {
    return guarded(c, [&] {
        if (!ptr) throw std::runtime_error("ptr must not be NULL");
        if (!c->wide.p) c->wide.alloc(CMS_CELLS);
        *ptr = c->wide.p;
    });
}
int kmn_table_widen(kmn_ctx* c) {
    return guarded(c, [&] {
        if (c->finalized) throw std::runtime_error("table is finalized");
        if (!c->wide.p) c->wide.alloc(CMS_CELLS);
        materialize_table(c);
        cms_widen_kernel<<<grid_for(c, 8), 256, 0, c->stream>>>(reinterpret_cast<const uint2*>(c->tab.p),
                                                              reinterpret_cast<uint4*>(c->wide.p), CMS_CELLS / 4);
        check_launch(c);
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        c->widened = true;
    });
}
int kmn_table_device_ptr(kmn_ctx* c, void** ptr) {
    return guarded(c, [&] {
        if (!ptr) throw std::runtime_error("ptr must not be NULL");
        if (c->tab_zero) {   // the lazy zeroing runs on the library's stream: finish it before another stream may touch the table
            materialize_table(c);
            KMN_CUDA(cudaStreamSynchronize(c->stream));
        }
        *ptr = const_cast<uint16_t*>(c->finalized ? c->final_tab : c->tab.p);
    });
}
int kmn_exchange_export(kmn_ctx* c, uint8_t* acc_handle64, uint8_t* table_handle64) {
    return guarded(c, [&] {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        alloc_exchange_buffers(c);
        cudaIpcMemHandle_t h;
        KMN_CUDA(cudaIpcGetMemHandle(&h, c->tab.p));
        std::memcpy(acc_handle64, &h, 64);
        KMN_CUDA(cudaIpcGetMemHandle(&h, c->xtab.p));
        std::memcpy(table_handle64, &h, 64);
    });
}

int kmn_exchange_setup(kmn_ctx* c, int world, int rank, const uint8_t* acc_handles, const uint8_t* table_handles) {
    return guarded(c, [&] {
        if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) throw std::runtime_error("invalid world / rank");
        if (CMS_CELLS / 8 % uint64_t(world)) throw std::runtime_error("world size must divide the table");
        alloc_exchange_buffers(c);
        for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
        c->ipc_opened.clear();
        c->peers = PeerTables{};
        c->peers.world = world;
        c->peers.rank = rank;
        c->planes = PeerPlanes{};
        c->planes.world = world;
        c->planes.rank = rank;
        for (int p = 0; p < world; p++) {
            if (p == rank) {
                c->peers.part[p] = reinterpret_cast<const uint4*>(c->tab.p);
                c->peers.full[p] = reinterpret_cast<uint4*>(c->xtab.p);
                c->peers.flags[p] = reinterpret_cast<uint32_t*>(c->xtab.p + CMS_CELLS);
                set_peer_planes(c->planes, p, c->xtab.p);
                continue;
            }
            cudaIpcMemHandle_t h;
            void* ptr = nullptr;
            std::memcpy(&h, acc_handles + 64 * p, 64);
            KMN_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(ptr);
            c->peers.part[p] = reinterpret_cast<const uint4*>(ptr);
            std::memcpy(&h, table_handles + 64 * p, 64);
            KMN_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(ptr);
            c->peers.full[p] = reinterpret_cast<uint4*>(ptr);
            c->peers.flags[p] = reinterpret_cast<uint32_t*>(static_cast<uint16_t*>(ptr) + CMS_CELLS);
            set_peer_planes(c->planes, p, ptr);
        }
        // a (re-)setup is collective: every rank starts again from epoch 0 with clean signal words
        c->exchange_epoch = 0;
        KMN_CUDA(cudaMemset(reinterpret_cast<uint8_t*>(c->xtab.p) + kXOffFlags, 0, 256));
        KMN_CUDA(cudaMemset(c->xsync.p, 0, 4 * sizeof(uint32_t)));
        KMN_CUDA(cudaDeviceSynchronize());   // cudaMemset runs on the legacy stream, which the context's non-blocking streams do not wait for
        c->peers_ready = true;
    });
}

// The exchange of one rank in two halves, so that one host thread can drive several GPUs (kmn_exchange_reduce_all):
// launch_exchange queues this rank's kernels (asynchronous: the kernel itself waits for the peers' kernels),
// finish_exchange waits for them and turns a recorded timeout into an error.
static void launch_exchange(kmn_ctx* c) {
    if (!c->peers_ready) throw std::runtime_error("kmn_exchange_setup has not been called");
    const int world = c->peers.world;
    const bool planes = c->exchange_planes && (world == 2 || world == 4 || world == 8);
    // the epoch only advances once this rank's kernel was launched: a failure before that must not leave
    // this rank one exchange ahead of its peers
    uint32_t epoch = c->exchange_epoch + 1;
    uint32_t* xs = c->xsync.p;
    ExchangeGuard guard = exchange_guard(c);
    KMN_CUDA(cudaMemsetAsync(c->xsync.p + 3, 0, sizeof(uint32_t), c->stream));   // re-arm the abort word
    materialize_table(c);
    KMN_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (planes) {
        // byte planes: split this rank's table, reduce + scatter planes over NVLink, rebuild the u16 table
        uint8_t* xb = reinterpret_cast<uint8_t*>(c->xtab.p);
        planes_split_kernel<<<grid_for(c, 8), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(c->tab.p),
                                                                 reinterpret_cast<uint4*>(xb + kXOffPLo),
                                                                 reinterpret_cast<uint4*>(xb + kXOffPHi),
                                                                 reinterpret_cast<uint32_t*>(xb + kXOffPBm));
        check_launch(c);
        uint64_t per = PLANE_BLOCKS / uint64_t(world), b0 = per * c->peers.rank, b1 = per * (c->peers.rank + 1);
        void* args[] = {&c->planes, &b0, &b1, &epoch, &xs, &guard};
        const void* fn = world == 2 ? reinterpret_cast<const void*>(exchange_planes_kernel<2>)
                       : world == 4 ? reinterpret_cast<const void*>(exchange_planes_kernel<4>)
                                    : reinterpret_cast<const void*>(exchange_planes_kernel<8>);
        if (c->exchange_grid == 0) {   // every CTA must be resident: the ranks synchronise inside the kernel
            int per_sm = 0;
            KMN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0));
            if (per_sm < 1) throw std::runtime_error("exchange_planes_kernel does not fit on an SM");
            c->exchange_grid = c->sm_count * std::min(per_sm, 8);
        }
        KMN_CUDA(cudaLaunchCooperativeKernel(fn, dim3(c->exchange_grid), dim3(256), args, 0, c->stream));
        check_launch(c);
        c->exchange_epoch = epoch;
        planes_join_kernel<<<grid_for(c, 8), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(xb + kXOffFLo),
                                                                reinterpret_cast<const uint4*>(xb + kXOffFHi),
                                                                reinterpret_cast<const uint32_t*>(xb + kXOffFBm),
                                                                reinterpret_cast<uint4*>(c->xtab.p));
    } else {
        const uint64_t n_vec = CMS_CELLS / 8, per = n_vec / uint64_t(world);   // uint4 = 8 cells
        uint64_t v0 = per * c->peers.rank, v1 = per * (c->peers.rank + 1);
        void* args[] = {&c->peers, &v0, &v1, &epoch, &xs, &guard};
        const void* fn;
        switch (world) {
            case 2: fn = reinterpret_cast<const void*>(exchange_reduce_kernel<2, 4>); break;
            case 4: fn = reinterpret_cast<const void*>(exchange_reduce_kernel<4, 2>); break;
            case 8: fn = reinterpret_cast<const void*>(exchange_reduce_kernel<8, 2>); break;
            default: fn = reinterpret_cast<const void*>(exchange_reduce_kernel<0, 1>); break;
        }
        if (c->exchange_grid == 0) {   // every CTA must be resident: the ranks synchronise inside the kernel
            int per_sm = 0;
            KMN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0));
            if (per_sm < 1) throw std::runtime_error("exchange_reduce_kernel does not fit on an SM");
            c->exchange_grid = c->sm_count * std::min(per_sm, 8);
        }
        KMN_CUDA(cudaLaunchCooperativeKernel(fn, dim3(c->exchange_grid), dim3(256), args, 0, c->stream));
        c->exchange_epoch = epoch;
    }
    check_launch(c);
    KMN_CUDA(cudaEventRecord(c->ev[1], c->stream));
}

static void finish_exchange(kmn_ctx* c) {
    check_exchange_error(c);   // synchronises the stream
    KMN_CUDA(cudaEventElapsedTime(&c->stage_ms[3], c->ev[0], c->ev[1]));
    c->final_tab = c->xtab.p; c->tab8_src = c->tab_max_src = nullptr;
    c->finalized = true;       // valid once every rank's kernel has finished
}

int kmn_exchange_reduce(kmn_ctx* c) {
    return guarded(c, [&] {
        launch_exchange(c);
        finish_exchange(c);
    });
}

// ---- in-process multi-GPU: one host process drives `world` contexts, one per device -------------------------
int kmn_exchange_setup_local(kmn_ctx* const* ctxs, int world) {
    try {
        if (!ctxs || world < 1 || world > MAX_PEERS) throw std::runtime_error("invalid context list");
        if (CMS_CELLS / 8 % uint64_t(world)) throw std::runtime_error("world size must divide the table");
        for (int r = 0; r < world; r++) {
            if (!ctxs[r]) throw std::runtime_error("NULL context");
            for (int q = 0; q < r; q++)
                if (ctxs[q]->device == ctxs[r]->device) throw std::runtime_error("every rank needs its own device");
        }
        for (int r = 0; r < world; r++) {   // peer access in both directions, exchange buffers
            DeviceGuard dg(ctxs[r]->device);
            for (int q = 0; q < world; q++) {
                if (q == r) continue;
                int can = 0;
                KMN_CUDA(cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device));
                if (!can) throw std::runtime_error("device " + std::to_string(ctxs[r]->device) + " cannot access device " +
                                                   std::to_string(ctxs[q]->device) + " (no peer path)");
                const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
                else KMN_CUDA(e);
            }
            alloc_exchange_buffers(ctxs[r]);
        }
        for (int r = 0; r < world; r++) {
            kmn_ctx* c = ctxs[r];
            DeviceGuard dg(c->device);
            for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
            c->ipc_opened.clear();
            c->peers = PeerTables{};
            c->peers.world = world;
            c->peers.rank = r;
            c->planes = PeerPlanes{};
            c->planes.world = world;
            c->planes.rank = r;
            for (int p = 0; p < world; p++) {   // same address space: the peers' buffers are used as they are
                c->peers.part[p] = reinterpret_cast<const uint4*>(ctxs[p]->tab.p);
                c->peers.full[p] = reinterpret_cast<uint4*>(ctxs[p]->xtab.p);
                c->peers.flags[p] = reinterpret_cast<uint32_t*>(ctxs[p]->xtab.p + CMS_CELLS);
                set_peer_planes(c->planes, p, ctxs[p]->xtab.p);
            }
            c->exchange_epoch = 0;
            KMN_CUDA(cudaMemset(reinterpret_cast<uint8_t*>(c->xtab.p) + kXOffFlags, 0, 256));
            KMN_CUDA(cudaMemset(c->xsync.p, 0, 4 * sizeof(uint32_t)));
            KMN_CUDA(cudaDeviceSynchronize());
            c->peers_ready = true;
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

int kmn_exchange_reduce_all(kmn_ctx* const* ctxs, int world) {
    try {
        if (!ctxs || world < 1) throw std::runtime_error("invalid context list");
        // queue every rank's kernels first (each waits inside for the others), then wait for all of them
        for (int r = 0; r < world; r++) {
            if (!ctxs[r]) throw std::runtime_error("NULL context");
            DeviceGuard dg(ctxs[r]->device);
            launch_exchange(ctxs[r]);
        }
        std::string first_err;
        for (int r = 0; r < world; r++) {
            try {
                DeviceGuard dg(ctxs[r]->device);
                finish_exchange(ctxs[r]);
            } catch (const std::exception& e) {
                if (first_err.empty()) first_err = "rank " + std::to_string(r) + ": " + e.what();
            }
        }
        if (!first_err.empty()) throw std::runtime_error(first_err);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

int kmn_exchange_abort(kmn_ctx* c) {
    // any host thread: a 4-byte copy on the copy stream lands while the exchange kernel spins on the compute stream
    return guarded(c, [&] {
        if (!c->xabort_host || !c->xsync.p) return;   // no exchange has been set up
        KMN_CUDA(cudaMemcpyAsync(c->xsync.p + 3, c->xabort_host, sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream));
        KMN_CUDA(cudaStreamSynchronize(c->copy_stream));
    });
}

int kmn_table_checksum(kmn_ctx* c, uint64_t* row_sums4, uint64_t* hash) {
    return guarded(c, [&] {
        if (!row_sums4 || !hash) throw std::runtime_error("NULL argument");
        const uint16_t* t = c->finalized ? c->final_tab : c->tab.p;
        if (!c->finalized) materialize_table(c);
        c->new_counts.ensure(8);
        KMN_CUDA(cudaMemsetAsync(c->new_counts.p, 0, 5 * sizeof(unsigned long long), c->stream));
        table_checksum_kernel<<<grid_for(c, 8), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(t), c->new_counts.p);
        check_launch(c);
        unsigned long long h[5];
        KMN_CUDA(cudaMemcpyAsync(h, c->new_counts.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < 4; r++) row_sums4[r] = h[r];
        *hash = h[4];
    });
}

int kmn_synchronize(kmn_ctx* c) {
    return guarded(c, [&] {
        KMN_CUDA(cudaStreamSynchronize(c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->copy_stream));   // copies of kmn_stage_batch_async nobody has consumed yet
        for (auto& b : c->batches)                         // ... and nothing of the caller's buffers is needed afterwards
            if (b) check_deferred_offsets(*b);
    });
}
int kmn_stream(kmn_ctx* c, void** stream) { return guarded(c, [&] { *stream = c->stream; }); }

int kmn_scan_batch(kmn_ctx* c, uint32_t batch_id, uint32_t min_needed, uint64_t* n_starts, uint64_t* n_ends) {
    return guarded(c, [&] { scan_batch(c, get_batch(c, batch_id), min_needed, n_starts, n_ends); });
}

int kmn_fetch_hits(kmn_ctx* c, uint32_t batch_id, uint32_t sp, kmn_hit* starts, uint64_t cap_starts,
                   uint64_t* n_starts, kmn_hit* ends, uint64_t cap_ends, uint64_t* n_ends) {
    return guarded(c, [&] {
        Batch& b = get_batch(c, batch_id);
        if (!b.scanned) throw std::runtime_error("batch has not been scanned (kmn_scan_batch)");
        if (sp >= b.n_sp) throw std::runtime_error("species index out of range");
        const uint64_t s0 = b.h_sp_starts_off[sp], s1 = b.h_sp_starts_off[sp + 1];
        const uint64_t e0 = b.h_sp_ends_off[sp], e1 = b.h_sp_ends_off[sp + 1];
        if (n_starts) *n_starts = s1 - s0;
        if (n_ends) *n_ends = e1 - e0;
        if (!starts && !ends) return;
        if ((starts && cap_starts < s1 - s0) || (ends && cap_ends < e1 - e0)) throw std::runtime_error("hit buffers too small");
        if (!b.hits_expanded) {   // the scan keeps compact hits; the full form is built once, when somebody asks for it
            b.full_starts.ensure(b.tot_starts);
            b.full_ends.ensure(b.tot_ends);
            if (b.n_inv) {
                const int g = grid_for(c, 8);
                expand_hits_kernel<<<g, SCAN_THREADS, 0, c->stream>>>(b.codes.p, b.inv_pos.p, b.n_inv, c->k, b.starts.p, b.seg_s_off.p,
                                                                     b.rec_cstart.p, b.n_rec, b.sp_cstart.p, b.sp_rec_off.p, b.n_sp,
                                                                     b.full_starts.p);
                check_launch(c);
                expand_hits_kernel<<<g, SCAN_THREADS, 0, c->stream>>>(b.codes.p, b.inv_pos.p, b.n_inv, c->k, b.ends.p, b.seg_e_off.p,
                                                                     b.rec_cstart.p, b.n_rec, b.sp_cstart.p, b.sp_rec_off.p, b.n_sp,
                                                                     b.full_ends.p);
                check_launch(c);
            }
            b.hits_expanded = true;
        }
        if (starts && s1 > s0)
            KMN_CUDA(cudaMemcpyAsync(starts, b.full_starts.p + s0, (s1 - s0) * sizeof(kmn_hit), cudaMemcpyDeviceToHost, c->stream));
        if (ends && e1 > e0)
            KMN_CUDA(cudaMemcpyAsync(ends, b.full_ends.p + e0, (e1 - e0) * sizeof(kmn_hit), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int kmn_pair_batch(kmn_ctx* c, uint32_t batch_id, uint32_t length_middle, uint64_t* n_pairs) {
    return guarded(c, [&] { pair_batch(c, get_batch(c, batch_id), length_middle, n_pairs); });
}

int kmn_fetch_pairs(kmn_ctx* c, uint32_t batch_id, uint32_t sp, kmn_pair* out, uint64_t cap, uint64_t* n) {
    return guarded(c, [&] {
        Batch& b = get_batch(c, batch_id);
        if (!b.paired) throw std::runtime_error("batch has not been paired (kmn_pair_batch)");
        if (sp >= b.n_sp) throw std::runtime_error("species index out of range");
        const uint64_t p0 = b.h_sp_pairs_off[sp], p1 = b.h_sp_pairs_off[sp + 1];
        if (n) *n = p1 - p0;
        if (!out || p1 == p0) return;
        if (cap < p1 - p0) throw std::runtime_error("pair buffer too small");
        KMN_CUDA(cudaMemcpyAsync(out, b.pairs.p + p0, (p1 - p0) * sizeof(kmn_pair), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int kmn_group_observations(kmn_ctx* c, uint64_t n, const uint64_t* left, const uint64_t* right, const uint32_t* sid,
                           const uint32_t* len, uint32_t min_needed, uint32_t* perm, kmn_group* groups,
                           uint64_t cap_groups, uint64_t* n_groups) {
    return guarded(c, [&] { group_observations(c, n, left, right, sid, len, min_needed, perm, groups, cap_groups, n_groups); });
}

int kmn_group_batches(kmn_ctx* c, const uint32_t* batch_ids, uint32_t n_batches, const uint32_t* sid_base, uint32_t sid_stride,
                      uint32_t min_needed, uint32_t* perm, uint64_t cap_perm, kmn_group* groups, uint64_t cap_groups,
                      uint64_t* n_obs, uint64_t* n_groups) {
    return guarded(c, [&] {
        group_batches(c, batch_ids, n_batches, sid_base, sid_stride, min_needed, perm, cap_perm, groups, cap_groups, n_obs, n_groups);
    });
}

int kmn_fetch_occupancy(kmn_ctx* c, uint32_t batch_id, uint32_t sp, uint16_t* occ, uint64_t cap, uint64_t* n) {
    return guarded(c, [&] {
        Batch& b = get_batch(c, batch_id);
        if (!b.scanned) throw std::runtime_error("batch has not been scanned (kmn_scan_batch)");
        if (sp >= b.n_sp) throw std::runtime_error("species index out of range");
        const uint32_t r0 = uint32_t(b.h_sp_rec_off[sp]), r1 = uint32_t(b.h_sp_rec_off[sp + 1]);
        const uint64_t n_res = uint64_t(b.h_sp_cstart[sp + 1] - b.h_sp_cstart[sp]) - (r1 - r0);
        if (n) *n = n_res;
        if (!occ || n_res == 0) return;
        if (cap < n_res) throw std::runtime_error("occupancy buffer too small");
        if (b.occ_narrow) {   // occ[] is exact only where >= min_needed: the estimates below it need the u16 table
            wide_occupancy_sweeps(c, b, uint32_t((size_t(b.n_codes) + WARP_TILE - 1) / WARP_TILE));
            b.occ_narrow = false;
        }
        DevBuf<uint16_t> tmp;
        tmp.alloc(n_res);
        occupancy_gather_kernel<<<std::min<uint32_t>(r1 - r0, 4096u), 128, 0, c->stream>>>(b.occ.p, b.rec_cstart.p, r0, r1, tmp.p);
        check_launch(c);
        KMN_CUDA(cudaMemcpyAsync(occ, tmp.p, n_res * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
        KMN_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int kmn_pair_counts(kmn_ctx* c, const uint8_t* mapped, uint32_t n, uint64_t L, uint32_t* valid, uint32_t* mismatch) {
    return guarded(c, [&] {
        if (!mapped || !valid || !mismatch) throw std::runtime_error("NULL argument");
        if (n < 2) return;
        pair_counts(c, mapped, n, L, 0, n, valid, mismatch);
    });
}

int kmn_pair_counts_rows(kmn_ctx* c, const uint8_t* mapped, uint32_t n, uint64_t L, uint32_t row_begin, uint32_t row_end,
                         uint32_t* valid, uint32_t* mismatch) {
    return guarded(c, [&] {
        if (!mapped) throw std::runtime_error("NULL argument");
        pair_counts(c, mapped, n, L, row_begin, row_end, valid, mismatch);
    });
}

int kmn_pair_counts_device(kmn_ctx* c, const uint32_t** valid, const uint32_t** mismatch, uint32_t* rows, uint32_t* n) {
    return guarded(c, [&] {
        if (!valid || !mismatch || !rows || !n) throw std::runtime_error("NULL argument");
        *valid = c->nj_valid.p;
        *mismatch = c->nj_mism.p;
        *rows = c->nj_rows;
        *n = c->nj_n;
    });
}

int kmn_last_stage_ms(kmn_ctx* c, int which, float* ms) {
    return guarded(c, [&] {
        if (which < 0 || which > 12 || !ms) throw std::runtime_error("invalid stage index");
        *ms = which == 12 ? c->h2d_gbs : c->stage_ms[which];
    });
}

uint64_t kmn_launch_count(kmn_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
