/* kamino_b200.h — C ABI of the B200-native kamino front end (libkamino_b200.so).
 *
 * Drop-in boundary for kamino's data-parallel front end.  The reference (Rust) has no FFI
 * layer; the seams this ABI replaces are crate-private functions, cited per entry point as
 * file:line under /root/reference.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success and non-zero on failure; kmn_last_error() returns a
 *     thread-local, NUL-terminated description of the last failure on the calling thread;
 *   - there is NO CPU fallback: a missing device, a CUDA error or an out-of-memory condition is
 *     an error return, never a silent host path;
 *   - one context drives one GPU (one process per GPU); calls on one context must be serialised
 *     by the caller; buffers passed in stay caller-owned and may be freed after the call returns;
 *   - "residues" are the bytes seq_io's rec.seq() yields for each FASTA record, concatenated:
 *     line terminators and other ASCII whitespace {0x20,0x09,0x0A,0x0C,0x0D} may be present and
 *     are skipped ON THE DEVICE exactly like group_extraction.rs:376-378; any other byte that is
 *     not one of the 20 amino-acid letters (either case) of the chosen scheme resets the roll
 *     (group_extraction.rs:380-384).
 */
#ifndef KAMINO_B200_H
#define KAMINO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kmn_ctx kmn_ctx;

/* RecodeScheme (src/recode.rs:9-16) */
enum { KMN_DAYHOFF6 = 0, KMN_SR6 = 1, KMN_KGB6 = 2 };

/* proba_filter.rs:13-19 — fixed by the reference; exported so hosts can size buffers. */
#define KMN_BLOOM_BITS  (1u << 25)
#define KMN_CMS_WIDTH   (1u << 26)
#define KMN_CMS_DEPTH   4u
#define KMN_CMS_CELLS   ((uint64_t)KMN_CMS_WIDTH * KMN_CMS_DEPTH)

/* One boundary hit of pass 2 (KmerHit, group_extraction.rs:131-139; end = start + k and
 * shared == (occ >= min_needed) are implied). */
typedef struct kmn_hit {
    uint64_t kmer;       /* packed 3-bit k-mer code                                              */
    uint32_t rec;        /* record (protein) index inside its species                            */
    uint32_t seg_start;  /* whitespace-stripped offset, inside the record, of the valid segment   */
    uint32_t start;      /* k-mer start inside that segment (KmerHit.start)                      */
    uint32_t occ;        /* CountMinSketch::estimate of the k-mer                                */
} kmn_hit;

/* One anchor pair of pass 2 ("bubble", group_extraction.rs:178-242): the amino-acid block is
 * aa[seg_start + left_start .. seg_start + left_start + len) of the whitespace-stripped, upper-cased record. */
typedef struct kmn_pair {
    uint64_t left;       /* packed k-mer of the selected start anchor                                   */
    uint64_t right;      /* packed k-mer of its best end anchor                                         */
    uint32_t rec;        /* record (protein) index inside its species                                   */
    uint32_t seg_start;  /* whitespace-stripped offset, inside the record, of the valid segment          */
    uint32_t left_start; /* start of the left anchor inside that segment                                */
    uint32_t len;        /* block length: right.start + k - left.start (both anchors included)          */
} kmn_pair;

/* One run of observations sharing an anchor pair after kmn_group_observations. */
typedef struct kmn_group {
    uint32_t first;      /* position of the run in the sorted order (index into perm)                   */
    uint32_t count;      /* observations in the run                                                     */
    uint32_t best_len;   /* block length with the largest support (distinct species)                    */
    uint32_t coverage;   /* that support                                                                */
    uint32_t flags;      /* KMN_GROUP_VALID: unique best length, coverage >= min_needed, length >= 2k+1;
                            KMN_GROUP_HOST: not decided on the device (more than 12 distinct lengths or
                            observations not species-major): the caller applies group_sorting.rs:90-143 */
} kmn_group;
#define KMN_GROUP_VALID 1u
#define KMN_GROUP_HOST  2u

const char* kmn_last_error(void);
/* ABI version of this header; bumped on any signature change. */
uint32_t kmn_abi_version(void);   /* currently 7 */

/* Context: owns the device, its stream, the count-min table (AtomicCountMinSketch::new,
 * proba_filter.rs:129-143), the scratch of the Bloom / count-min partition passes and all staged batches.
 * k in 1..=21 (lib.rs:208-218); scheme = KMN_*.  device = CUDA ordinal. */
int  kmn_create(kmn_ctx** out, int device, int k, int scheme);
void kmn_destroy(kmn_ctx* ctx);

/* Page-locked host memory for the batch buffers handed to kmn_count_batch / kmn_stage_batch (io.rs side of
 * the pipeline: FASTA bytes are gunzipped straight into it).  Optional — any host pointer works — but only
 * pinned buffers let the host-to-device copies of kmn_count_batch overlap its kernels at full link speed
 * (that holds for rec_off as well: it travels in slices with the residues, 8 bytes per record). */
int  kmn_host_alloc(void** ptr, uint64_t bytes);
void kmn_host_free(void* ptr);

/* ---- pass 1: table build ------------------------------------------------------------------
 * Replaces count_species_kmers (group_extraction.rs:359-395) driven by extract_groups stage 2
 * (group_extraction.rs:557-567) for a batch of WHOLE species.
 *   residues      concatenated rec.seq() bytes of all records of the batch (host memory)
 *   rec_off       [n_rec+1] byte offsets of the records inside `residues`, non-decreasing
 *   sp_rec_off    [n_sp+1]  record ranges of the species (file order inside a species)
 *   new_kmers_per_species (optional, [n_sp]) number of k-mers that passed the per-species Bloom
 *                 filter, i.e. the number of AtomicCountMinSketch::increment calls for it.
 * The batch stays resident in HBM (recoded, whitespace-free) for pass 2; *batch_id (optional)
 * receives its handle.  n_residues < 2^31 per batch.
 */
int kmn_count_batch(kmn_ctx* ctx, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                    const uint64_t* sp_rec_off, uint32_t n_sp, uint64_t* new_kmers_per_species,
                    uint32_t* batch_id);
/* The same in two steps: host->device staging only, then the kernels on the staged batch.
 * kmn_count_staged may be called repeatedly on one staged batch (benchmarks: inputs resident in
 * HBM); each call adds the batch to the table again, so pair it with kmn_reset_table. */
int kmn_stage_batch(kmn_ctx* ctx, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                    const uint64_t* sp_rec_off, uint32_t n_sp, uint32_t* batch_id);
/* kmn_stage_batch without the wait: the host-to-device copies are queued on the context's copy stream and the call
 * returns; the first call that uses the batch (kmn_count_staged, kmn_scan_batch) makes the device wait for them.  In a
 * job of several batches the copies of batch i+1 then overlap the kernels of batch i (the reference overlaps reading
 * the next files with counting the current ones the same way: rayon over species, group_extraction.rs:558-567):
 *     stage_async(b0); for i: { stage_async(b[i+1]); count_staged(b[i]); release_batch(b[i]); }
 * The host buffers must stay valid and unchanged until a call that uses the batch has returned (or kmn_release_batch /
 * kmn_release_batches / kmn_synchronize after it); pinned buffers (kmn_host_alloc) are needed for the overlap.
 * The order check of rec_off (one pass over 8 bytes per record) is deferred to the next count call of the context,
 * where it runs under that call's kernels, or to the batch's own first use: an unsorted rec_off makes the call that
 * first uses the batch fail ("rec_off must be non-decreasing"); all other argument errors are reported here. */
int kmn_stage_batch_async(kmn_ctx* ctx, const uint8_t* residues, const uint64_t* rec_off, uint64_t n_rec,
                          const uint64_t* sp_rec_off, uint32_t n_sp, uint32_t* batch_id);
int kmn_count_staged(kmn_ctx* ctx, uint32_t batch_id, uint64_t* new_kmers_per_species);

/* ---- FASTA framing on the device ("next" row 3 of SURVEY 8f) --------------------------------
 * kmn_count_fasta = the record splitting of seq_io::fasta::Reader as count_species_kmers drives
 * it (group_extraction.rs:368-375) + kmn_count_batch: the caller hands over whole (decompressed)
 * FASTA files and the device finds the records.
 *   bytes     n_files files back to back, file f = bytes[file_off[f] .. file_off[f+1]),
 *             file_off[0] = 0; one file = one species, in species order.  Every non-empty file
 *             must end with '\n' (append one to a file that lacks it: that only adds whitespace to
 *             its last sequence).  Pinned memory (kmn_host_alloc) makes the copy faster.
 * A record starts at every line whose first byte is '>' (lines end with '\n'); its sequence is
 * everything up to the next such line; only empty lines ("\n", "\r\n") may precede the first
 * record of a file, otherwise the call fails with "FASTA parse error in file <f>: expected '>' at
 * record start" and counts nothing.  Record indices (kmn_hit.rec, kmn_pair.rec) count the '>'
 * lines of a file in order, like the reference's protein ids.
 *   n_records (optional out): records found in the whole batch. */
int kmn_count_fasta(kmn_ctx* ctx, const uint8_t* bytes, const uint64_t* file_off, uint32_t n_files,
                    uint64_t* new_kmers_per_species, uint32_t* batch_id, uint64_t* n_records);
/* Record table of a kmn_count_fasta batch: hdr_off[r] = offset in `bytes` of the '>' of record r
 * (the header runs to the next '\n', the sequence from there to hdr_off[r+1] or to the end of the
 * file), sp_rec_off[f] (optional, n_files + 1 values) = first record of file f.  *n = records;
 * hdr_off may be NULL to query the count. */
int kmn_fetch_records(kmn_ctx* ctx, uint32_t batch_id, uint64_t* hdr_off, uint64_t cap, uint64_t* n,
                      uint64_t* sp_rec_off);

/* Zero the table (a fresh AtomicCountMinSketch). */
int kmn_reset_table(kmn_ctx* ctx);
/* Drop every staged batch (frees their HBM). */
int kmn_release_batches(kmn_ctx* ctx);
/* Release ONE staged batch (its buffers are recycled by the next stage / count call); its id is not reused. */
int kmn_release_batch(kmn_ctx* ctx, uint32_t batch_id);

/* AtomicCountMinSketch::snapshot (proba_filter.rs:171-177).  The device table is kept as the
 * reference's saturating u16 table at all times (every update is +1, so adding a batch's exact
 * per-cell count with a clamp at 65535 equals the CAS loop of proba_filter.rs:151-166): on one GPU
 * this only marks the table read-only until kmn_reset_table; after kmn_table_widen + an external sum
 * it clamps the summed u32 cells back into the table. */
int kmn_finalize_table(kmn_ctx* ctx);
/* Copy the finalized table to host: 4 * 2^26 u16, row-major == CountMinSketch.table. */
int kmn_table_snapshot(kmn_ctx* ctx, uint16_t* out);
/* Load a finalized table from host (pass 2 against an externally built table). */
int kmn_table_load(kmn_ctx* ctx, const uint16_t* table);
/* Multi-GPU plumbing, NCCL variant.  Partial tables of several GPUs combine exactly as
 * min(sum over ranks, 65535) (a clamped partial only ever replaces a value >= 65535).
 * kmn_table_widen copies this rank's u16 cells into a u32 buffer (KMN_CMS_CELLS entries) whose device
 * pointer kmn_table_accum_device_ptr returns (stable for the context's life); the caller sums that
 * buffer across ranks (ncclSum on u32) and then calls kmn_finalize_table.
 * kmn_table_device_ptr: the u16 table (the finalized one once finalized / exchanged). */
int kmn_table_accum_device_ptr(kmn_ctx* ctx, void** ptr);
int kmn_table_widen(kmn_ctx* ctx);
int kmn_table_device_ptr(kmn_ctx* ctx, void** ptr);
/* Fused NVLink exchange (one process per GPU): every rank exports CUDA IPC handles of its own table
 * and of its exchanged table (64 bytes each); after an all-gather of the handles by the caller,
 * kmn_exchange_setup maps the peers' buffers.  kmn_exchange_reduce then runs ONE kernel per rank that
 * loads this rank's 1/world cell range from every peer over NVLink (as byte planes: low bytes always, high
 * bytes only for the 16-cell blocks that have any), adds with saturation
 * (== snapshot of the summed sketch, proba_filter.rs:171-177) and stores the result into EVERY peer's
 * exchanged table: reduce-scatter + snapshot + all-gather without NCCL on the data path.  The ranks
 * synchronise INSIDE the kernel through signal words in peer memory (cooperative launch): it starts moving
 * data once every rank's counting kernels have finished and returns once every peer is done with this
 * rank's buffers, so no host barrier is needed; every rank must call it the same number of times. */
int kmn_exchange_export(kmn_ctx* ctx, uint8_t* acc_handle64, uint8_t* table_handle64);
int kmn_exchange_setup(kmn_ctx* ctx, int world, int rank, const uint8_t* acc_handles, const uint8_t* table_handles);
int kmn_exchange_reduce(kmn_ctx* ctx);
/* The same exchange when ONE host process drives several GPUs (the drop-in for the reference's -t N species
 * parallelism, group_extraction.rs:550-567 / :589-608: species shard over the devices, one exchange per job):
 * ctxs[r] is the context of rank r, each created on its own device.  kmn_exchange_setup_local enables peer
 * access between the devices and wires the ranks' buffers directly (no IPC handles); kmn_exchange_reduce_all
 * queues every rank's exchange kernel and then waits for all of them (same kernels, same in-kernel rank
 * synchronisation as kmn_exchange_reduce).  Calls on different contexts may otherwise run on different host
 * threads at the same time (pass 1 and pass 2 of each device's own species). */
int kmn_exchange_setup_local(kmn_ctx* const* ctxs, int world);
int kmn_exchange_reduce_all(kmn_ctx* const* ctxs, int world);
/* Every in-kernel wait of the exchange is bounded (KMN_EXCHANGE_TIMEOUT_MS, default 30 000): when a peer does
 * not arrive, kmn_exchange_reduce returns an error that names the rank instead of hanging the GPU; the ranks then
 * re-run kmn_exchange_setup before exchanging again.  kmn_exchange_abort (callable from ANY host thread while
 * kmn_exchange_reduce is blocked) makes the waiting kernel give up at once. */
int kmn_exchange_abort(kmn_ctx* ctx);
/* Row sums (4 x u64) and a position-sensitive 64-bit hash of the table (the finalized one once finalized):
 * 40 bytes that let ranks, runs and implementations compare 512 MiB tables.  While no cell saturates,
 * row_sums4[r] == the number of AtomicCountMinSketch::increment calls (proba_filter.rs:146-168), i.e. the
 * sum of new_kmers_per_species over everything counted.  hash = sum over cells of cell * mix64(index + 0x51),
 * index = row * 2^26 + column, arithmetic modulo 2^64 (mix64 = proba_filter.rs:30-37). */
int kmn_table_checksum(kmn_ctx* ctx, uint64_t* row_sums4, uint64_t* hash);
/* Block until all work queued on the context's streams has finished (staged copies included; the caller's buffers
 * of kmn_stage_batch_async calls are not touched afterwards). */
int kmn_synchronize(kmn_ctx* ctx);
/* CUDA stream of the context (cudaStream_t as void*), for event timing by the caller. */
int kmn_stream(kmn_ctx* ctx, void** stream);

/* ---- pass 2: occupancy scan ---------------------------------------------------------------
 * Replaces CountMinSketch::estimate (proba_filter.rs:103-112) per k-mer position plus the
 * segmentation and boundary rule of extract_species_blocks / extract_from_valid_segment
 * (group_extraction.rs:446-498, :264-303) for every species of a staged batch.  Results stay on
 * the device until fetched.  min_needed = ceil(f32(min_freq) * f32(n)) (group_extraction.rs:544).
 */
int kmn_scan_batch(kmn_ctx* ctx, uint32_t batch_id, uint32_t min_needed, uint64_t* n_starts,
                   uint64_t* n_ends);
/* Start candidates / end anchors of species `sp` (index inside the batch), in reference order
 * (record, segment, position).  Two-call protocol: with starts == NULL only the counts are set. */
int kmn_fetch_hits(kmn_ctx* ctx, uint32_t batch_id, uint32_t sp, kmn_hit* starts, uint64_t cap_starts,
                   uint64_t* n_starts, kmn_hit* ends, uint64_t cap_ends, uint64_t* n_ends);
/* Anchor selection + bubble pairing on the device for every segment of a scanned batch: replaces
 * select_best_starts_per_local_cluster (group_extraction.rs:151-175) and the anchor search of
 * extract_bubble_pairs (group_extraction.rs:178-242); the caller only slices the amino-acid blocks.
 * kmn_fetch_pairs: pairs of species `sp` in reference order (record, segment, left anchor); with
 * out == NULL only *n is set. */
int kmn_pair_batch(kmn_ctx* ctx, uint32_t batch_id, uint32_t length_middle, uint64_t* n_pairs);
int kmn_fetch_pairs(kmn_ctx* ctx, uint32_t batch_id, uint32_t sp, kmn_pair* out, uint64_t cap, uint64_t* n);
/* Grouping of block observations by anchor pair + per-group length selection on the device: replaces the
 * HashMap merge of merge_species_extraction (group_extraction.rs:335-357) by a stable radix sort of the
 * observation indices by (left, right) and select_unique_best_supported_length (group_sorting.rs:90-143)
 * by a segmented reduction.  Inputs are host arrays over the n observations of ALL species in emission
 * order (species-major): anchor k-mers, species index, block length.  perm[n] receives the sorted order
 * (ties keep emission order), groups[] one entry per distinct anchor pair in ascending (left, right). */
int kmn_group_observations(kmn_ctx* ctx, uint64_t n, const uint64_t* left, const uint64_t* right,
                           const uint32_t* sid, const uint32_t* len, uint32_t min_needed, uint32_t* perm,
                           kmn_group* groups, uint64_t cap_groups, uint64_t* n_groups);
/* The same grouping straight from the anchor pairs kmn_pair_batch left on the device (no host arrays are
 * uploaded): the observations are the pairs of the given batches, batch after batch and, inside a batch,
 * species after species in emission order (== kmn_fetch_pairs order); species i of batch b carries the species
 * index sid_base[b] + i * sid_stride.  perm indexes that concatenation.  With perm == NULL only *n_obs is set. */
int kmn_group_batches(kmn_ctx* ctx, const uint32_t* batch_ids, uint32_t n_batches, const uint32_t* sid_base,
                      uint32_t sid_stride, uint32_t min_needed, uint32_t* perm, uint64_t cap_perm, kmn_group* groups,
                      uint64_t cap_groups, uint64_t* n_obs, uint64_t* n_groups);
/* Occupancy per whitespace-stripped residue of species `sp` (records concatenated): estimate()
 * of the k-mer ENDING at that residue, 0 where none ends (roll shorter than k).  *n = residues.
 * (A parity / debugging call: kmn_scan_batch keeps exact estimates only where they are >= min_needed when it can
 * read the narrow table; this call then recomputes the batch's occupancies from the u16 table first.) */
int kmn_fetch_occupancy(kmn_ctx* ctx, uint32_t batch_id, uint32_t sp, uint16_t* occ, uint64_t cap,
                        uint64_t* n);

/* ---- --nj: pairwise counts ----------------------------------------------------------------
 * Replaces the integer part of lg_f81_distance (phylo.rs:84-95) for all pairs i<j of
 * compute_distance_matrix (phylo.rs:114-141).  mapped = n x L codes 0..19, 20 = gap/unknown
 * (map_sequence, phylo.rs:75-79).  valid / mismatch are n x n row-major u32; the strict upper
 * triangle is written.  The f64 correction (phylo.rs:96-105) stays on the host (glibc log). */
int kmn_pair_counts(kmn_ctx* ctx, const uint8_t* mapped, uint32_t n, uint64_t L, uint32_t* valid,
                    uint32_t* mismatch);
/* Rows [row_begin, row_end) of the same matrix: the row block one rank computes when the rayon loop
 * over rows i of compute_distance_matrix (phylo.rs:124-138) is sharded over GPUs.  Every rank passes
 * the whole `mapped`; row_begin must be a multiple of 64 and row_end a multiple of 64 or n.
 * valid / mismatch are (row_end - row_begin) x n row-major; entry (i - row_begin, j) is written for
 * j > i, the rest is 0.  kmn_pair_counts == rows [0, n).  valid / mismatch may be NULL: the counts
 * then stay on the device only (kmn_pair_counts_device), e.g. to be gathered by a collective. */
int kmn_pair_counts_rows(kmn_ctx* ctx, const uint8_t* mapped, uint32_t n, uint64_t L,
                         uint32_t row_begin, uint32_t row_end, uint32_t* valid, uint32_t* mismatch);
/* Device copies of the counts of the last kmn_pair_counts / kmn_pair_counts_rows call: *rows x *n
 * row-major u32 each, owned by the context, valid until the next such call or kmn_destroy. */
int kmn_pair_counts_device(kmn_ctx* ctx, const uint32_t** valid, const uint32_t** mismatch,
                           uint32_t* rows, uint32_t* n);

/* Device time (ms) of the last kmn_count_staged / kmn_scan_batch / kmn_pair_counts call on this
 * context, per stage, measured with CUDA events on the context's stream.
 * which: 0 frame, 1 bloom, 2 cms, 3 finalize (or exchange_reduce), 4 scan, 5 pair_counts,
 *        6 cms_partition_kernel alone, 7 bloom_bucket_kernel alone (single launches inside 2 / 1),
 *        8 pair_batch, 9 group_observations (kernels only),
 *        10 bloom_partition_kernel alone, 11 cms_apply_kernel (+ the idle spill / direct launches) alone;
 *        12 is not a time: the host-to-device rate (GB/s) the last kmn_count_batch call measured over its range
 *        copies (0 before the first such call; it steers the next call's range schedule). */
int kmn_last_stage_ms(kmn_ctx* ctx, int which, float* ms);
/* Number of kernel launches issued by this context since creation. */
uint64_t kmn_launch_count(kmn_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* KAMINO_B200_H */
