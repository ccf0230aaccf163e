// fasta_kernels.cuh — FASTA record / line detection on the device ("next" row 3 of SURVEY 8f).
//
// Replaces the record splitting of seq_io::fasta::Reader as the reference drives it
// (/root/reference/src/group_extraction.rs:368-375: `while let Some(rec) = reader.next()`, `rec.seq()`), restated
// on the CPU by oracle/fasta_io.cpp parse_fasta_bytes:
//   * a record starts at every LINE whose first byte is '>' (lines end with '\n' only);
//   * its header is the rest of that line, its sequence every byte up to the next such line (line terminators
//     included — the framing kernels skip them like any ASCII whitespace);
//   * before the first record only empty lines ("\n" or "\r\n") may appear.
// Input: the bytes of whole files back to back; every non-empty file ends with '\n' (the caller appends one
// to a file that lacks it — that adds whitespace to the last sequence, nothing else), so "first byte of a
// line" is a local property: offset 0, or the byte after a '\n'.
//
// Whether a byte belongs to a header line is a scan: per 16-byte chunk the state after the chunk is either
// the state before it (no line start inside) or decided by the chunk's last line start.  Three small kernels:
// per-tile summaries, a single-CTA composition over the tiles, and the pass that blanks the header bytes to
// spaces IN PLACE (a space is whitespace for the framing kernels and, unlike '\n', cannot fake a line start
// for a neighbouring thread that still has to look at the byte) and writes the record offsets.
// After that the existing pipeline (frame_kernels.cuh) runs unchanged on (raw, rec_off).
#pragma once
#include "frame_kernels.cuh"

namespace kmn {

struct ChunkScan {
    uint32_t hdr_starts;   // bit j: byte j starts a header line
    uint32_t line_starts;  // bit j: byte j is the first byte of a line
};

// c: the 16 bytes of a chunk, prev_nl: the byte before the chunk is '\n' (or the chunk starts the buffer)
__device__ __forceinline__ ChunkScan scan_chunk(const uint4& v, bool prev_nl) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    ChunkScan s{0u, 0u};
    bool ls = prev_nl;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t c = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        if (ls) {
            s.line_starts |= 1u << j;
            if (c == '>') s.hdr_starts |= 1u << j;
        }
        ls = c == '\n';
    }
    return s;
}

__device__ __forceinline__ bool last_byte_is_nl(const uint4& v) { return (v.w >> 24) == '\n'; }

// is the byte before this thread's chunk a '\n'?  (previous lane's last byte; lane 0 reads it from memory)
__device__ __forceinline__ bool prev_is_newline(const uint8_t* __restrict__ raw, uint64_t i0, const uint4& v) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t up = __shfl_up_sync(0xffffffffu, uint32_t(last_byte_is_nl(v)), 1);
    if (lane != 0) return up != 0;
    return i0 == 0 || raw[i0 - 1] == '\n';
}

// Pass A: per tile of FRAME_TILE bytes -> summary (bit 0: the tile has a line start, bit 1: the state after
// the tile's last line start = "inside a header line") and the number of header starts.
__global__ void __launch_bounds__(FRAME_THREADS)
fasta_summary_kernel(const uint8_t* __restrict__ raw, uint8_t* __restrict__ tile_sum, uint32_t* __restrict__ tile_hdr) {
    __shared__ uint32_t s_cnt[FRAME_WARPS], s_sum[FRAME_WARPS];
    const uint64_t i0 = uint64_t(blockIdx.x) * FRAME_TILE + uint64_t(threadIdx.x) * CHUNK;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(raw + i0));
    const ChunkScan s = scan_chunk(v, prev_is_newline(raw, i0, v));
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t cnt = __reduce_add_sync(0xffffffffu, uint32_t(__popc(s.hdr_starts)));
    const bool has = s.line_starts != 0;
    const bool val = has && ((s.hdr_starts >> (31 - __clz(s.line_starts))) & 1u);
    const uint32_t m_has = __ballot_sync(0xffffffffu, has), m_val = __ballot_sync(0xffffffffu, val);
    if (lane == 0) {
        s_cnt[warp] = cnt;
        s_sum[warp] = m_has ? (1u | (((m_val >> (31 - __clz(m_has))) & 1u) << 1)) : 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0, sum = 0;
#pragma unroll
        for (int i = 0; i < FRAME_WARPS; i++) {
            t += s_cnt[i];
            if (s_sum[i]) sum = s_sum[i];
        }
        tile_hdr[blockIdx.x] = t;
        tile_sum[blockIdx.x] = uint8_t(sum);
    }
}

// Pass B (one CTA): tile_in[t] = "the first byte of tile t continues a header line" = the state set by the
// nearest earlier tile that has a line start (offset 0 is one, so tile 0 always decides for itself).
__global__ void __launch_bounds__(1024)
fasta_tile_state_kernel(const uint8_t* __restrict__ tile_sum, uint32_t n_tiles, uint8_t* __restrict__ tile_in) {
    __shared__ uint32_t s_has[32], s_val[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t t = base + threadIdx.x;
        const uint32_t sum = t < n_tiles ? tile_sum[t] : 0u;
        const bool has = sum & 1u, val = (sum >> 1) & 1u;
        const uint32_t m_has = __ballot_sync(0xffffffffu, has), m_val = __ballot_sync(0xffffffffu, val);
        if (lane == 0) {
            s_has[warp] = m_has != 0;
            s_val[warp] = m_has ? (m_val >> (31 - __clz(m_has))) & 1u : 0u;
        }
        __syncthreads();
        uint32_t in = s_carry;
        for (unsigned w = 0; w < warp; w++)
            if (s_has[w]) in = s_val[w];
        const uint32_t before = m_has & ((1u << lane) - 1u);
        if (before) in = (m_val >> (31 - __clz(before))) & 1u;
        if (t < n_tiles) tile_in[t] = uint8_t(in);
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = has ? uint32_t(val) : in;
        __syncthreads();
    }
}

// Per file: only empty lines may precede the first record (seq_io: anything else is a parse error), and a
// non-empty file must end with '\n' (this library's input contract).  err = 1 + index of the first bad file,
// bit 31 set for the contract violation.
__global__ void fasta_check_kernel(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ file_off, uint32_t n_files,
                                   uint32_t* __restrict__ err) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_files) return;
    const uint64_t f0 = file_off[f], f1 = file_off[f + 1];
    if (f0 == f1) return;
    if (raw[f1 - 1] != '\n') { atomicMin(err, (f + 1) | 0x80000000u); }
    uint64_t p = f0;
    while (p < f1) {
        if (raw[p] == '\n') { p++; continue; }
        if (raw[p] == '\r' && p + 1 < f1 && raw[p + 1] == '\n') { p += 2; continue; }
        break;
    }
    if (p < f1 && raw[p] != '>') atomicMin(err, f + 1);
}

// Pass C: blank the header lines to spaces in place, write the header offset of every record
// (hdr_off[r], r = tile_rec_base[tile] + rank inside the tile) and the record offsets the framing kernels
// use: rec_off[r] = hdr_off[r] (the blanked header frames to nothing), rec_off[0] = 0 (leading empty lines).
__global__ void __launch_bounds__(FRAME_THREADS)
fasta_blank_kernel(uint8_t* __restrict__ raw, const uint8_t* __restrict__ tile_in, const uint32_t* __restrict__ tile_rec_base,
                   uint64_t* __restrict__ hdr_off, uint64_t* __restrict__ rec_off) {
    __shared__ uint32_t s_cnt[FRAME_WARPS], s_has[FRAME_WARPS], s_val[FRAME_WARPS];
    const uint64_t i0 = uint64_t(blockIdx.x) * FRAME_TILE + uint64_t(threadIdx.x) * CHUNK;
    uint4 v = *reinterpret_cast<const uint4*>(raw + i0);
    const ChunkScan s = scan_chunk(v, prev_is_newline(raw, i0, v));
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n = __popc(s.hdr_starts);
    uint32_t incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= unsigned(d)) incl += y;
    }
    const bool has = s.line_starts != 0;
    const bool val = has && ((s.hdr_starts >> (31 - __clz(s.line_starts))) & 1u);
    const uint32_t m_has = __ballot_sync(0xffffffffu, has), m_val = __ballot_sync(0xffffffffu, val);
    if (lane == 31) s_cnt[warp] = incl;
    if (lane == 0) {
        s_has[warp] = m_has != 0;
        s_val[warp] = m_has ? (m_val >> (31 - __clz(m_has))) & 1u : 0u;
    }
    __syncthreads();
    uint32_t rank = tile_rec_base[blockIdx.x] + incl - n;
    uint32_t in = tile_in[blockIdx.x];
    for (unsigned w = 0; w < warp; w++) {
        rank += s_cnt[w];
        if (s_has[w]) in = s_val[w];
    }
    const uint32_t before = m_has & ((1u << lane) - 1u);
    if (before) in = (m_val >> (31 - __clz(before))) & 1u;
    if (!in && !s.hdr_starts) return;   // a chunk of sequence lines (the common case): nothing to blank or record
    // walk the chunk: state flips only at line starts
    uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    bool hdr = in != 0, changed = false;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if ((s.line_starts >> j) & 1u) hdr = (s.hdr_starts >> j) & 1u;
        if ((s.hdr_starts >> j) & 1u) {
            hdr_off[rank] = i0 + j;
            rec_off[rank] = rank == 0 ? 0ull : i0 + j;
            rank++;
        }
        const uint32_t c = (w4[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        if (hdr && c != '\n' && c != ' ') {
            w4[j >> 2] = (w4[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (0x20u << (8 * (j & 3)));
            changed = true;
        }
    }
    if (changed) *reinterpret_cast<uint4*>(raw + i0) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
}

// sp_rec_off[s] = number of records that start before file s = lower_bound(hdr_off, file_off[s]);
// sp_rec_off[n_files] = n_rec.  Also closes rec_off with rec_off[n_rec] = n_raw.
__global__ void fasta_species_kernel(const uint64_t* __restrict__ hdr_off, uint32_t n_rec, const uint64_t* __restrict__ file_off,
                                     uint32_t n_files, uint64_t n_raw, uint64_t* __restrict__ sp_rec_off,
                                     uint64_t* __restrict__ rec_off) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) rec_off[n_rec] = n_raw;
    if (s > n_files) return;
    const uint64_t x = file_off[s];
    uint32_t lo = 0, hi = n_rec;   // first r with hdr_off[r] >= x
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (hdr_off[mid] < x) lo = mid + 1; else hi = mid;
    }
    sp_rec_off[s] = lo;
}

}  // namespace kmn
