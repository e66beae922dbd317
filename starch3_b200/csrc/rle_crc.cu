// rle_crc.cu -- kernel (3a): bzip2's initial run-length coding, block cut and
// block CRC for a batch of streams (one stream per chromosome).
//
// Reference: the byte-serial state machine of ADD_CHAR_TO_BLOCK /
// add_pair_to_block / copy_input_until_stop / handle_compress in the
// reference's vendored libbz2 (bz/bzlib.c:225-338, :370-412), for a stream fed
// with one BZ_FINISH action.  Restated for parallel execution:
//   * a "chunk" is a maximal run of equal bytes split every 255 bytes; a chunk
//     of length L emits min(L,4) copies plus, if L >= 4, the byte L-4;
//   * the emitted length is a prefix sum over input bytes, so output offsets
//     come from a device-wide scan and every byte is placed independently;
//   * blocks are whole chunks: a block closes after the first chunk that takes
//     it to >= nblockMAX bytes -- unless exactly one input byte is left in the
//     stream, which libbz2 flushes into the same block (bz/bzlib.c:393-396);
//   * the block CRC (CRC-32/BZIP2, bz/bzlib_private.h:157-172) covers the
//     pre-RLE bytes of the block's chunks, a contiguous input range.
#include "common.cuh"
#include "scan.cuh"

namespace s3g {

constexpr int RT = 256;            // threads per tile
constexpr int RB = 16;             // bytes per thread
constexpr int RTILE = RT * RB;     // 4096 input bytes per tile

struct StreamMap {
    const uint64_t *soff;   // [n_streams + 1]
    uint64_t n_streams;
    // is position i the first byte of a stream?  (only evaluated where bytes repeat)
    __device__ __forceinline__ bool is_start(uint64_t i) const
    {
        uint64_t lo = 0, hi = n_streams;            // find soff[k] == i
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            uint64_t v = soff[mid];
            if (v == i) return true;
            if (v < i) lo = mid + 1; else hi = mid;
        }
        return false;
    }
};

// One thread's 16 input bytes plus the byte before and after, kept as words so that the
// run structure comes from SIMD-in-register byte compares instead of per-byte branches.
struct Win {
    uint32_t x[4];       // bytes pos0 .. pos0+15, little endian
    uint32_t prevb;      // byte pos0-1 (0 if none)
    uint32_t nextb;      // byte pos0+16 (0 if none)
    uint64_t pos0;
    int cnt;             // valid bytes (0..RB)
};

__device__ __forceinline__ void load_win(const uint8_t *in, uint64_t n, uint64_t pos0, Win &w)
{
    w.pos0 = pos0;
    w.cnt = pos0 >= n ? 0 : (n - pos0 >= RB ? RB : (int)(n - pos0));
    if (w.cnt == RB) {
        uint4 v = *reinterpret_cast<const uint4 *>(in + pos0);
        w.x[0] = v.x; w.x[1] = v.y; w.x[2] = v.z; w.x[3] = v.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) if (q * 4 + k < w.cnt) v |= (uint32_t)in[pos0 + q * 4 + k] << (8 * k);
            w.x[q] = v;
        }
    }
    w.prevb = (pos0 > 0 && w.cnt > 0) ? in[pos0 - 1] : 0;
    w.nextb = (w.cnt == RB && pos0 + RB < n) ? in[pos0 + RB] : 0;
}

__device__ __forceinline__ uint32_t byte_of(const Win &w, int k) { return (w.x[k >> 2] >> (8 * (k & 3))) & 0xffu; }

// Stream starts inside the current tile, found once per CTA (streams are few and sorted).
struct TileStarts {
    uint64_t k0;        // index of the first stream start >= tile begin
    uint32_t count;     // stream starts in [tile begin, tile end]   (tile end = first byte of the next tile)
};
__device__ __forceinline__ void find_tile_starts(const StreamMap &sm, uint64_t tile_begin, TileStarts *ts)
{
    uint64_t lo = 0, hi = sm.n_streams + 1;          // soff has n_streams + 1 entries (the last is the end)
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (sm.soff[mid] < tile_begin) lo = mid + 1; else hi = mid; }
    uint64_t k = lo; uint32_t c = 0;
    while (k + c < sm.n_streams + 1 && sm.soff[k + c] <= tile_begin + RTILE) c++;
    ts->k0 = k; ts->count = c;
}

// bit k (0..16) set <=> a run starts at window byte k (bit 16: at the byte after the window).
// Bits at or beyond the end of the input are set: the input end closes the last chunk.
__device__ __forceinline__ uint32_t run_start_mask(const Win &w, const StreamMap &sm, const TileStarts &ts, uint64_t n)
{
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t prevw = q == 0 ? w.prevb : (w.x[q - 1] >> 24);
        uint32_t sh = (w.x[q] << 8) | prevw;                      // every byte's predecessor
        uint32_t ne = __vcmpne4(w.x[q], sh) & 0x01010101u;
        m |= ((ne * 0x01020408u) >> 24) << (4 * q);
    }
    if ((w.x[3] >> 24) != w.nextb) m |= 1u << 16;
    if (w.pos0 == 0) m |= 1u;
    for (uint32_t q = 0; q < ts.count; q++) {                     // almost always zero iterations
        uint64_t st = sm.soff[ts.k0 + q];
        if (st >= w.pos0 && st <= w.pos0 + RB) m |= 1u << (uint32_t)(st - w.pos0);
    }
    // positions >= n count as starts
    uint64_t left = n > w.pos0 ? n - w.pos0 : 0;
    if (left <= RB) m |= ~0u << (uint32_t)left;
    return m & 0x1ffffu;
}

// position + 1 of the last run start before `tile_begin` (0: none) -- what the max-scan over k_rle_runs' aggregates used to
// deliver.  A run starts where a byte differs from the one before it, at a stream start, and at position 0; the search
// walks back from the tile 32 bytes at a time (long runs are rare).  One warp; every lane gets the result.
__device__ __forceinline__ uint64_t tile_run_carry(const uint8_t *__restrict__ in, const StreamMap &sm, const TileStarts &ts, uint64_t tile_begin)
{
    if (tile_begin == 0) return 0;
    const unsigned l = threadIdx.x & 31;
    // the last stream start before the tile bounds the search
    const uint64_t floor = ts.k0 > 0 ? sm.soff[ts.k0 - 1] : 0;      // soff[0] == 0: a start at or before every position
    uint64_t hi = tile_begin;                                       // candidates p in [floor, hi)
    while (hi > floor) {
        const bool inr = hi >= (uint64_t)l + 1 && hi - 1 - l >= floor;
        const uint64_t p = hi - 1 - l;
        const bool st = inr && (p == floor || in[p] != in[p - 1]);
        const unsigned m = __ballot_sync(0xffffffffu, st);
        if (m) return hi - 1 - (uint64_t)(__ffs((int)m) - 1) + 1;
        if (hi < 32 + floor) break;
        hi -= 32;
    }
    return floor + 1;
}

// ---- pass 1 (kept for reference by the stage tests): last run start per tile (max-scan aggregate) --------------------
__global__ void __launch_bounds__(RT) k_rle_runs(const uint8_t *in, uint64_t n, StreamMap sm, uint64_t *agg)
{
    __shared__ uint64_t s[33];
    __shared__ TileStarts ts;
    if (threadIdx.x == 0) find_tile_starts(sm, (uint64_t)blockIdx.x * RTILE, &ts);
    __syncthreads();
    Win w;
    load_win(in, n, (uint64_t)blockIdx.x * RTILE + (uint64_t)threadIdx.x * RB, w);
    uint32_t m = w.cnt ? run_start_mask(w, sm, ts, n) & ((1u << w.cnt) - 1) : 0;
    uint64_t last = m ? w.pos0 + (31 - __clz(m)) + 1 : 0;   // position + 1 of the last run start among my bytes
    uint64_t tot;
    block_excl_max<uint64_t>(last, s, &tot);
    if (threadIdx.x == 0) agg[blockIdx.x] = tot;
}

// exclusive sum of the per-tile counts in two levels (1024 tiles per CTA, then the CTA totals): what one CTA of
// k_scan_agg did in 70 us
constexpr int SS_T = 1024;
__global__ void __launch_bounds__(SS_T) k_sum_tiles(uint64_t *v, uint64_t n, uint64_t *span_tot)
{
    __shared__ uint64_t sm[33];
    const uint64_t i = (uint64_t)blockIdx.x * SS_T + threadIdx.x;
    uint64_t a = i < n ? v[i] : 0, tot;
    uint64_t ex = block_excl_sum<uint64_t>(a, sm, &tot);
    if (i < n) v[i] = ex;
    if (threadIdx.x == 0) span_tot[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SS_T) k_sum_spans(uint64_t *v, uint64_t n, uint64_t *span, uint64_t nspans, uint64_t *total)
{
    // the spans' totals are few (n / 1024): every CTA scans them in shared memory, then adds its span's prefix to its tiles
    __shared__ uint64_t sm[33];
    __shared__ uint64_t s_pre, s_total;
    uint64_t carry = 0, mine = 0;
    for (uint64_t base = 0; base < nspans; base += SS_T) {
        const uint64_t k = base + threadIdx.x;
        uint64_t a = k < nspans ? span[k] : 0, tot;
        uint64_t ex = block_excl_sum<uint64_t>(a, sm, &tot);
        if (k == blockIdx.x) mine = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (blockIdx.x / SS_T * SS_T + threadIdx.x == blockIdx.x) s_pre = mine;     // the thread that met span blockIdx.x
    if (threadIdx.x == 0) s_total = carry;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * SS_T + threadIdx.x;
    if (i < n) v[i] += s_pre;
    if (blockIdx.x == 0 && threadIdx.x == 0 && total) *total = s_total;
}

struct MaxU64 {
    typedef uint64_t T;
    __host__ __device__ static T identity() { return 0; }
    __host__ __device__ static T op(T a, T b) { return a > b ? a : b; }
};
struct SumU64b {
    typedef uint64_t T;
    __host__ __device__ static T identity() { return 0; }
    __host__ __device__ static T op(T a, T b) { return a + b; }
};

// Per-thread RLE1 result for its RB bytes.
struct RleLocal {
    uint32_t emit_mask;   // bit k: byte k is copied (chunk offset < 4)
    uint32_t cnt_mask;    // bit k: byte k ends a chunk of length >= 4 (emits the count byte)
    uint32_t cnt_val[RB / 4];   // count bytes, packed like the input words
    uint32_t total;       // emitted bytes
};

// `carry` = position+1 of the last run start before this tile (exclusive max-scan).
__device__ __forceinline__ void rle_local(const Win &w, uint32_t startmask, uint64_t carry, uint64_t *s_max, RleLocal &r,
                                          uint32_t *c_first = nullptr)
{
    uint32_t own = w.cnt ? startmask & ((1u << w.cnt) - 1) : 0;
    uint64_t last = own ? w.pos0 + (31 - __clz(own)) + 1 : 0;
    uint64_t tot;
    uint64_t before = block_excl_max<uint64_t>(last, s_max, &tot);   // run start in effect before my first byte
    if (carry > before) before = carry;
    r.emit_mask = 0; r.cnt_mask = 0; r.total = 0;
#pragma unroll
    for (int q = 0; q < RB / 4; q++) r.cnt_val[q] = 0;
    if (w.cnt == 0) return;
    // chunk offset of my first byte
    uint32_t c;
    if (startmask & 1u) c = 0;
    else {
        uint64_t o64 = w.pos0 - (before - 1);
        c = o64 < 0xffffffffull ? (uint32_t)o64 % 255u : (uint32_t)(o64 % 255u);
    }
    if (c_first) *c_first = c;
    // fast path: every chunk offset in this window stays <= 2 -> every byte is copied, no count bytes.
    // Offsets grow by one per continuing byte, so that holds iff no three continue-bits are adjacent
    // and the run entering the window (offset c) does not reach 3 either.
    {
        uint32_t eqv = ~startmask & ((1u << w.cnt) - 1);      // bit k: byte k continues the run of byte k-1
        uint32_t lead = (uint32_t)__ffs((int)~(eqv >> 1)) - 1;  // continuing bytes right after byte 0
        if ((eqv & (eqv >> 1) & (eqv >> 2)) == 0 && c + lead <= 2) {
            r.emit_mask = (1u << w.cnt) - 1; r.total = (uint32_t)w.cnt;
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < RB; k++) {
        if (k < w.cnt) {
            if (k > 0) { c = (startmask >> k) & 1u ? 0u : (c == 254u ? 0u : c + 1u); }
            bool last_in_chunk = c == 254u || ((startmask >> (k + 1)) & 1u);
            if (c < 4) { r.emit_mask |= 1u << k; r.total++; }
            if (last_in_chunk && c >= 3) { r.cnt_mask |= 1u << k; r.cnt_val[k >> 2] |= (c - 3) << (8 * (k & 3)); r.total++; }
        }
    }
}

// ---- pass 2: emitted bytes per tile -----------------------------------------
__global__ void __launch_bounds__(RT) k_rle_emit_count(const uint8_t *in, uint64_t n, StreamMap sm, uint64_t *run_carry,
                                                        uint64_t *agg, uint16_t *win_e, uint8_t *win_c)
{
    __shared__ uint64_t s_max[33];
    __shared__ uint32_t s_sum[33];
    __shared__ TileStarts ts;
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) find_tile_starts(sm, (uint64_t)blockIdx.x * RTILE, &ts);
    __syncthreads();
    if (threadIdx.x < 32) {
        uint64_t c = tile_run_carry(in, sm, ts, (uint64_t)blockIdx.x * RTILE);
        if (threadIdx.x == 0) { s_carry = c; run_carry[blockIdx.x] = c; }       // k_rle_write reads it back
    }
    __syncthreads();
    Win w;
    load_win(in, n, (uint64_t)blockIdx.x * RTILE + (uint64_t)threadIdx.x * RB, w);
    uint32_t sm_mask = run_start_mask(w, sm, ts, n);
    RleLocal r;
    uint32_t c0 = 0;
    rle_local(w, sm_mask, s_carry, s_max, r, &c0);
    uint32_t tot;
    uint32_t ex = block_excl_sum<uint32_t>(r.total, s_sum, &tot);
    if (threadIdx.x == 0) agg[blockIdx.x] = tot;
    // per-window prefix (relative to the tile) and chunk offset of the window's first byte: the block
    // cut walks from a window start instead of re-deriving the state of a whole tile
    win_e[(uint64_t)blockIdx.x * RT + threadIdx.x] = (uint16_t)ex;
    win_c[(uint64_t)blockIdx.x * RT + threadIdx.x] = (uint8_t)c0;
}

// ---- pass 3: block cut, one warp per stream -----------------------------------
// The cut is a sequential chain over the blocks of a stream (each block starts where the
// previous one ended), so a stream is walked by one warp; the search for each block end is
// what the 32 lanes share: a 32-ary search over the per-tile prefix e_base and a cooperative
// scan of the one tile that holds the crossing.
struct CutWalker {
    const uint8_t *in; uint64_t n; StreamMap sm;
    const uint64_t *e_base;      // per tile: emitted bytes before the tile (exclusive scan), [ntiles] = total
    const uint16_t *win_e;       // per 16-byte window: emitted bytes before the window, relative to its tile
    const uint8_t *win_c;        // per window: chunk offset (0..254) of its first byte

    // serial walk from the start of window `win`; stops at the first chunk end x <= s1 with E(x) >= target
    // (or, when stop_at != ~0, exactly at position stop_at) and returns x and E(x)
    __device__ void walk(uint64_t win, uint64_t target, uint64_t s1, uint64_t stop_at, uint64_t *x_out, uint64_t *e_out) const
    {
        uint64_t i = win * RB;
        uint64_t e = e_base[win / RT] + win_e[win];
        uint32_t c = win_c[win];
        bool first = true;
        uint64_t x = s1;
        for (; i < s1; i++) {
            if (i == stop_at) { x = i; break; }
            uint8_t ch = in[i];
            if (!first) {
                bool st = ch != in[i - 1] || sm.is_start(i);
                c = st ? 0u : (c == 254u ? 0u : c + 1u);
            }
            first = false;
            bool last;
            if (c == 254u || i + 1 >= n) last = true;
            else if (in[i + 1] != ch) last = true;
            else last = sm.is_start(i + 1);
            if (c < 4) e++;
            if (last && c >= 3) e++;
            if (stop_at == ~0ull && last && e >= target) { x = i + 1; break; }
        }
        *x_out = x; *e_out = e;
    }

    // E(x): emitted bytes of all input bytes < x   (x is a stream boundary, hence a chunk boundary)
    __device__ uint64_t e_at(uint64_t x) const
    {
        uint64_t tile = x / RTILE;
        if (x == tile * RTILE) return e_base[tile];
        uint64_t xo, eo;
        walk(x / RB, 0, x, x, &xo, &eo);
        return eo;
    }

    // first chunk end x (<= s1) with E(x) >= target, given the last tile whose prefix is below the target
    __device__ void find(uint64_t tile, uint64_t target, uint64_t s1, uint64_t *x_out, uint64_t *e_out) const
    {
        const unsigned l = threadIdx.x & 31;
        // the crossing byte lies in the last window whose prefix is still below the target
        uint64_t eb = e_base[tile];
        uint32_t below = 0;
#pragma unroll
        for (int q = 0; q < RT / 32; q++) {
            uint64_t wi = tile * RT + (uint64_t)l * (RT / 32) + q;
            if (wi * RB < n && eb + win_e[wi] < target) below++;
        }
        uint32_t total = warp_incl_sum<uint32_t>(below);
        total = __shfl_sync(0xffffffffu, total, 31);
        uint64_t win = tile * RT + (total ? total - 1 : 0);
        uint64_t x = 0, e = 0;
        if (l == 0) walk(win, target, s1, ~0ull, &x, &e);
        *x_out = __shfl_sync(0xffffffffu, x, 0);
        *e_out = __shfl_sync(0xffffffffu, e, 0);
    }
};

// Blocks of stream s are written at provisional slots prov(s) + k and compacted afterwards.
__global__ void __launch_bounds__(32) k_rle_cut(CutWalker cw, uint32_t nmax, uint64_t slot_cap, BlockInfo *prov,
                                                uint32_t *blocks_per_stream, uint64_t *prov_base)
{
    const uint64_t s = blockIdx.x;
    const unsigned l = threadIdx.x & 31;
    if (s >= cw.sm.n_streams) return;
    uint64_t s0 = cw.sm.soff[s], s1 = cw.sm.soff[s + 1];
    uint64_t e0 = 0, e1 = 0;
    if (l == 0) { e0 = cw.e_at(s0); e1 = cw.e_at(s1); }
    e0 = __shfl_sync(0xffffffffu, e0, 0); e1 = __shfl_sync(0xffffffffu, e1, 0);
    uint64_t slot = e0 / nmax + s;     // upper bound on the blocks of earlier streams
    if (l == 0) prov_base[s] = slot;
    uint32_t nb = 0;
    uint64_t start = s0, base = e0;
    while (start < s1) {
        uint64_t target = base + nmax;
        uint64_t x = s1, ex = e1;
        if (e1 >= target) {
            // the last tile whose prefix is still below the target.  RLE1 rarely changes the length of text, so the tile is
            // first looked for in a window of 32 around where it would be if it changed nothing; the 32-ary search over
            // all tiles of the stream runs only if the window does not hold the crossing
            uint64_t lo = start / RTILE, hi = (s1 - 1) / RTILE;
            {
                const uint64_t guess = (start + nmax) / RTILE;
                const uint64_t w0 = guess > lo + 20 ? guess - 20 : lo;
                const uint64_t probe = w0 + l;
                const bool in = probe <= hi;
                const bool below = in && cw.e_base[probe] < target;
                const unsigned mb = __ballot_sync(0xffffffffu, below), mi = __ballot_sync(0xffffffffu, in);
                // monotone: the lanes below the target come first.  Usable if the first probe is below (or is the stream's
                // first tile) and some probe inside the stream is not below (or the window reaches the stream's last tile)
                const bool left_ok = (mb & 1u) || w0 == lo;
                const bool right_ok = (mi & ~mb) != 0 || w0 + 31 >= hi;
                if (left_ok && right_ok && mi) {
                    const unsigned cnt = __popc(mb);
                    lo = hi = cnt ? w0 + cnt - 1 : lo;
                }
            }
            while (lo < hi) {
                uint64_t span = hi - lo, stepw = (span + 31) / 32;
                uint64_t probe = lo + (uint64_t)(l + 1) * stepw;
                bool below = probe <= hi && cw.e_base[probe] < target;
                unsigned m = __ballot_sync(0xffffffffu, below);
                unsigned cnt = __popc(m);               // monotone: the first cnt probes are below
                uint64_t nlo = lo + (uint64_t)cnt * stepw;
                uint64_t nhi = lo + (uint64_t)(cnt + 1) * stepw - 1;
                lo = nlo; if (nhi < hi) hi = nhi;
                if (lo > hi) hi = lo;
            }
            cw.find(lo, target, s1, &x, &ex);
            if (x + 1 >= s1) { x = s1; ex = e1; }   // a single trailing byte is flushed into this block
        }
        if (l == 0 && slot + nb < slot_cap) {
            BlockInfo bi;
            memset(&bi, 0, sizeof bi);
            bi.in_start = start; bi.in_end = x; bi.e_base = base; bi.nblock = (uint32_t)(ex - base);
            bi.chrom = (uint32_t)s; bi.orig_ptr = -1;
            prov[slot + nb] = bi;
        }
        nb++;
        start = x; base = ex;
    }
    if (l == 0) blocks_per_stream[s] = nb;
}

__global__ void k_stream_block_scan(const uint32_t *blocks_per_stream, uint64_t n_streams, uint64_t *first_block,
                                    uint64_t *total)
{
    // tiny serial scan (streams are few); one thread
    if (blockIdx.x || threadIdx.x) return;
    uint64_t acc = 0;
    for (uint64_t s = 0; s < n_streams; s++) { first_block[s] = acc; acc += blocks_per_stream[s]; }
    first_block[n_streams] = acc;
    *total = acc;
}

__global__ void k_compact_blocks(const BlockInfo *prov, const uint64_t *prov_base, const uint64_t *first_block,
                                 uint64_t n_streams, BlockInfo *blocks, s3g_chrom *chroms)
{
    uint64_t s = blockIdx.x;
    if (s >= n_streams) return;
    uint64_t nb = first_block[s + 1] - first_block[s];
    for (uint64_t k = threadIdx.x; k < nb; k += blockDim.x) blocks[first_block[s] + k] = prov[prov_base[s] + k];
    if (threadIdx.x == 0 && chroms) chroms[s].n_blocks = (uint32_t)nb;
}

// packed layout of the block bytes: blk_off = exclusive sum of the slot sizes (one CTA, chunks of 1024 blocks)
// Only blocks [b_lo, b_hi) get a slot (a GPU that compresses a share of the blocks holds only their bytes).
__global__ void __launch_bounds__(1024) k_block_offsets_dev(BlockInfo *blocks, uint64_t b_lo, uint64_t b_hi, uint64_t *total)
{
    __shared__ uint64_t sm[33];
    uint64_t carry = 0;
    for (uint64_t b0 = b_lo; b0 < b_hi; b0 += 1024) {
        uint64_t b = b0 + threadIdx.x;
        uint64_t w = b < b_hi ? blk_slot_bytes(blocks[b].nblock) : 0, tot;
        uint64_t ex = block_excl_sum<uint64_t>(w, sm, &tot);
        if (b < b_hi) blocks[b].blk_off = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// ---- pass 4: write the RLE1 bytes into the block slots ----------------------
// A tile that lies inside one block (all but one tile in ~220) stages its output in shared memory at
// the same 16-byte phase as its destination and writes it out with aligned vector stores; the bytes
// in use are collected in a per-CTA flag table.  Tiles that touch a block boundary take the
// byte-by-byte path.
constexpr int RW_BUF = RTILE + RTILE / 4 + 32;        // a tile emits at most 5/4 of its bytes, plus the phase

__global__ void __launch_bounds__(RT) k_rle_write(const uint8_t *in, uint64_t n, StreamMap sm, const uint64_t *run_carry,
                                                   const uint64_t *e_base, const BlockInfo *blocks, uint64_t n_blocks,
                                                   uint8_t *blk_bytes, uint8_t *in_use, uint64_t tile0, uint64_t b_lo, uint64_t b_hi)
{
    __shared__ uint64_t s_max[33];
    __shared__ uint32_t s_sum[33];
    __shared__ TileStarts ts;
    __shared__ uint64_t s_bi;
    __shared__ int s_fast;
    __shared__ __align__(16) uint8_t s_buf[RW_BUF];
    __shared__ uint8_t s_used[256];
    const uint64_t tile = tile0 + blockIdx.x;            // the grid covers the tiles of the input range of blocks [b_lo, b_hi)
    const uint64_t tile_begin = tile * RTILE;
    if (threadIdx.x == 0) {
        find_tile_starts(sm, tile_begin, &ts);
        // block containing the tile's first byte: last block with in_start <= tile_begin
        uint64_t lo = 0, hi = n_blocks - 1;
        while (lo < hi) {
            uint64_t mid = (lo + hi + 1) >> 1;
            if (blocks[mid].in_start <= tile_begin) lo = mid; else hi = mid - 1;
        }
        uint64_t tile_end = tile_begin + RTILE < n ? tile_begin + RTILE : n;
        s_bi = lo;
        s_fast = blocks[lo].in_start <= tile_begin && tile_end <= blocks[lo].in_end;
    }
    s_used[threadIdx.x] = 0;
    __syncthreads();
    Win w;
    load_win(in, n, tile_begin + (uint64_t)threadIdx.x * RB, w);
    uint32_t sm_mask = run_start_mask(w, sm, ts, n);
    RleLocal r;
    rle_local(w, sm_mask, run_carry[tile], s_max, r);
    uint32_t tot;
    uint32_t ex = block_excl_sum<uint32_t>(r.total, s_sum, &tot);
    if (s_fast) {
        const uint64_t bi = s_bi;
        if (bi < b_lo || bi >= b_hi) return;               // another GPU's block
        const uint64_t d0 = e_base[tile] - blocks[bi].e_base;     // offset of the tile's output inside the block slot
        const uint32_t ph = (uint32_t)d0 & 15u;
        if (w.cnt) {
            uint32_t o = ph + ex;
#pragma unroll
            for (int k = 0; k < RB; k++) {
                if (k < w.cnt) {
                    uint8_t by = (uint8_t)byte_of(w, k);
                    if (r.emit_mask & (1u << k)) { s_buf[o++] = by; s_used[by] = 1; }
                    if (r.cnt_mask & (1u << k)) { uint8_t cv = (uint8_t)((r.cnt_val[k >> 2] >> (8 * (k & 3))) & 0xffu); s_buf[o++] = cv; s_used[cv] = 1; }
                }
            }
        }
        __syncthreads();
        uint8_t *dst = blk_bytes + blocks[bi].blk_off + (d0 - ph);        // 16-byte aligned
        const uint32_t lo_b = ph, hi_b = ph + tot;                                // valid bytes of s_buf
        for (uint32_t c = threadIdx.x * 16; c < hi_b; c += RT * 16) {
            if (c >= lo_b && c + 16 <= hi_b) *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(s_buf + c);
            else for (uint32_t j = c > lo_b ? c : lo_b; j < c + 16 && j < hi_b; j++) dst[j] = s_buf[j];
        }
        if (s_used[threadIdx.x]) in_use[bi * 256 + threadIdx.x] = 1;
        return;
    }
    if (w.cnt == 0) return;
    uint64_t e = e_base[tile] + ex;
    // block containing my first byte: last block with in_start <= pos0
    uint64_t lo = 0, hi = n_blocks - 1;
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (blocks[mid].in_start <= w.pos0) lo = mid; else hi = mid - 1;
    }
    uint64_t bi = lo;
    uint64_t b_end = blocks[bi].in_end, b_e0 = blocks[bi].e_base;
    uint8_t *dst = blk_bytes + blocks[bi].blk_off;
    uint8_t *use = in_use + bi * 256;
#pragma unroll
    for (int k = 0; k < RB; k++) {
        if (k < w.cnt) {
            uint64_t i = w.pos0 + k;
            if (i >= b_end) {
                bi++;
                b_end = blocks[bi].in_end; b_e0 = blocks[bi].e_base;
                dst = blk_bytes + blocks[bi].blk_off; use = in_use + bi * 256;
            }
            uint8_t by = (uint8_t)byte_of(w, k);
            const bool mine = bi >= b_lo && bi < b_hi;
            if (r.emit_mask & (1u << k)) { if (mine) { dst[e - b_e0] = by; use[by] = 1; } e++; }
            if (r.cnt_mask & (1u << k)) { uint8_t cv = (uint8_t)((r.cnt_val[k >> 2] >> (8 * (k & 3))) & 0xffu); if (mine) { dst[e - b_e0] = cv; use[cv] = 1; } e++; }
        }
    }
}

// ---- pass 5: block CRC -------------------------------------------------------
__device__ __forceinline__ uint32_t gf_mulmod(uint32_t a, uint32_t b)
{
    // a*b mod P over GF(2), P = x^32 + 0x04C11DB7, bit 31 = x^31
    uint32_t r = 0;
#pragma unroll 1
    for (int i = 31; i >= 0; i--) {
        uint32_t msb = r & 0x80000000u;
        r <<= 1;
        if (msb) r ^= 0x04C11DB7u;
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
// x^(8 * 2^j) mod P for j = 0 .. 39, computed once on the host (a kernel argument: nothing per device to set up); a shift by n
// bytes is the product of the entries of n's set bits -- ten multiplications on average where squaring on the device took forty
struct XPow8 { uint32_t p[40]; };
static uint32_t gf_mulmod_host(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 31; i >= 0; i--) {
        uint32_t msb = r & 0x80000000u;
        r <<= 1;
        if (msb) r ^= 0x04C11DB7u;
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
static const XPow8 &xpow8_table()
{
    static const XPow8 t = [] {
        XPow8 x;
        uint32_t sq = 0x100;           // x^8
        for (int j = 0; j < 40; j++) { x.p[j] = sq; sq = gf_mulmod_host(sq, sq); }
        return x;
    }();
    return t;
}
// c * x^(8*nbytes) mod P
__device__ __forceinline__ uint32_t gf_shift_bytes(uint32_t c, uint64_t nbytes, const XPow8 &X)
{
    for (int j = 0; nbytes && j < 40; j++, nbytes >>= 1)
        if (nbytes & 1) c = gf_mulmod(c, X.p[j]);
    return c;
}

constexpr int CRC_T = 512;
// CRC-32/BZIP2 (MSB first, bz/bzlib_private.h:157-172) of the pre-RLE bytes of a block.  The block is cut into `parts`
// pieces, a CTA each (a batch of few blocks would otherwise leave most SMs idle: ten blocks took as long as 148), every
// piece into 512 segments, each by slicing-by-4 (four table look-ups per aligned 32-bit word, 16-byte loads).  A segment's
// register is moved to the end of the BLOCK by a multiplication with x^(8 * bytes after it) mod P; what is left is a sum
// (XOR) over segments and pieces: the pieces meet in BlockInfo.crc by atomicXor, k_crc_finish complements the sum.
__global__ void __launch_bounds__(CRC_T) k_block_crc(const uint8_t *in, BlockInfo *blocks, uint32_t parts, const __grid_constant__ XPow8 X)
{
    __shared__ uint32_t tab[4][256];       // tab[k][b]: the CRC register after byte b and k zero bytes
    __shared__ uint32_t part[CRC_T];
    for (int i = threadIdx.x; i < 256; i += CRC_T) {
        uint32_t c = (uint32_t)i << 24;
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
        tab[0][i] = c;
    }
    __syncthreads();
    for (int k = 1; k < 4; k++) {
        for (int i = threadIdx.x; i < 256; i += CRC_T) { uint32_t c = tab[k - 1][i]; tab[k][i] = (c << 8) ^ tab[0][c >> 24]; }
        __syncthreads();
    }
    BlockInfo *bi = &blocks[blockIdx.x / parts];
    const uint32_t pc = blockIdx.x % parts;
    const uint64_t a = bi->in_start, len = bi->in_end - bi->in_start;
    // this CTA's piece [p0, p1) of the block, cut at multiples of 16 bytes
    const uint64_t per_piece = ((len + parts - 1) / parts + 15) & ~15ull;
    const uint64_t p0 = min(pc * per_piece, len), p1 = min(p0 + per_piece, len);
    const uint64_t plen = p1 - p0;
    uint64_t per = ((plen + CRC_T - 1) / CRC_T + 15) & ~15ull;
    uint64_t lo = p0 + (uint64_t)threadIdx.x * per, hi = lo + per;
    if (lo > p1) lo = p1;
    if (hi > p1) hi = p1;
    uint32_t c = (pc == 0 && threadIdx.x == 0) ? 0xFFFFFFFFu : 0u;     // only the block's first segment carries the init state
    const uint8_t *p = in + a;
    uint64_t i = lo;
    auto word = [&](uint32_t w) {
        c ^= __byte_perm(w, 0, 0x0123);                    // the first byte in memory is the most significant
        c = tab[3][c >> 24] ^ tab[2][(c >> 16) & 255u] ^ tab[1][(c >> 8) & 255u] ^ tab[0][c & 255u];
    };
    for (; i < hi && ((uintptr_t)(p + i) & 15); i++) c = (c << 8) ^ tab[0][(c >> 24) ^ p[i]];
    for (; i + 16 <= hi; i += 16) {
        uint4 v = *reinterpret_cast<const uint4 *>(p + i);
        word(v.x); word(v.y); word(v.z); word(v.w);
    }
    for (; i < hi; i++) c = (c << 8) ^ tab[0][(c >> 24) ^ p[i]];
    // shift by the bytes of the block that follow this segment
    uint64_t after = len - hi;
    if (c != 0 && after) c = gf_shift_bytes(c, after, X);
    part[threadIdx.x] = c;
    __syncthreads();
    for (int d = CRC_T / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) part[threadIdx.x] ^= part[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0 && (part[0] || pc == 0)) atomicXor(&bi->crc, part[0]);
}
__global__ void k_crc_finish(BlockInfo *blocks, uint64_t nb)
{
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) blocks[b].crc = ~blocks[b].crc;
}
// pieces per block for a batch of nb blocks
static uint32_t crc_parts(uint64_t nb)
{
    uint32_t parts = 1;
    while (parts < 16 && (uint64_t)parts * 2 * nb <= 2 * (uint64_t)SM_COUNT) parts *= 2;
    return parts;
}

// unseqToSeq map + nInUse per block (makeMaps_e, bz/compress.c:106-115)
__global__ void k_block_maps(const uint8_t *in_use, BlockInfo *blocks, uint8_t *seq_map)
{
    __shared__ uint32_t s[33];
    uint64_t b = blockIdx.x;
    uint32_t u = in_use[b * 256 + threadIdx.x] ? 1u : 0u;
    uint32_t tot;
    uint32_t ex = block_excl_sum<uint32_t>(u, s, &tot);
    seq_map[b * 256 + threadIdx.x] = (uint8_t)ex;
    if (threadIdx.x == 0) blocks[b].n_in_use = tot;
}

// block plan of the streams in d_in: cut points, sizes (ctx->blocks, mirrored in ctx->h_blocks).  d_in and d_soff must
// stay valid until the last run_rle_fill of the plan.
int run_rle_plan(Ctx *ctx, const uint8_t *d_in, uint64_t n, const uint64_t *d_soff, uint64_t n_streams, int level,
                 CutResult *out)
{
    *out = CutResult();
    uint32_t nmax = 100000u * (uint32_t)level - 19;     // bz/bzlib.c:194
    uint64_t ntiles = (n + RTILE - 1) / RTILE;
    if (ntiles == 0) ntiles = 1;
    if (ntiles > 0x7fffffffull) { set_error("stream buffer too large"); return S3G_E_LIMIT; }
    S3G_TRY(ctx->rle_carry.ensure((ntiles + 2) * 8));
    S3G_TRY(ctx->rle_ebase.ensure((ntiles + 2) * 8));
    S3G_TRY(ctx->scalars.ensure(64 * 8));
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    uint64_t *run_carry = ctx->rle_carry.as<uint64_t>(), *e_base = ctx->rle_ebase.as<uint64_t>();
    StreamMap sm{d_soff, n_streams};
    S3G_TRY(ctx->io_c.ensure((ntiles + 1) * RT * 2));
    S3G_TRY(ctx->io_e.ensure((ntiles + 1) * RT));
    S3G_BYTES(ctx, n + n * 3 / 16);
    S3G_LAUNCH(ctx, k_rle_emit_count, (unsigned)ntiles, RT, 0, d_in, n, sm, run_carry, e_base, ctx->io_c.as<uint16_t>(), ctx->io_e.as<uint8_t>());
    const uint64_t nspans = (ntiles + SS_T - 1) / SS_T;
    S3G_TRY(ctx->io_d.ensure((nspans + 1) * 8));
    S3G_LAUNCH(ctx, k_sum_tiles, (unsigned)nspans, SS_T, 0, e_base, ntiles, ctx->io_d.as<uint64_t>());
    S3G_LAUNCH(ctx, k_sum_spans, (unsigned)nspans, SS_T, 0, e_base, ntiles, ctx->io_d.as<uint64_t>(), nspans, d_sc + 16);
    // e_base[ntiles] = total, so E() can be evaluated at n
    S3G_CUDA(cudaMemcpyAsync(e_base + ntiles, d_sc + 16, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 16, d_sc + 16, 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    S3G_TRY(check_launch("rle scan"));
    uint64_t e_total = ctx->h_scalars[16];
    uint64_t slot_cap = e_total / nmax + n_streams + 2;
    S3G_TRY(ctx->blk_prov.ensure(slot_cap * sizeof(BlockInfo)));       // provisional slots
    S3G_TRY(ctx->blocks.ensure(slot_cap * sizeof(BlockInfo)));
    S3G_TRY(ctx->stream_tab.ensure((n_streams + 2) * 8 * 3));
    uint64_t *prov_base = ctx->stream_tab.as<uint64_t>();
    uint64_t *first_block = prov_base + (n_streams + 2);
    uint32_t *bps = reinterpret_cast<uint32_t *>(first_block + (n_streams + 2));
    CutWalker cw{d_in, n, sm, e_base, ctx->io_c.as<uint16_t>(), ctx->io_e.as<uint8_t>()};
    S3G_LAUNCH(ctx, k_rle_cut, (unsigned)n_streams, 32, 0, cw, nmax, slot_cap, ctx->blk_prov.as<BlockInfo>(), bps, prov_base);
    S3G_LAUNCH(ctx, k_stream_block_scan, 1, 1, 0, bps, n_streams, first_block, d_sc + 17);
    S3G_LAUNCH(ctx, k_compact_blocks, (unsigned)n_streams, 64, 0, ctx->blk_prov.as<BlockInfo>(), prov_base, first_block,
               n_streams, ctx->blocks.as<BlockInfo>(), (s3g_chrom *)nullptr);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 17, d_sc + 17, 8, cudaMemcpyDeviceToHost, ctx->stream));
    // host mirror of the block table (sizes drive batching and the byte accounting of the later stages)
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    S3G_TRY(check_launch("rle cut"));
    uint64_t nb = ctx->h_scalars[17];
    out->n_blocks = nb;
    out->rle_bytes = e_total;
    ctx->rle_in = d_in; ctx->rle_n = n; ctx->rle_soff = d_soff; ctx->rle_streams = n_streams; ctx->rle_blocks = nb;
    ctx->h_blocks.resize(nb);
    if (nb == 0) return S3G_OK;
    S3G_TRY(ctx->in_use.ensure(nb * 256));
    S3G_TRY(ctx->seq_map.ensure(nb * 256));
    S3G_CUDA(cudaMemcpyAsync(ctx->h_blocks.data(), ctx->blocks.p, nb * sizeof(BlockInfo), cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

// RLE1 bytes, CRC, bytes in use and unseqToSeq maps of blocks [b_lo, b_hi) of the plan left by run_rle_plan
int run_rle_fill(Ctx *ctx, uint64_t b_lo, uint64_t b_hi)
{
    const uint64_t nb = ctx->rle_blocks;
    if (b_lo >= b_hi || b_hi > nb) return S3G_OK;
    const uint8_t *d_in = ctx->rle_in;
    const uint64_t n = ctx->rle_n;
    StreamMap sm{ctx->rle_soff, ctx->rle_streams};
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    uint64_t *run_carry = ctx->rle_carry.as<uint64_t>(), *e_base = ctx->rle_ebase.as<uint64_t>();
    BlockInfo *blocks = ctx->blocks.as<BlockInfo>();
    uint64_t packed = 0;
    for (uint64_t b = b_lo; b < b_hi; b++) { ctx->h_blocks[b].blk_off = packed; packed += blk_slot_bytes(ctx->h_blocks[b].nblock); }
    S3G_TRY(ctx->blk_bytes.ensure(packed + 256));
    S3G_LAUNCH(ctx, k_block_offsets_dev, 1, 1024, 0, blocks, b_lo, b_hi, d_sc + 18);
    S3G_CUDA(cudaMemsetAsync(ctx->in_use.as<uint8_t>() + b_lo * 256, 0, (b_hi - b_lo) * 256, ctx->stream));
    const uint64_t in_lo = ctx->h_blocks[b_lo].in_start, in_hi = ctx->h_blocks[b_hi - 1].in_end;
    const uint64_t tile0 = in_lo / RTILE, tile1 = (in_hi + RTILE - 1) / RTILE;
    double e_own = 0;
    for (uint64_t b = b_lo; b < b_hi; b++) e_own += ctx->h_blocks[b].nblock;
    S3G_BYTES(ctx, (double)(in_hi - in_lo) + e_own);
    if (tile1 > tile0)
        S3G_LAUNCH(ctx, k_rle_write, (unsigned)(tile1 - tile0), RT, 0, d_in, n, sm, run_carry, e_base, blocks, nb,
                   ctx->blk_bytes.as<uint8_t>(), ctx->in_use.as<uint8_t>(), tile0, b_lo, b_hi);
    S3G_BYTES(ctx, (double)(in_hi - in_lo));
    {
        // the block table's crc fields are zero (the cut leaves them so): the pieces XOR into them
        const uint32_t parts = crc_parts(b_hi - b_lo);
        S3G_LAUNCH(ctx, k_block_crc, (unsigned)((b_hi - b_lo) * parts), CRC_T, 0, d_in, blocks + b_lo, parts, xpow8_table());
        S3G_LAUNCH(ctx, k_crc_finish, (unsigned)((b_hi - b_lo + 255) / 256), 256, 0, blocks + b_lo, b_hi - b_lo);
    }
    S3G_LAUNCH(ctx, k_block_maps, (unsigned)(b_hi - b_lo), 256, 0, ctx->in_use.as<uint8_t>() + b_lo * 256, blocks + b_lo, ctx->seq_map.as<uint8_t>() + b_lo * 256);
    // the mirror learns CRC and alphabet size (the later stages size their launches from it)
    S3G_CUDA(cudaMemcpyAsync(ctx->h_blocks.data() + b_lo, blocks + b_lo, (b_hi - b_lo) * sizeof(BlockInfo), cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return check_launch("rle write");
}

// CRC-32/BZIP2 of d_in[in_start, in_end) of each of the nb descriptors -> BlockInfo.crc (the decoder checks its output with it)
int run_block_crc(Ctx *ctx, const uint8_t *d_in, BlockInfo *d_blocks, uint64_t nb, double bytes)
{
    if (!nb) return S3G_OK;
    S3G_BYTES(ctx, bytes);
    const uint32_t parts = crc_parts(nb);                  // the descriptors' crc fields must be zero on entry
    S3G_LAUNCH(ctx, k_block_crc, (unsigned)(nb * parts), CRC_T, 0, d_in, d_blocks, parts, xpow8_table());
    S3G_LAUNCH(ctx, k_crc_finish, (unsigned)((nb + 255) / 256), 256, 0, d_blocks, nb);
    return check_launch("block crc");
}

int run_rle_cut(Ctx *ctx, const uint8_t *d_in, uint64_t n, const uint64_t *d_soff, uint64_t n_streams, int level,
                CutResult *out)
{
    S3G_TRY(run_rle_plan(ctx, d_in, n, d_soff, n_streams, level, out));
    return run_rle_fill(ctx, 0, out->n_blocks);
}

}  // namespace s3g
