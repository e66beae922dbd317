// tokenize_transform.cu -- kernels (1) and (2) of the hot path, fused.
//
// (1) newline scan + tab split + integer parse + chromosome-boundary flags:
//     replaces produce_line / consume_line (hpp:158-199, :201-309) and the
//     strcmp chromosome test (hpp:325-342).
// (2) the starch coordinate transform update_transformation_state (hpp:428-504)
//     with the per-chromosome reset (hpp:523-532) and the statistics the
//     reference declares but never computes (hpp:61-62).
// hpp = /root/reference/include/starch3api.hpp.
//
// The per-line dependencies are radius-1 (previous stop, previous length, previous chromosome); everything
// else is prefix sums (output offsets, line and chromosome indices) and one segmented running maximum
// (unique bases).  Nothing per line is kept in HBM: the input is read twice, in tiles of 8 KiB, and only the
// transformed bytes are written.
//
//   k_front_measure  a tile owns the lines that END in it.  It stages its bytes (plus the 512 bytes before
//                    them) in shared memory with 16-byte loads, builds newline and tab bit masks, parses its
//                    lines (a thread per line, fields found in the tab mask), and reduces
//                    {lines, output bytes, chromosome starts, running max of stop}.  The exclusive prefix of
//                    every tile comes from a decoupled look-back over the earlier tiles in the same launch.
//   k_front_write    the same tile walk with the prefix known: formats the transformed lines into shared
//                    memory at the destination's 16-byte phase and stores them as aligned vectors; chromosome
//                    table seeds, per-chromosome sums (one atomic per tile), optionally the per-line arrays
//                    (the s3g_tokenize parity entry point).
//   k_chrom_finish   chromosome table from the seeds and sums.
#include "common.cuh"
#include "scan.cuh"

namespace s3g {

constexpr int FT = 8192;                 // input bytes per tile
constexpr int FBACK = 512;               // staged bytes before the tile (the line that ends before it, and the one before)
constexpr int FTH = 256;                 // threads per tile
constexpr int FSTAGE = FBACK + FT;
constexpr int FMAXL = FT / 3 + 2;        // a line with three fields has at least 3 bytes ("\t\t\n")
constexpr int FOBUF = FT + 2048;         // staged output bytes of a tile (more: direct stores)

// scalar slots (ctx->scalars, u64 each)
enum { SC_MALFORMED = 1, SC_LASTNL = 2, SC_LINE1 = 3, SC_TICKET = 4, SC_ERROR = 5, SC_TOTAL = 8 /* FAgg: 5 slots */ };

// what a tile (or a prefix of tiles) contributes; fagg_op is associative, operands in input order
struct FAgg {
    uint64_t lines, out, chroms;
    int64_t v;            // largest stop since the last chromosome start (that line included)
    uint32_t seg, pad;    // a chromosome start lies inside
};
__host__ __device__ inline FAgg fagg_identity() { FAgg a; a.lines = 0; a.out = 0; a.chroms = 0; a.v = INT64_MIN; a.seg = 0; a.pad = 0; return a; }
__host__ __device__ inline FAgg fagg_op(const FAgg &a, const FAgg &b)
{
    FAgg r;
    r.lines = a.lines + b.lines; r.out = a.out + b.out; r.chroms = a.chroms + b.chroms; r.pad = 0;
    if (b.seg) { r.v = b.v; r.seg = 1; }
    else { r.v = a.v > b.v ? a.v : b.v; r.seg = a.seg; }
    return r;
}

// segmented running max as a scan.cuh functor
struct SegMax {
    int64_t v; int32_t seg; int32_t pad;
};
struct SegMaxF {
    typedef SegMax T;
    __host__ __device__ static T identity() { T t; t.v = INT64_MIN; t.seg = 0; t.pad = 0; return t; }
    __host__ __device__ static T op(T a, T b)
    {
        if (b.seg) return b;
        T r; r.v = a.v > b.v ? a.v : b.v; r.seg = a.seg; r.pad = 0; return r;
    }
};

// bit k of the result set <=> byte k of the 16 bytes in w equals the byte replicated in `pat`
__device__ __forceinline__ unsigned eq_mask16(const uint32_t w[4], uint32_t pat)
{
    unsigned m = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        unsigned b = __vcmpeq4(w[q], pat) & 0x01010101u;
        b = (b | (b >> 7) | (b >> 14) | (b >> 21)) & 0xfu;
        m |= b << (4 * q);
    }
    return m;
}

__constant__ uint64_t POW10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull,
                                   10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull,
                                   1000000000000000ull, 10000000000000000ull, 100000000000000000ull, 1000000000000000000ull,
                                   10000000000000000000ull};
// decimal digits of a (1..20): floor(bits * log10(2)) is the count or one short of it (no division)
__device__ __forceinline__ int dec_digits(uint64_t a)
{
    int b = 64 - __clzll((long long)(a | 1));
    int t = (b * 1233) >> 12;
    int d = t + (a >= POW10[t] ? 1 : 0);
    return d < 1 ? 1 : d;
}
__device__ __forceinline__ int dec_len(int64_t v)
{
    // printed length of "%lld": n_digits (hpp:559-581) plus the sign it does not count
    uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    return dec_digits(a) + (v < 0);
}
__device__ __forceinline__ uint32_t put_dec(uint8_t *dst, int64_t v)
{
    uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    uint32_t k = 0;
    if (v < 0) dst[k++] = '-';
    const int nd = dec_digits(a);
    if (a <= 0xffffffffull) {            // the usual case: 32-bit divisions by a constant
        uint32_t x = (uint32_t)a;
        for (int j = nd - 1; j >= 0; j--) { uint32_t q = x / 10u; dst[k + j] = (uint8_t)('0' + (x - q * 10u)); x = q; }
    } else {
        for (int j = nd - 1; j >= 0; j--) { uint64_t q = a / 10u; dst[k + j] = (uint8_t)('0' + (uint32_t)(a - q * 10u)); a = q; }
    }
    return k + nd;
}

// ---- one tile's view of the input -------------------------------------------------------------------
struct TileSh {
    __align__(16) uint8_t buf[FSTAGE];       // bytes [lo, lo + FSTAGE)
    uint32_t nlw[FSTAGE / 32];               // newline bits, bit i of word w <=> byte lo + 32 w + i
    uint32_t tbw[FSTAGE / 32];               // tab bits
    uint16_t nl[FMAXL];                      // newlines inside the tile, offsets from tile0, ascending
    int64_t st[FTH], sp[FTH];                // start / stop of the lines of the current round
    uint32_t sum_a[33], sum_b[33];
    SegMax seg_sm[FTH / 32];
    uint64_t lo, tile0, start0, prev_start;
    int64_t carry_start, carry_stop;         // start / stop of the line before the current round
    uint32_t k, has_prev, has_prev2, tile;
};

struct Parsed {
    int64_t start, stop;
    uint32_t rem_off, name_len, malformed;
};

// sscanf("%lld") over the documented domain (hpp:306-307): [sign] digits, up to the first other byte
__device__ __forceinline__ int64_t parse_int(const uint8_t *p, uint32_t len)
{
    uint64_t acc = 0;
    int neg = 0, st = 0;                      // 0 = expecting sign/digit, 1 = in digits
    for (uint32_t i = 0; i < len; i++) {
        uint8_t c = p[i];
        if (st == 0 && (c == '-' || c == '+')) { neg = (c == '-'); st = 1; }
        else if (c >= '0' && c <= '9') { acc = acc * 10 + (uint64_t)(c - '0'); st = 1; }
        else break;
    }
    return neg ? (int64_t)(0 - acc) : (int64_t)acc;
}

// first tab in [p, e) of the staged bytes, or e
__device__ __forceinline__ uint64_t next_tab(const TileSh &S, uint64_t p, uint64_t e)
{
    if (p >= e) return e;
    uint32_t i = (uint32_t)(p - S.lo), iend = (uint32_t)(e - S.lo);
    uint32_t w = i >> 5;
    uint32_t m = S.tbw[w] & (0xffffffffu << (i & 31));
    while (m == 0) {
        w++;
        if ((w << 5) >= iend) return e;
        m = S.tbw[w];
    }
    uint32_t q = (w << 5) + (uint32_t)__ffs((int)m) - 1;
    return q < iend ? S.lo + q : e;
}

// line [s, e), bed[e] == '\n'; fields as consume_line finds them (hpp:220-309)
__device__ __forceinline__ Parsed parse_line(const TileSh &S, const uint8_t *bed, uint64_t s, uint64_t e)
{
    Parsed P;
    P.start = 0; P.stop = 0; P.malformed = 0;
    uint64_t t1, t2, t3;
    const uint8_t *src;                                // src[p - off] = byte at absolute position p
    uint64_t off;
    if (s >= S.lo) {                                   // staged: tabs from the bit mask
        t1 = next_tab(S, s, e);
        t2 = t1 < e ? next_tab(S, t1 + 1, e) : e;
        t3 = t2 < e ? next_tab(S, t2 + 1, e) : e;
        src = S.buf; off = S.lo;
    } else {                                           // a line longer than the staged window: bytes from global memory
        uint64_t q = s;
        while (q < e && bed[q] != '\t') q++;
        t1 = q;
        if (q < e) { q++; while (q < e && bed[q] != '\t') q++; }
        t2 = q;
        if (q < e) { q++; while (q < e && bed[q] != '\t') q++; }
        t3 = q;
        src = bed; off = 0;
    }
    P.name_len = (uint32_t)(t1 - s);
    if (t2 >= e) {                                     // fewer than three fields
        P.malformed = 1;
        P.name_len = (uint32_t)((t1 < e ? t1 : e) - s);
        P.rem_off = (uint32_t)(e - s);
        return P;
    }
    P.start = parse_int(src + (t1 + 1 - off), (uint32_t)(t2 - t1 - 1));
    P.stop = parse_int(src + (t2 + 1 - off), (uint32_t)(t3 - t2 - 1));
    P.rem_off = (uint32_t)((t3 < e ? t3 + 1 : e) - s);
    return P;
}

__device__ __forceinline__ uint8_t byte_at(const TileSh &S, const uint8_t *bed, uint64_t pos)
{
    return pos >= S.lo ? S.buf[pos - S.lo] : bed[pos];
}

// largest p in [floor, hi) with bed[p] == '\n', or -1; one warp
__device__ __forceinline__ int64_t scan_back(const uint8_t *bed, uint64_t hi, uint64_t floor)
{
    const unsigned l = threadIdx.x & 31;
    while (hi > floor) {
        bool hit = hi >= (uint64_t)l + 1 && hi - 1 - l >= floor && bed[hi - 1 - l] == '\n';
        unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) return (int64_t)(hi - 1 - (uint64_t)(__ffs((int)m) - 1));
        if (hi < 32 + floor) break;
        hi -= 32;
    }
    return -1;
}

// stage the tile, build the masks, list its newlines, find where its first line and the line before it start
__device__ __forceinline__ void tile_setup(TileSh &S, const uint8_t *__restrict__ bed, uint64_t n, uint32_t skip, uint32_t tile)
{
    const uint64_t tile0 = (uint64_t)tile * FT;
    const uint32_t back = tile ? FBACK : 0;
    const uint64_t lo = tile0 - back;
    if (threadIdx.x == 0) { S.lo = lo; S.tile0 = tile0; S.tile = tile; }
    uint16_t *nl16 = reinterpret_cast<uint16_t *>(S.nlw), *tb16 = reinterpret_cast<uint16_t *>(S.tbw);
    for (uint32_t c = threadIdx.x; c < (back + FT) / 16; c += FTH) {
        uint64_t pos = lo + 16ull * c;
        uint32_t w[4] = {0, 0, 0, 0};
        if (pos + 16 <= n) {
            uint4 v = *reinterpret_cast<const uint4 *>(bed + pos);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else if (pos < n) {
            for (uint32_t k = 0; k < (uint32_t)(n - pos); k++) w[k >> 2] |= (uint32_t)bed[pos + k] << (8 * (k & 3));
        }
        *reinterpret_cast<uint4 *>(S.buf + 16 * c) = make_uint4(w[0], w[1], w[2], w[3]);
        unsigned mn = eq_mask16(w, 0x0a0a0a0au), mt = eq_mask16(w, 0x09090909u);
        if (pos < skip) mn &= 0xffffffffu << (skip - (uint32_t)pos);     // the bytes before the range: not ours (skip < 16)
        nl16[c] = (uint16_t)mn; tb16[c] = (uint16_t)mt;
    }
    __syncthreads();
    // newlines inside the tile -> compact list
    {
        uint32_t w = S.nlw[back / 32 + threadIdx.x];
        uint32_t tot;
        uint32_t ex = block_excl_sum<uint32_t>(__popc(w), S.sum_a, &tot);
        while (w) {
            int b = __ffs((int)w) - 1;
            w &= w - 1;
            if (ex < FMAXL) S.nl[ex] = (uint16_t)(threadIdx.x * 32 + b);
            ex++;
        }
        if (threadIdx.x == 0) S.k = tot;
    }
    // the last two newlines before the tile
    if (threadIdx.x < 32) {
        const unsigned l = threadIdx.x;
        int64_t nl1 = -1, nl2 = -1;
        if (tile) {
            uint32_t w = l < FBACK / 32 ? S.nlw[l] : 0;
            unsigned nz = __ballot_sync(0xffffffffu, w != 0);
            if (nz) {
                int hl = 31 - __clz((int)nz);
                uint32_t hw = __shfl_sync(0xffffffffu, w, hl);
                int hb = 31 - __clz((int)hw);
                nl1 = (int64_t)(lo + (uint64_t)hl * 32 + hb);
                uint32_t hw2 = hw & ~(1u << hb);
                unsigned nz2 = nz & ~(1u << hl);
                if (hw2) nl2 = (int64_t)(lo + (uint64_t)hl * 32 + (31 - __clz((int)hw2)));
                else if (nz2) {
                    int hl2 = 31 - __clz((int)nz2);
                    uint32_t w2 = __shfl_sync(0xffffffffu, w, hl2);
                    nl2 = (int64_t)(lo + (uint64_t)hl2 * 32 + (31 - __clz((int)w2)));
                }
                if (nl2 < 0) nl2 = scan_back(bed, lo, skip);
            } else {
                nl1 = scan_back(bed, lo, skip);
                if (nl1 >= 0) nl2 = scan_back(bed, (uint64_t)nl1, skip);
            }
        }
        if (l == 0) {
            S.has_prev = nl1 >= 0; S.has_prev2 = nl2 >= 0;
            S.start0 = nl1 >= 0 ? (uint64_t)nl1 + 1 : skip;
            S.prev_start = nl2 >= 0 ? (uint64_t)nl2 + 1 : skip;
        }
    }
    __syncthreads();
    // start / stop of the line before the tile's first line: what its first line's deltas refer to
    if (threadIdx.x == 0) {
        S.carry_start = 0; S.carry_stop = 0;
        if (S.has_prev) {
            Parsed P = parse_line(S, bed, S.prev_start, S.start0 - 1);
            S.carry_start = P.start; S.carry_stop = P.stop;
        }
    }
    __syncthreads();
}

// what a line contributes, given the line before it
struct LineOut {
    uint64_t s, e;
    Parsed P;
    uint32_t flag;          // chromosome differs from the previous line (hpp:331), or first line of the input
    uint32_t first_global;  // first line of the input
    uint32_t out_len;
    int64_t pstop, plen;
};

// round `base`: thread t takes the tile's line base + t
__device__ __forceinline__ bool line_of_round(TileSh &S, const uint8_t *bed, uint32_t base, uint32_t halo, LineOut &L)
{
    const uint32_t j = base + threadIdx.x;
    const bool act = j < S.k;
    L.flag = 0; L.first_global = 0; L.out_len = 0; L.pstop = 0; L.plen = 0;
    L.P.start = 0; L.P.stop = 0; L.P.rem_off = 0; L.P.name_len = 0; L.P.malformed = 0; L.s = 0; L.e = 0;
    if (act) {
        L.s = j == 0 ? S.start0 : S.tile0 + S.nl[j - 1] + 1;
        L.e = S.tile0 + S.nl[j];
        L.P = parse_line(S, bed, L.s, L.e);
        L.first_global = (j == 0 && !S.has_prev) ? 1u : 0u;
        if (L.first_global) L.flag = 1;
        else {
            // strcmp(chr, previous chr) != 0 (hpp:331); the previous line's field ends at its first tab
            uint64_t ps = j == 0 ? S.prev_start : (j == 1 ? S.start0 : S.tile0 + S.nl[j - 2] + 1);
            uint32_t cl = L.P.name_len;
            bool diff = false;
            for (uint32_t q = 0; q < cl; q++)
                if (byte_at(S, bed, L.s + q) != byte_at(S, bed, ps + q)) { diff = true; break; }
            if (!diff && byte_at(S, bed, ps + cl) != '\t') diff = true;
            L.flag = diff ? 1u : 0u;
        }
        S.st[threadIdx.x] = L.P.start; S.sp[threadIdx.x] = L.P.stop;
    }
    __syncthreads();
    if (act) {
        int64_t a, b;
        if (threadIdx.x == 0) { a = S.carry_start; b = S.carry_stop; }
        else { a = S.st[threadIdx.x - 1]; b = S.sp[threadIdx.x - 1]; }
        if (!L.flag) { L.pstop = b; L.plen = (int64_t)((uint64_t)b - (uint64_t)a); }      // hpp:523-532 resets both at a chromosome start
        int64_t len = (int64_t)((uint64_t)L.P.stop - (uint64_t)L.P.start), d = (int64_t)((uint64_t)L.P.start - (uint64_t)L.pstop);
        uint32_t rem_len = (uint32_t)(L.e - L.s) - L.P.rem_off;
        uint32_t o = (uint32_t)dec_len(d) + 1 + (rem_len ? rem_len + 1 : 0);
        if (len != L.plen) o += 2 + (uint32_t)dec_len(len);
        L.out_len = (halo && L.first_global) ? 0 : o;     // the halo line hands over its stop, length and chromosome only
    }
    return act;
}
// after the scans of a round (they synchronise): the round's last line becomes the carry
__device__ __forceinline__ void round_carry(TileSh &S, uint32_t base, const LineOut &L)
{
    if (threadIdx.x == FTH - 1 && base + FTH - 1 < S.k) { S.carry_start = L.P.start; S.carry_stop = L.P.stop; }
}

// ---- pass 1 -------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_agg(FAgg *dst, const FAgg &a)
{
    volatile uint64_t *d = reinterpret_cast<volatile uint64_t *>(dst);
    d[0] = a.lines; d[1] = a.out; d[2] = a.chroms; d[3] = (uint64_t)a.v; d[4] = (uint64_t)a.seg;
}
__device__ __forceinline__ FAgg load_agg(const FAgg *src)
{
    const volatile uint64_t *s = reinterpret_cast<const volatile uint64_t *>(src);
    FAgg a;
    a.lines = s[0]; a.out = s[1]; a.chroms = s[2]; a.v = (int64_t)s[3]; a.seg = (uint32_t)s[4]; a.pad = 0;
    return a;
}
__device__ __forceinline__ FAgg shfl_down_agg(const FAgg &a, int d)
{
    FAgg r;
    r.lines = __shfl_down_sync(0xffffffffu, a.lines, d); r.out = __shfl_down_sync(0xffffffffu, a.out, d);
    r.chroms = __shfl_down_sync(0xffffffffu, a.chroms, d); r.v = __shfl_down_sync(0xffffffffu, a.v, d);
    r.seg = __shfl_down_sync(0xffffffffu, a.seg, d); r.pad = 0;
    return r;
}

// status[tile]: 0 = nothing yet, 1 = agg[tile] valid, 2 = inc[tile] (inclusive prefix) valid
__global__ void __launch_bounds__(FTH) k_front_measure(const uint8_t *__restrict__ bed, uint64_t n, uint32_t skip, uint32_t halo, uint32_t ntiles,
                                                       uint32_t *status, FAgg *agg, FAgg *inc, unsigned long long *sc)
{
    __shared__ TileSh S;
    __shared__ uint32_t s_ticket;
    // tiles are handed out in launch order, so the tiles a look-back waits for are always running or done
    if (threadIdx.x == 0) s_ticket = (uint32_t)atomicAdd(&sc[SC_TICKET], 1ull);
    __syncthreads();
    const uint32_t tile = s_ticket;
    tile_setup(S, bed, n, skip, tile);
    const uint32_t k = S.k;
    FAgg A = fagg_identity();
    A.lines = k;
    uint32_t malformed = 0;
    if (k > FMAXL) malformed = 1;                          // lines shorter than three bytes
    else {
        for (uint32_t base = 0; base < k; base += FTH) {
            LineOut L;
            bool act = line_of_round(S, bed, base, halo, L);
            uint32_t tot_out, tot_ch;
            block_excl_sum<uint32_t>(L.out_len, S.sum_a, &tot_out);
            block_excl_sum<uint32_t>(L.flag, S.sum_b, &tot_ch);
            SegMax mine = SegMaxF::identity();
            if (act) { mine.v = (halo && L.first_global) ? INT64_MIN : L.P.stop; mine.seg = (int32_t)L.flag; }
            SegMax tot_seg;
            block_scan_partials<SegMaxF, FTH>(mine, S.seg_sm, &tot_seg);
            malformed |= (uint32_t)__syncthreads_or(act && L.P.malformed);
            // the flag of the input's second line (does the range continue the halo line's chromosome?)
            if (act) {
                uint32_t j = base + threadIdx.x;
                if ((j == 1 && !S.has_prev) || (j == 0 && S.has_prev && !S.has_prev2)) sc[SC_LINE1] = L.flag;
            }
            FAgg R; R.lines = 0; R.out = tot_out; R.chroms = tot_ch; R.v = tot_seg.v; R.seg = (uint32_t)tot_seg.seg; R.pad = 0;
            A = fagg_op(A, R);
            round_carry(S, base, L);
        }
    }
    if (threadIdx.x == 0) {
        if (malformed) atomicAdd(&sc[SC_MALFORMED], 1ull);
        if (k) atomicMax(&sc[SC_LASTNL], (unsigned long long)(S.tile0 + S.nl[(k <= FMAXL ? k : FMAXL) - 1] + 1));
    }
    if (k > FMAXL && threadIdx.x == 0) {
        // the compact list is cut short: the true last newline of the tile comes from the mask
        uint64_t last = 0;
        for (int w = FT / 32 - 1; w >= 0; w--) {
            uint32_t m = S.nlw[(tile ? FBACK : 0) / 32 + w];
            if (m) { last = S.tile0 + (uint64_t)w * 32 + (31 - __clz((int)m)) + 1; break; }
        }
        atomicMax(&sc[SC_LASTNL], (unsigned long long)last);
    }
    // ---- decoupled look-back (warp 0) ----
    if (threadIdx.x < 32) {
        const unsigned l = threadIdx.x;
        FAgg excl = fagg_identity();
        if (tile > 0) {
            if (l == 0) { store_agg(&agg[tile], A); __threadfence(); *reinterpret_cast<volatile uint32_t *>(&status[tile]) = 1u; }
            int64_t basei = (int64_t)tile - 1;
            long long t_begin = clock64();
            bool failed = false;
            while (true) {
                int64_t idx = basei - (int64_t)l;
                uint32_t f = idx >= 0 ? *reinterpret_cast<volatile uint32_t *>(&status[idx]) : 2u;   // before tile 0: an empty prefix
                unsigned m2 = __ballot_sync(0xffffffffu, f == 2u), m0 = __ballot_sync(0xffffffffu, f == 0u);
                int first2 = m2 ? __ffs((int)m2) - 1 : 32;
                unsigned need = first2 >= 31 ? 0xffffffffu : ((2u << first2) - 1u);
                if (m0 & need) {
                    // an earlier tile has not published yet: poll again (bounded: a lost tile must not hang the device)
                    bool give_up = clock64() - t_begin > 4000000000ll || *reinterpret_cast<volatile unsigned long long *>(&sc[SC_ERROR]) != 0;
                    if (__any_sync(0xffffffffu, give_up)) { failed = true; break; }
                    continue;
                }
                __threadfence();
                FAgg p = fagg_identity();
                if (idx >= 0) {
                    if ((int)l < first2) p = load_agg(&agg[idx]);
                    else if ((int)l == first2) p = load_agg(&inc[idx]);
                }
                // lane l holds tile basei - l: higher lanes are earlier tiles, so they are the left operand
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    FAgg q = shfl_down_agg(p, d);
                    if (l + d < 32) p = fagg_op(q, p);
                }
                p.lines = __shfl_sync(0xffffffffu, p.lines, 0); p.out = __shfl_sync(0xffffffffu, p.out, 0);
                p.chroms = __shfl_sync(0xffffffffu, p.chroms, 0); p.v = __shfl_sync(0xffffffffu, p.v, 0);
                p.seg = __shfl_sync(0xffffffffu, p.seg, 0);
                excl = fagg_op(p, excl);
                if (first2 < 32) break;
                basei -= 32;
            }
            if (failed && l == 0) sc[SC_ERROR] = 1ull;
        }
        if (l == 0) {
            FAgg I = fagg_op(excl, A);
            store_agg(&inc[tile], I);
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(&status[tile]) = 2u;
            if (tile == ntiles - 1) {
                sc[SC_TOTAL + 0] = I.lines; sc[SC_TOTAL + 1] = I.out; sc[SC_TOTAL + 2] = I.chroms;
                sc[SC_TOTAL + 3] = (unsigned long long)I.v; sc[SC_TOTAL + 4] = I.seg;
            }
        }
    }
}

// ---- pass 2 -------------------------------------------------------------------------------------------
struct ChromSeed {
    uint64_t first_line, name_off, tf_off;
    uint32_t name_len, pad;
};
struct DumpArrays {            // the per-line arrays of the s3g_tokenize entry point (null: not wanted)
    uint64_t *line_start;
    int64_t *start, *stop;
    uint32_t *rem_off;
    uint8_t *flags;
};

template <bool DUMP>
__global__ void __launch_bounds__(FTH) k_front_write(const uint8_t *__restrict__ bed, uint64_t n, uint32_t skip, uint32_t halo, int64_t carry_max,
                                                     const FAgg *__restrict__ inc, uint64_t n_lines, uint8_t *__restrict__ tf,
                                                     ChromSeed *seeds, unsigned long long *stat_len, unsigned long long *stat_uniq, DumpArrays da)
{
    __shared__ TileSh S;
    __shared__ __align__(16) uint8_t s_out[FOBUF + 16];
    __shared__ unsigned long long s_red[2][FTH / 32];
    const uint32_t tile = blockIdx.x;
    tile_setup(S, bed, n, skip, tile);
    const uint32_t k = S.k;
    if (k == 0 || k > FMAXL) return;
    const FAgg ex0 = tile ? inc[tile - 1] : fagg_identity();
    const uint64_t o_begin = ex0.out, o_end = inc[tile].out;
    const bool staged = o_end - o_begin <= FOBUF;
    const uint32_t ph = (uint32_t)o_begin & 15u;
    uint64_t run_out = 0, run_ch = 0;
    SegMax run_seg; run_seg.v = ex0.v; run_seg.seg = (int32_t)ex0.seg; run_seg.pad = 0;
    for (uint32_t base = 0; base < k; base += FTH) {
        LineOut L;
        bool act = line_of_round(S, bed, base, halo, L);
        uint32_t tot_out, tot_ch;
        uint32_t ex_out = block_excl_sum<uint32_t>(L.out_len, S.sum_a, &tot_out);
        uint32_t ex_ch = block_excl_sum<uint32_t>(L.flag, S.sum_b, &tot_ch);
        SegMax mine = SegMaxF::identity();
        if (act) { mine.v = (halo && L.first_global) ? INT64_MIN : L.P.stop; mine.seg = (int32_t)L.flag; }
        SegMax tot_seg;
        SegMax ex_seg = block_scan_partials<SegMaxF, FTH>(mine, S.seg_sm, &tot_seg);
        unsigned long long my_len = 0, my_uniq = 0;
        uint64_t chrom = 0;
        if (act) {
            const uint64_t g = ex0.lines + base + threadIdx.x;               // line index in the input
            const uint64_t o = o_begin + run_out + ex_out;
            chrom = ex0.chroms + run_ch + ex_ch + L.flag - 1;
            const bool is_halo = halo && L.first_global;
            int64_t rm = L.flag ? INT64_MIN : SegMaxF::op(run_seg, ex_seg).v;  // largest stop of the earlier lines of this chromosome
            if (halo && chrom == 0 && carry_max > rm) rm = carry_max;          // ... including those on other GPUs
            const int64_t s = L.P.start, t = L.P.stop;
            const int64_t len = (int64_t)((uint64_t)t - (uint64_t)s), d = (int64_t)((uint64_t)s - (uint64_t)L.pstop);
            if (!is_halo) {
                int64_t lo = s > rm ? s : rm;
                my_len = (unsigned long long)len;
                my_uniq = t > lo ? (unsigned long long)(t - lo) : 0ull;
                uint8_t *w = staged ? s_out + ph + (uint32_t)(o - o_begin) : tf + o;
                if (len != L.plen) { *w++ = 'p'; w += put_dec(w, len); *w++ = '\n'; }      // hpp:438-455
                w += put_dec(w, d);                                                        // hpp:456-500
                uint64_t r = L.s + L.P.rem_off;
                if (r < L.e) {
                    *w++ = '\t';
                    const uint8_t *src = L.s >= S.lo ? S.buf + (r - S.lo) : bed + r;
                    const uint32_t rl = (uint32_t)(L.e - r);
                    for (uint32_t q = 0; q < rl; q++) w[q] = src[q];
                    w += rl;
                }
                *w = '\n';
            }
            if (L.flag) {
                ChromSeed cs; cs.first_line = g; cs.name_off = L.s; cs.tf_off = o; cs.name_len = L.P.name_len; cs.pad = 0;
                seeds[chrom] = cs;
            }
            if (DUMP) {
                da.line_start[g] = L.s;
                if (g + 1 == n_lines) da.line_start[g + 1] = L.e + 1;
                da.start[g] = s; da.stop[g] = t; da.rem_off[g] = L.P.rem_off; da.flags[g] = (uint8_t)(L.flag | (L.P.malformed << 1));
            }
        }
        // per-chromosome sums: one atomic per tile round unless a chromosome starts inside it
        if (tot_ch == 0) {
            unsigned long long a = my_len, b = my_uniq;
#pragma unroll
            for (int dd = 16; dd; dd >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, dd); b += __shfl_xor_sync(0xffffffffu, b, dd); }
            if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = a; s_red[1][threadIdx.x >> 5] = b; }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long ta = 0, tb = 0;
                for (int w = 0; w < FTH / 32; w++) { ta += s_red[0][w]; tb += s_red[1][w]; }
                const uint64_t c0 = ex0.chroms + run_ch - 1;      // every line of the round belongs to it
                if (ta) atomicAdd(&stat_len[c0], ta);
                if (tb) atomicAdd(&stat_uniq[c0], tb);
            }
        } else if (act) {
            if (my_len) atomicAdd(&stat_len[chrom], my_len);
            if (my_uniq) atomicAdd(&stat_uniq[chrom], my_uniq);
        }
        run_out += tot_out; run_ch += tot_ch; run_seg = SegMaxF::op(run_seg, tot_seg);
        round_carry(S, base, L);
    }
    if (!staged) return;
    __syncthreads();
    uint8_t *dst = tf + (o_begin - ph);                                       // 16-byte aligned (tf comes from cudaMalloc)
    const uint32_t lo_b = ph, hi_b = ph + (uint32_t)(o_end - o_begin);
    for (uint32_t c = threadIdx.x * 16; c < hi_b; c += FTH * 16) {
        if (c >= lo_b && c + 16 <= hi_b) *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(s_out + c);
        else for (uint32_t j = c > lo_b ? c : lo_b; j < c + 16 && j < hi_b; j++) dst[j] = s_out[j];
    }
}

__global__ void k_chrom_finish(const ChromSeed *seeds, const unsigned long long *stat_len, const unsigned long long *stat_uniq,
                               uint64_t n_chroms, uint64_t n_lines, uint64_t tf_total, s3g_chrom *out, uint32_t halo)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chroms) return;
    ChromSeed a = seeds[c];
    s3g_chrom r;
    r.name_off = a.name_off; r.name_len = a.name_len; r.n_blocks = 0;
    r.tf_off = a.tf_off;
    r.tf_len = (c + 1 < n_chroms ? seeds[c + 1].tf_off : tf_total) - a.tf_off;
    r.line_count = (int64_t)((c + 1 < n_chroms ? seeds[c + 1].first_line : n_lines) - a.first_line) - (halo && c == 0 ? 1 : 0);
    r.bases_nonunique = (int64_t)stat_len[c];
    r.bases_unique = (int64_t)stat_uniq[c];
    r.bz_off = 0; r.bz_len = 0;
    out[c] = r;
}

// kernel (1): measures the range d_bed[0, n).  `halo` = 1: the range's first line is the last line BEFORE it (multi-GPU
// ranges, shard.cu): it hands its stop, length and chromosome to the second line and is itself neither written nor counted.
int run_tokenize(Ctx *ctx, const uint8_t *d_bed, uint64_t n, uint32_t skip, TfResult *out, uint32_t halo)
{
    *out = TfResult();
    uint64_t ntiles = (n + FT - 1) / FT;
    if (ntiles == 0) ntiles = 1;
    if (ntiles > 0x7fffffffull) { set_error("input too large"); return S3G_E_LIMIT; }
    S3G_TRY(ctx->tile_cnt.ensure(ntiles * 4));
    S3G_TRY(ctx->scan_a.ensure(ntiles * sizeof(FAgg)));
    S3G_TRY(ctx->scan_b.ensure(ntiles * sizeof(FAgg)));
    S3G_TRY(ctx->scalars.ensure(64 * 8));
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    S3G_CUDA(cudaMemsetAsync(d_sc, 0, 64 * 8, ctx->stream));
    S3G_CUDA(cudaMemsetAsync(ctx->tile_cnt.p, 0, ntiles * 4, ctx->stream));
    S3G_BYTES(ctx, n);
    S3G_LAUNCH(ctx, k_front_measure, (unsigned)ntiles, FTH, 0, d_bed, n, skip, halo, (uint32_t)ntiles, ctx->tile_cnt.as<uint32_t>(),
               ctx->scan_a.as<FAgg>(), ctx->scan_b.as<FAgg>(), (unsigned long long *)d_sc);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_sc, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    S3G_TRY(check_launch("front measure"));
    const uint64_t *h = ctx->h_scalars;
    if (h[SC_ERROR]) { set_error("tokenizer look-back timed out"); return S3G_E_CUDA; }
    out->n_lines = h[SC_TOTAL + 0];
    out->tf_len = h[SC_TOTAL + 1];
    out->n_chroms = h[SC_TOTAL + 2];
    out->dropped = out->n_lines ? n - h[SC_LASTNL] : n - skip;      // unterminated tail (hpp:181-190); the skip bytes are not ours
    ctx->front_halo = halo; ctx->front_skip = skip;
    ctx->front_tail_max = (int64_t)h[SC_TOTAL + 3];
    ctx->front_line1_flag = (uint32_t)h[SC_LINE1];
    if (out->n_lines && h[SC_MALFORMED]) {
        set_error("malformed BED line(s): fewer than three fields (in %llu 8 KiB tile(s) of the input)", (unsigned long long)h[SC_MALFORMED]);
        return S3G_E_MALFORMED;
    }
    return S3G_OK;
}

// summary of the measured range for the hand-over between GPUs (run_tokenize already has it on the host)
int run_range_summary(Ctx *ctx, uint64_t n_lines, uint32_t halo, int64_t *tail_max, uint64_t *last_flag, uint32_t *continues)
{
    *tail_max = INT64_MIN; *last_flag = 0; *continues = 0;
    if (n_lines == 0) return S3G_OK;
    *tail_max = ctx->front_tail_max;
    *last_flag = ctx->h_scalars[SC_TOTAL + 2] > 1 ? 1 : 0;         // non-zero: a chromosome starts after the first line
    *continues = (halo && n_lines > 1 && !(ctx->front_line1_flag & 1)) ? 1u : 0u;
    return S3G_OK;
}

// kernel (2) over the range measured by run_tokenize; dump = also leave the per-line arrays in ctx
int run_transform_rest(Ctx *ctx, const uint8_t *d_bed, uint64_t n, TfResult *out, uint32_t halo, int64_t carry_max, bool dump)
{
    const uint64_t n_lines = out->n_lines, n_chroms = out->n_chroms;
    uint64_t ntiles = (n + FT - 1) / FT;
    if (ntiles == 0) ntiles = 1;
    S3G_TRY(ctx->tf.ensure(out->tf_len + 64));
    S3G_TRY(ctx->chrom_first.ensure((n_chroms + 1) * sizeof(ChromSeed)));
    S3G_TRY(ctx->stat_b.ensure((n_chroms + 1) * 16));
    S3G_TRY(ctx->chroms.ensure((n_chroms + 1) * sizeof(s3g_chrom)));
    unsigned long long *stat_len = ctx->stat_b.as<unsigned long long>(), *stat_uniq = stat_len + (n_chroms + 1);
    S3G_CUDA(cudaMemsetAsync(stat_len, 0, (n_chroms + 1) * 16, ctx->stream));
    DumpArrays da = {nullptr, nullptr, nullptr, nullptr, nullptr};
    S3G_BYTES(ctx, n + out->tf_len);
    if (dump) {
        S3G_TRY(ctx->line_start.ensure((n_lines + 1) * 8));
        S3G_TRY(ctx->start.ensure(n_lines * 8));
        S3G_TRY(ctx->stop.ensure(n_lines * 8));
        S3G_TRY(ctx->rem_off.ensure(n_lines * 4));
        S3G_TRY(ctx->flags.ensure(n_lines));
        da.line_start = ctx->line_start.as<uint64_t>(); da.start = ctx->start.as<int64_t>(); da.stop = ctx->stop.as<int64_t>();
        da.rem_off = ctx->rem_off.as<uint32_t>(); da.flags = ctx->flags.as<uint8_t>();
        S3G_LAUNCH(ctx, k_front_write<true>, (unsigned)ntiles, FTH, 0, d_bed, n, ctx->front_skip, halo, carry_max, ctx->scan_b.as<FAgg>(), n_lines,
                   ctx->tf.as<uint8_t>(), ctx->chrom_first.as<ChromSeed>(), stat_len, stat_uniq, da);
    } else {
        S3G_LAUNCH(ctx, k_front_write<false>, (unsigned)ntiles, FTH, 0, d_bed, n, ctx->front_skip, halo, carry_max, ctx->scan_b.as<FAgg>(), n_lines,
                   ctx->tf.as<uint8_t>(), ctx->chrom_first.as<ChromSeed>(), stat_len, stat_uniq, da);
    }
    S3G_LAUNCH(ctx, k_chrom_finish, (unsigned)((n_chroms + 127) / 128), 128, 0, ctx->chrom_first.as<ChromSeed>(), stat_len, stat_uniq,
               n_chroms, n_lines, out->tf_len, ctx->chroms.as<s3g_chrom>(), halo);
    return check_launch("transform write");
}

int run_transform(Ctx *ctx, const uint8_t *d_bed, uint64_t n, TfResult *out, bool tokenize_only, uint32_t skip)
{
    S3G_TRY(run_tokenize(ctx, d_bed, n, skip, out, 0));
    if (out->n_lines == 0) return S3G_OK;
    S3G_TRY(run_transform_rest(ctx, d_bed, n, out, 0, INT64_MIN, tokenize_only));
    if (tokenize_only) {
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        return check_launch("tokenize");
    }
    return S3G_OK;
}

}  // namespace s3g
