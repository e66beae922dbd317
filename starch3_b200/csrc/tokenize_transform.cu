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
// (unique bases).  Nothing per line is kept in HBM: the input is read twice and only the transformed bytes
// are written.  The unit of work is a CHUNK of 2 KiB handled by ONE WARP, which owns the lines that END in
// the chunk; warps never wait for each other inside a pass (no block barrier in the line loops).
//
//   k_front_measure  lane l loads bytes [64 l, 64 l + 64) of the chunk as four 16-byte vectors and turns them
//                    into a newline and a tab bit mask; the newline offsets become the warp's line list, the tab
//                    masks (plus those of the 256 bytes before the chunk) let a lane find the fields of its line
//                    without touching the bytes.  32 lines per round, a lane per line: integer parse,
//                    chromosome test against the line before, output length; warp reductions give the chunk's
//                    {lines, output bytes, chromosome starts, running max of stop}.  The eight chunks of a CTA
//                    are combined once at its end.
//   k_front_scan     exclusive scan of the per-tile aggregates (one CTA; 40 bytes per 16 KiB of input).
//   k_front_write    the same walk with the prefix known: formats the transformed lines into the warp's strip of
//                    shared memory at the destination's 16-byte phase and stores them as aligned vectors;
//                    chromosome table seeds, per-chromosome sums (one atomic pair per chunk, spread over 32
//                    slots), optionally the per-line arrays (the s3g_tokenize parity entry point).
//   k_chrom_finish   chromosome table from the seeds and sums.
#include "common.cuh"
#include "scan.cuh"

namespace s3g {

constexpr int FCH = 2048;                // input bytes per chunk (one warp)
constexpr int FWARPS = 8;                // chunks per CTA ("tile": 16 KiB)
constexpr int FTH = FWARPS * 32;
constexpr int FBACK = 256;               // bytes before the chunk whose tabs and newlines are in the masks too
constexpr int FMAXL = FCH / 3 + 2;       // a line with three fields has at least 3 bytes ("\t\t\n")
constexpr int FOB = FCH + 512;           // staged output bytes of a chunk (more: direct stores)

// scalar slots (ctx->scalars, u64 each)
enum { SC_MALFORMED = 1, SC_LASTNL = 2, SC_LINE1 = 3, SC_UNSORTED = 4, SC_CRLF = 5, SC_TOTAL = 8 /* FAgg: 5 slots */ };

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
        const uint32_t t = w[q] ^ pat;                                             // zero byte <=> equal
        const uint32_t z = ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu);  // 0x80 in every zero byte, exact
        m |= (((z >> 7) * 0x01020408u) >> 24) << (4 * q);                          // bits 0, 8, 16, 24 -> bits 0..3
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

// ---- one chunk as its warp sees it ---------------------------------------------------------------------
struct WarpSh {
    uint64_t tb[FBACK / 64 + FCH / 64];      // tab masks: entry q covers bytes [chunk0 - FBACK + 64 q, + 64)
    uint16_t nl[FMAXL + 2];                  // newlines inside the chunk, offsets from chunk0, ascending
};
struct ChunkGeom {                           // the same in every lane
    uint64_t chunk0, start0, prev_start;     // start0: where the chunk's first line starts; prev_start: the line before it
    int64_t mask_lo;                         // position of bit 0 of tb[0]
    uint64_t n;                              // bytes of the buffer
    uint32_t k, has_prev, has_prev2;         // k: lines that end in the chunk
};

struct Parsed {
    int64_t start, stop;
    uint32_t rem_off, name_len, malformed;
};

// eight bytes starting at position p (little endian: the byte at p is the lowest); needs [p & ~7, (p & ~7) + 16) readable
__device__ __forceinline__ uint64_t load8(const uint8_t *__restrict__ bed, uint64_t p)
{
    const uint64_t *q = reinterpret_cast<const uint64_t *>(bed + (p & ~7ull));
    const uint64_t lo = q[0], hi = q[1];
    const uint32_t sh = (uint32_t)(p & 7) * 8;
    return sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
}
// the decimal number in the `len` (1..8) HIGH bytes of x (string order = memory order), or false if one of them is not a digit
__device__ __forceinline__ bool swar_digits(uint64_t x, uint32_t len, uint64_t *val)
{
    const uint32_t drop = 8 * (8 - len);
    x = drop ? ((x >> drop) << drop) | (0x3030303030303030ull >> (64 - drop)) : x;      // the bytes before the number become '0'
    const uint64_t d = x - 0x3030303030303030ull;
    if (((d + 0x7676767676767676ull) | d | x) & 0x8080808080808080ull) return false;    // a byte outside '0'..'9'
    uint64_t v = (d * 2561ull) >> 8;                                                     // pairs: 10 a + b
    v = ((v & 0x00ff00ff00ff00ffull) * 6553601ull) >> 16;                                // 100 ab + cd
    v = ((v & 0x0000ffff0000ffffull) * 42949672960001ull) >> 32;                         // 10000 abcd + efgh
    *val = v;
    return true;
}
// sscanf("%lld") over the documented domain (hpp:306-307): [sign] digits, up to the first other byte.  The field is
// bed[a, b); n = bytes of the buffer.  Fields of up to 16 digits are converted eight bytes at a time.
__device__ __forceinline__ int64_t parse_int(const uint8_t *__restrict__ bed, uint64_t n, uint64_t a, uint64_t b)
{
    if (a >= b) return 0;
    int neg = 0;
    const uint8_t c0 = bed[a];
    if (c0 == '-' || c0 == '+') { neg = c0 == '-'; a++; }
    const uint32_t len = (uint32_t)(b - a);
    if (len >= 1 && len <= 16 && b >= 16 && ((b + 7) & ~7ull) + 8 <= n) {
        uint64_t lo, hi = 0;
        const uint32_t l8 = len > 8 ? 8 : len;
        bool ok = swar_digits(load8(bed, b - 8), l8, &lo);
        if (ok && len > 8) ok = swar_digits(load8(bed, b - 16), len - 8, &hi);
        if (ok) {
            const uint64_t acc = hi * 100000000ull + lo;
            return neg ? (int64_t)(0 - acc) : (int64_t)acc;
        }
    }
    uint64_t acc = 0;
    for (uint64_t i = a; i < b; i++) {
        uint32_t d = (uint32_t)bed[i] - '0';
        if (d > 9u) break;
        acc = acc * 10 + d;
    }
    return neg ? (int64_t)(0 - acc) : (int64_t)acc;
}

// first tab in [p, e), or e; p at or after the first masked byte
__device__ __forceinline__ uint64_t next_tab(const WarpSh &W, int64_t mask_lo, uint64_t p, uint64_t e)
{
    if (p >= e) return e;
    uint32_t i = (uint32_t)((int64_t)p - mask_lo), iend = (uint32_t)((int64_t)e - mask_lo);
    uint32_t w = i >> 6;
    uint64_t m = W.tb[w] & (~0ull << (i & 63));
    while (m == 0) {
        w++;
        if ((w << 6) >= iend) return e;
        m = W.tb[w];
    }
    uint32_t q = (w << 6) + (uint32_t)__ffsll((long long)m) - 1;
    return q < iend ? (uint64_t)(mask_lo + q) : e;
}

// line [s, e), bed[e] == '\n'; fields as consume_line finds them (hpp:220-309)
__device__ __forceinline__ Parsed parse_line(const WarpSh &W, const ChunkGeom &G, const uint8_t *__restrict__ bed, uint64_t s, uint64_t e)
{
    Parsed P;
    P.start = 0; P.stop = 0; P.malformed = 0;
    uint64_t t1, t2, t3;
    if ((int64_t)s >= G.mask_lo) {                     // tabs from the bit masks
        t1 = next_tab(W, G.mask_lo, s, e);
        t2 = t1 < e ? next_tab(W, G.mask_lo, t1 + 1, e) : e;
        t3 = t2 < e ? next_tab(W, G.mask_lo, t2 + 1, e) : e;
    } else {                                           // a line that starts before the masked bytes
        uint64_t q = s;
        while (q < e && bed[q] != '\t') q++;
        t1 = q;
        if (q < e) { q++; while (q < e && bed[q] != '\t') q++; }
        t2 = q;
        if (q < e) { q++; while (q < e && bed[q] != '\t') q++; }
        t3 = q;
    }
    P.name_len = (uint32_t)(t1 - s);
    if (t2 >= e) {                                     // fewer than three fields
        P.malformed = 1;
        P.rem_off = (uint32_t)(e - s);
        return P;
    }
    P.start = parse_int(bed, G.n, t1 + 1, t2);
    P.stop = parse_int(bed, G.n, t2 + 1, t3);
    P.rem_off = (uint32_t)((t3 < e ? t3 + 1 : e) - s);
    return P;
}

// largest p in [floor, hi) with bed[p] == '\n', or -1; one warp
__device__ __forceinline__ int64_t scan_back(const uint8_t *__restrict__ bed, uint64_t hi, uint64_t floor)
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

// 64 bytes -> newline and tab masks (bit i <=> byte i)
__device__ __forceinline__ void masks64(const uint8_t *__restrict__ bed, uint64_t n, uint64_t pos, uint64_t *nlm, uint64_t *tbm)
{
    uint64_t mn = 0, mt = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint64_t p = pos + 16ull * q;
        uint32_t w[4] = {0, 0, 0, 0};
        if (p + 16 <= n) {
            uint4 v = *reinterpret_cast<const uint4 *>(bed + p);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else if (p < n) {
            for (uint32_t k = 0; k < (uint32_t)(n - p); k++) {
                uint32_t by = (uint32_t)bed[p + k] << (8 * (k & 3));
                if ((k >> 2) == 0) w[0] |= by; else if ((k >> 2) == 1) w[1] |= by; else if ((k >> 2) == 2) w[2] |= by; else w[3] |= by;
            }
        }
        mn |= (uint64_t)eq_mask16(w, 0x0a0a0a0au) << (16 * q);
        mt |= (uint64_t)eq_mask16(w, 0x09090909u) << (16 * q);
    }
    *nlm = mn; *tbm = mt;
}

// masks, line list, and where the chunk's first line and the line before it start; all lanes of the warp
__device__ __forceinline__ void chunk_setup(WarpSh &W, ChunkGeom &G, const uint8_t *__restrict__ bed, uint64_t n, uint32_t skip, uint64_t chunk)
{
    const unsigned l = threadIdx.x & 31;
    const uint64_t chunk0 = chunk * FCH;
    G.chunk0 = chunk0; G.n = n;
    G.mask_lo = (int64_t)chunk0 - FBACK;
    G.k = 0; G.has_prev = 0; G.has_prev2 = 0; G.start0 = skip; G.prev_start = skip;
    if (chunk0 >= n) return;
    uint64_t mn, mt;
    masks64(bed, n, chunk0 + 64ull * l, &mn, &mt);
    if (chunk == 0 && 64u * l < skip) mn &= ~0ull << skip;                 // the bytes before the range: not ours (skip < 16)
    W.tb[FBACK / 64 + l] = mt;
    int64_t nl1 = -1, nl2 = -1;
    if (chunk) {
        // the 256 bytes before the chunk: tab masks for the lines that start there, and the last two newlines
        uint64_t bn = 0, bt = 0;
        if (l < FBACK / 64) { masks64(bed, n, chunk0 - FBACK + 64ull * l, &bn, &bt); W.tb[l] = bt; }
        unsigned nz = __ballot_sync(0xffffffffu, bn != 0);
        if (nz) {
            int hl = 31 - __clz((int)nz);
            uint64_t hw = __shfl_sync(0xffffffffu, bn, hl);
            int hb = 63 - __clzll((long long)hw);
            nl1 = (int64_t)(chunk0 - FBACK + 64ull * hl + hb);
            uint64_t hw2 = hw & ~(1ull << hb);
            unsigned nz2 = nz & ~(1u << hl);
            if (hw2) nl2 = (int64_t)(chunk0 - FBACK + 64ull * hl + (63 - __clzll((long long)hw2)));
            else if (nz2) {
                int hl2 = 31 - __clz((int)nz2);
                uint64_t w2 = __shfl_sync(0xffffffffu, bn, hl2);
                nl2 = (int64_t)(chunk0 - FBACK + 64ull * hl2 + (63 - __clzll((long long)w2)));
            }
            if (nl2 < 0) nl2 = scan_back(bed, chunk0 - FBACK, skip);
        } else {
            nl1 = scan_back(bed, chunk0 - FBACK, skip);
            if (nl1 >= 0) nl2 = scan_back(bed, (uint64_t)nl1, skip);
        }
    }
    // the chunk's newlines -> compact list
    uint32_t c = (uint32_t)__popcll((long long)mn);
    uint32_t inc = warp_incl_sum<uint32_t>(c);
    uint32_t ex = inc - c;
    G.k = __shfl_sync(0xffffffffu, inc, 31);
    while (mn) {
        int b = __ffsll((long long)mn) - 1;
        mn &= mn - 1;
        if (ex < FMAXL) W.nl[ex] = (uint16_t)(64 * l + b);
        ex++;
    }
    G.has_prev = nl1 >= 0; G.has_prev2 = nl2 >= 0;
    G.start0 = nl1 >= 0 ? (uint64_t)nl1 + 1 : skip;
    G.prev_start = nl2 >= 0 ? (uint64_t)nl2 + 1 : skip;
    __syncwarp();
}

// what a lane holds in a round.  Item v = 32 * round + lane of a chunk: v = 0 is the line BEFORE the chunk's first line (it
// only hands its start, stop and name on), v = 1 .. k are the chunk's lines.
struct LineOut {
    uint64_t s, e;
    Parsed P;
    uint32_t owned;         // one of the chunk's lines
    uint32_t flag;          // chromosome differs from the previous line (hpp:331), or first line of the input
    uint32_t first_global;  // first line of the input
    uint32_t out_len;
    uint32_t unsorted;      // starts before the previous element of its chromosome (input hardening)
    int64_t pstop, plen;
};

// carry_*: start / stop of the last item of the round before (in/out); every lane of the warp calls
__device__ __forceinline__ void line_of_round(const WarpSh &W, const ChunkGeom &G, const uint8_t *__restrict__ bed, uint32_t round, uint32_t halo,
                                              int64_t &carry_start, int64_t &carry_stop, LineOut &L)
{
    const unsigned l = threadIdx.x & 31;
    const uint32_t v = round * 32 + l;
    const bool act = v <= G.k && (v >= 1 || G.has_prev);
    L.owned = act && v >= 1;
    L.flag = 0; L.first_global = 0; L.out_len = 0; L.pstop = 0; L.plen = 0; L.unsorted = 0;
    L.P.start = 0; L.P.stop = 0; L.P.rem_off = 0; L.P.name_len = 0; L.P.malformed = 0; L.s = 0; L.e = 0;
    if (act) {
        if (v == 0) { L.s = G.prev_start; L.e = G.start0 - 1; }
        else {
            L.s = v == 1 ? G.start0 : G.chunk0 + W.nl[v - 2] + 1;
            L.e = G.chunk0 + W.nl[v - 1];
        }
        L.P = parse_line(W, G, bed, L.s, L.e);
        if (v >= 1) {
            L.first_global = (v == 1 && !G.has_prev) ? 1u : 0u;
            if (L.first_global) L.flag = 1;
            else {
                // strcmp(chr, previous chr) != 0 (hpp:331); the previous line's field ends at its first tab
                const uint64_t ps = v == 1 ? G.prev_start : (v == 2 ? G.start0 : G.chunk0 + W.nl[v - 3] + 1);
                const uint32_t cl = L.P.name_len;
                bool diff = false;
                if (cl < 8 && ((L.s + 7) & ~7ull) + 16 <= G.n && !L.P.malformed) {
                    // name and the tab that ends it, both lines, eight bytes at once
                    const uint64_t x = load8(bed, L.s) ^ load8(bed, ps);
                    diff = (x << (8 * (7 - cl))) != 0;
                } else {
                    for (uint32_t q = 0; q < cl; q++)
                        if (bed[L.s + q] != bed[ps + q]) { diff = true; break; }
                    if (!diff && bed[ps + cl] != '\t') diff = true;
                }
                L.flag = diff ? 1u : 0u;
            }
        }
    }
    int64_t a = __shfl_up_sync(0xffffffffu, L.P.start, 1), b = __shfl_up_sync(0xffffffffu, L.P.stop, 1);
    if (l == 0) { a = carry_start; b = carry_stop; }
    carry_start = __shfl_sync(0xffffffffu, L.P.start, 31); carry_stop = __shfl_sync(0xffffffffu, L.P.stop, 31);
    if (L.owned) {
        if (!L.flag) { L.pstop = b; L.plen = (int64_t)((uint64_t)b - (uint64_t)a); L.unsorted = L.P.start < a; }   // hpp:523-532 resets both at a chromosome start
        int64_t len = (int64_t)((uint64_t)L.P.stop - (uint64_t)L.P.start), d = (int64_t)((uint64_t)L.P.start - (uint64_t)L.pstop);
        uint32_t rem_len = (uint32_t)(L.e - L.s) - L.P.rem_off;
        uint32_t o = (uint32_t)dec_len(d) + 1 + (rem_len ? rem_len + 1 : 0);
        if (len != L.plen) o += 2 + (uint32_t)dec_len(len);
        L.out_len = (halo && L.first_global) ? 0 : o;     // the halo line hands over its stop, length and chromosome only
    }
}

// inclusive segmented max over the lanes (a set flag restarts the maximum); returns the lane's (v, f) with f = a flag at or before it
__device__ __forceinline__ void warp_segmax_incl(int64_t &v, uint32_t &f)
{
    const unsigned l = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t ov = __shfl_up_sync(0xffffffffu, v, d);
        uint32_t of = __shfl_up_sync(0xffffffffu, f, d);
        if (l >= (unsigned)d) { if (!f) { v = ov > v ? ov : v; f = of; } }
    }
}

// one chunk's aggregate: lines, output bytes, chromosome starts, running max of stop.  Every lane of the warp calls; A and
// malformed are the same in all lanes on return.
__device__ __forceinline__ void measure_chunk(const WarpSh &W, const ChunkGeom &G, const uint8_t *__restrict__ bed, uint32_t halo,
                                              unsigned long long *sc, FAgg &A, uint32_t &malformed)
{
    const unsigned l = threadIdx.x & 31;
    A = fagg_identity();
    A.lines = G.k;
    malformed = 0;
    if (G.k > FMAXL) malformed = 1;                        // lines shorter than three bytes
    else if (G.k) {
        int64_t cs = 0, ce = 0;
        const uint32_t rounds = (G.k + 1 + 31) / 32;
        for (uint32_t r = 0; r < rounds; r++) {
            LineOut L;
            line_of_round(W, G, bed, r, halo, cs, ce, L);
            uint32_t tot_out = __reduce_add_sync(0xffffffffu, L.out_len);
            unsigned mflag = __ballot_sync(0xffffffffu, L.flag != 0);
            malformed |= __any_sync(0xffffffffu, L.owned && L.P.malformed) ? 1u : 0u;
            int64_t v = L.owned ? ((halo && L.first_global) ? INT64_MIN : L.P.stop) : INT64_MIN;
            uint32_t f = L.flag;
            warp_segmax_incl(v, f);
            FAgg R; R.lines = 0; R.out = tot_out; R.chroms = __popc(mflag); R.pad = 0;
            R.v = __shfl_sync(0xffffffffu, v, 31); R.seg = mflag != 0;
            A = fagg_op(A, R);
            // the flag of the input's second line (does the range continue the halo line's chromosome?)
            const uint32_t vv = r * 32 + l;
            if (L.owned && ((vv == 2 && !G.has_prev) || (vv == 1 && G.has_prev && !G.has_prev2))) sc[SC_LINE1] = L.flag;
        }
    }
}

// where the chunk's last line ends (+1), 0 if no line ends in it; every lane of the warp calls
__device__ __forceinline__ unsigned long long chunk_last_newline(const WarpSh &W, const ChunkGeom &G, const uint8_t *__restrict__ bed, uint64_t n)
{
    const unsigned l = threadIdx.x & 31;
    if (G.k == 0) return 0;
    if (G.k <= FMAXL) return G.chunk0 + W.nl[G.k - 1] + 1;
    // the compact list is cut short: the chunk's last newline from the masks
    uint64_t mn, mt;
    masks64(bed, n, G.chunk0 + 64ull * l, &mn, &mt);
    unsigned nz = __ballot_sync(0xffffffffu, mn != 0);
    int hl = 31 - __clz((int)nz);
    uint64_t hw = __shfl_sync(0xffffffffu, mn, hl);
    return G.chunk0 + 64ull * hl + (63 - __clzll((long long)hw)) + 1;
}

// ---- pass 1 -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FTH) k_front_measure(const uint8_t *__restrict__ bed, uint64_t n, uint32_t skip, uint32_t halo,
                                                       FAgg *__restrict__ tile_agg, FAgg *__restrict__ chunk_pre, uint32_t *__restrict__ chunk_out,
                                                       unsigned long long *sc)
{
    __shared__ WarpSh Ws[FWARPS];
    __shared__ FAgg s_agg[FWARPS];
    __shared__ unsigned long long s_lastnl[FWARPS];
    __shared__ uint32_t s_mal[FWARPS];
    const unsigned l = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t chunk = (uint64_t)blockIdx.x * FWARPS + wid;
    WarpSh &W = Ws[wid];
    ChunkGeom G;
    chunk_setup(W, G, bed, n, skip, chunk);
    FAgg A;
    uint32_t malformed;
    measure_chunk(W, G, bed, halo, sc, A, malformed);
    const unsigned long long last = chunk_last_newline(W, G, bed, n);
    if (l == 0) {
        s_agg[wid] = A; s_mal[wid] = malformed;
        s_lastnl[wid] = last;
        chunk_out[chunk] = (uint32_t)A.out;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        FAgg run = fagg_identity();
        unsigned long long lastm = 0; uint32_t mal = 0;
        for (int w = 0; w < FWARPS; w++) {
            chunk_pre[(uint64_t)blockIdx.x * FWARPS + w] = run;
            run = fagg_op(run, s_agg[w]);
            if (s_lastnl[w] > lastm) lastm = s_lastnl[w];
            mal |= s_mal[w];
        }
        tile_agg[blockIdx.x] = run;
        if (mal) atomicAdd(&sc[SC_MALFORMED], 1ull);
        if (lastm) atomicMax(&sc[SC_LASTNL], lastm);
    }
}

struct FAggF {
    typedef FAgg T;
    __host__ __device__ static T identity() { return fagg_identity(); }
    __host__ __device__ static T op(const T &a, const T &b) { return fagg_op(a, b); }
};

// exclusive scan of the tile aggregates in two levels: spans of 1024 tiles (a thread per tile), then the span totals.
// tile t's prefix = span_pre[t / 1024] (+) agg[t]; totals -> sc[SC_TOTAL ..]
constexpr int FSCAN_T = 1024;
__global__ void __launch_bounds__(FSCAN_T) k_front_scan_tiles(FAgg *agg, uint64_t ntiles, FAgg *span_tot)
{
    __shared__ FAgg sm[FSCAN_T / 32];
    const uint64_t t = (uint64_t)blockIdx.x * FSCAN_T + threadIdx.x;
    FAgg a = t < ntiles ? agg[t] : fagg_identity();
    FAgg tot;
    FAgg ex = block_scan_partials<FAggF, FSCAN_T>(a, sm, &tot);
    if (t < ntiles) agg[t] = ex;
    if (threadIdx.x == 0) span_tot[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(FSCAN_T) k_front_scan_spans(FAgg *span, uint64_t nspans, unsigned long long *sc)
{
    __shared__ FAgg sm[FSCAN_T / 32];
    FAgg carry = fagg_identity();
    for (uint64_t base = 0; base < nspans; base += FSCAN_T) {
        const uint64_t i = base + threadIdx.x;
        FAgg a = i < nspans ? span[i] : fagg_identity();
        FAgg tot;
        FAgg ex = block_scan_partials<FAggF, FSCAN_T>(a, sm, &tot);
        if (i < nspans) span[i] = fagg_op(carry, ex);
        carry = fagg_op(carry, tot);
    }
    if (threadIdx.x == 0) {
        sc[SC_TOTAL + 0] = carry.lines; sc[SC_TOTAL + 1] = carry.out; sc[SC_TOTAL + 2] = carry.chroms;
        sc[SC_TOTAL + 3] = (unsigned long long)carry.v; sc[SC_TOTAL + 4] = carry.seg;
    }
}

// ---- pass 2 -------------------------------------------------------------------------------------------
// where the transformed bytes go: up to 8 buffers (this GPU's first), the range's bytes at offset `off` in each
struct PeerDst {
    uint8_t *ptr[8];
    uint8_t *mc;          // NVLS multicast address of the same buffers (a store to it lands in every GPU's copy), or null
    uint64_t off;
    uint32_t n;
};
// one 16-byte store that the NVSwitch replicates into every GPU's copy of the buffer
__device__ __forceinline__ void multimem_st16(uint8_t *mc_addr, const uint4 &v)
{
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc_addr), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}
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

// the transformed lines of one chunk, given what everything before it leaves (ex0) and its own output bytes (o_len); every
// lane of the warp calls.  obuf: the warp's strip of shared memory; slot_salt spreads a chromosome's sums over stat_slots.
template <bool DUMP>
__device__ __forceinline__ void write_chunk(const WarpSh &W, const ChunkGeom &G, const uint8_t *__restrict__ bed, uint32_t halo, int64_t carry_max,
                                            const FAgg &ex0, uint32_t o_len, uint64_t n_lines, const PeerDst &pd, uint8_t *obuf, uint32_t slot_salt,
                                            ChromSeed *seeds, unsigned long long *stat_len, unsigned long long *stat_uniq, uint32_t stat_slots,
                                            uint64_t diag_chroms, unsigned long long *sc, const DumpArrays &da)
{
    const unsigned l = threadIdx.x & 31;
    const uint64_t o_begin = ex0.out;
    const bool staged = o_len <= FOB;
    uint8_t *const tf = pd.ptr[0] + pd.off;                // destination 0 (this GPU's own buffer); pd.off keeps 16-byte phase in mind:
    const uint32_t ph = (uint32_t)(pd.off + o_begin) & 15u; // every destination buffer is 16-byte aligned, so the phase is the same in all
    uint64_t run_out = 0, run_ch = 0;
    int64_t run_v = ex0.v;                                 // largest stop of the current chromosome so far
    const uint64_t c0 = ex0.chroms - 1;                    // the chromosome the chunk begins in (unless its first line starts one)
    unsigned long long acc_len = 0, acc_uniq = 0;          // sums of the lines that belong to c0
    uint32_t n_unsorted = 0, n_crlf = 0;                   // diagnostics over the lines of chromosomes < diag_chroms
    int64_t cs = 0, ce = 0;
    const uint32_t rounds = (G.k + 1 + 31) / 32;
    for (uint32_t r = 0; r < rounds; r++) {
        LineOut L;
        line_of_round(W, G, bed, r, halo, cs, ce, L);
        const uint32_t inc_out = warp_incl_sum<uint32_t>(L.out_len);
        const uint32_t tot_out = __shfl_sync(0xffffffffu, inc_out, 31);
        const unsigned mflag = __ballot_sync(0xffffffffu, L.flag != 0);
        const uint32_t ex_ch = __popc(mflag & ((1u << l) - 1u));
        int64_t v = L.owned ? ((halo && L.first_global) ? INT64_MIN : L.P.stop) : INT64_MIN;
        uint32_t f = L.flag;
        warp_segmax_incl(v, f);
        // exclusive: what the lanes before this one leave
        int64_t pv = __shfl_up_sync(0xffffffffu, v, 1);
        uint32_t pf = __shfl_up_sync(0xffffffffu, f, 1);
        if (l == 0) { pv = INT64_MIN; pf = 0; }
        const int64_t tot_v = __shfl_sync(0xffffffffu, v, 31);
        if (L.owned) {
            const uint64_t g = ex0.lines + r * 32 + l - 1;                       // line index in the input
            const uint64_t o = o_begin + run_out + (inc_out - L.out_len);
            const uint64_t chrom = ex0.chroms + run_ch + ex_ch + L.flag - 1;
            const bool is_halo = halo && L.first_global;
            if (chrom < diag_chroms && !is_halo) {
                n_unsorted += L.unsorted;
                n_crlf += (L.e > L.s && bed[L.e - 1] == '\r') ? 1u : 0u;
            }
            int64_t rm = L.flag ? INT64_MIN : (pf ? pv : (run_v > pv ? run_v : pv));   // largest stop of the earlier lines of this chromosome
            if (halo && chrom == 0 && carry_max > rm) rm = carry_max;                  // ... including those on other GPUs
            const int64_t s = L.P.start, t = L.P.stop;
            const int64_t len = (int64_t)((uint64_t)t - (uint64_t)s), d = (int64_t)((uint64_t)s - (uint64_t)L.pstop);
            if (!is_halo) {
                const int64_t lo = s > rm ? s : rm;
                const unsigned long long my_len = (unsigned long long)len, my_uniq = t > lo ? (unsigned long long)(t - lo) : 0ull;
                if (chrom == c0) { acc_len += my_len; acc_uniq += my_uniq; }
                else {
                    const uint64_t slot = chrom * stat_slots + (slot_salt & (stat_slots - 1));
                    if (my_len) atomicAdd(&stat_len[slot], my_len);
                    if (my_uniq) atomicAdd(&stat_uniq[slot], my_uniq);
                }
                uint8_t *w = staged ? obuf + ph + (uint32_t)(o - o_begin) : tf + o;
                if (len != L.plen) { *w++ = 'p'; w += put_dec(w, len); *w++ = '\n'; }      // hpp:438-455
                w += put_dec(w, d);                                                        // hpp:456-500
                const uint64_t rr = L.s + L.P.rem_off;
                if (rr < L.e) {
                    *w++ = '\t';
                    const uint8_t *src = bed + rr;
                    const uint32_t rl = (uint32_t)(L.e - rr);
                    for (uint32_t q = 0; q < rl; q++) w[q] = src[q];
                    w += rl;
                }
                *w = '\n';
            }
            if (L.flag) {
                ChromSeed sd; sd.first_line = g; sd.name_off = L.s; sd.tf_off = o; sd.name_len = L.P.name_len; sd.pad = 0;
                seeds[chrom] = sd;
            }
            if (DUMP) {
                da.line_start[g] = L.s;
                if (g + 1 == n_lines) da.line_start[g + 1] = L.e + 1;
                da.start[g] = s; da.stop[g] = t; da.rem_off[g] = L.P.rem_off; da.flags[g] = (uint8_t)(L.flag | (L.P.malformed << 1));
            }
        }
        run_out += tot_out; run_ch += __popc(mflag);
        run_v = mflag ? tot_v : (run_v > tot_v ? run_v : tot_v);
    }
    n_unsorted = __reduce_add_sync(0xffffffffu, n_unsorted); n_crlf = __reduce_add_sync(0xffffffffu, n_crlf);
    if (l == 0 && n_unsorted) atomicAdd(&sc[SC_UNSORTED], (unsigned long long)n_unsorted);
    if (l == 0 && n_crlf) atomicAdd(&sc[SC_CRLF], (unsigned long long)n_crlf);
    // per-chromosome sums of the chunk's first chromosome: one atomic pair per chunk
#pragma unroll
    for (int dd = 16; dd; dd >>= 1) { acc_len += __shfl_xor_sync(0xffffffffu, acc_len, dd); acc_uniq += __shfl_xor_sync(0xffffffffu, acc_uniq, dd); }
    if (l == 0 && ex0.chroms) {
        const uint64_t slot = c0 * stat_slots + (slot_salt & (stat_slots - 1));
        if (acc_len) atomicAdd(&stat_len[slot], acc_len);
        if (acc_uniq) atomicAdd(&stat_uniq[slot], acc_uniq);
    }
    __syncwarp();
    if (!staged) {
        // (a chunk whose output outgrew the strip was written straight into destination 0) the other destinations get copies
        for (uint32_t k = 1; k < pd.n; k++) {
            uint8_t *d = pd.ptr[k] + pd.off + o_begin;
            for (uint32_t j = l; j < o_len; j += 32) d[j] = tf[o_begin + j];
        }
        return;
    }
    // the strip leaves as aligned 16-byte vectors -- into this GPU's buffer and, in the N-GPU path, into every peer's copy of
    // the transformed buffer through its NVLink-mapped pointer: the all-gather of the transformed bytes is these stores
    const uint32_t lo_b = ph, hi_b = ph + o_len;
    if (pd.mc) {
        // whole vectors once, through the multicast address; the ragged ends byte by byte into every copy
        uint8_t *mdst = pd.mc + pd.off + o_begin - ph;
        for (uint32_t c = l * 16; c < hi_b; c += 32 * 16) {
            if (c >= lo_b && c + 16 <= hi_b) multimem_st16(mdst + c, *reinterpret_cast<const uint4 *>(obuf + c));
            else
                for (uint32_t k = 0; k < pd.n; k++) {
                    uint8_t *dst = pd.ptr[k] + pd.off + o_begin - ph;
                    for (uint32_t j = c > lo_b ? c : lo_b; j < c + 16 && j < hi_b; j++) dst[j] = obuf[j];
                }
        }
        return;
    }
    for (uint32_t k = 0; k < pd.n; k++) {
        uint8_t *dst = pd.ptr[k] + pd.off + o_begin - ph;                     // 16-byte aligned
        for (uint32_t c = l * 16; c < hi_b; c += 32 * 16) {
            if (c >= lo_b && c + 16 <= hi_b) *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(obuf + c);
            else for (uint32_t j = c > lo_b ? c : lo_b; j < c + 16 && j < hi_b; j++) dst[j] = obuf[j];
        }
    }
}

template <bool DUMP>
__global__ void __launch_bounds__(FTH, 3) k_front_write(const uint8_t *__restrict__ bed, uint64_t n, uint32_t skip, uint32_t halo, int64_t carry_max,
                                                     const FAgg *__restrict__ span_pre, const FAgg *__restrict__ tile_pre, const FAgg *__restrict__ chunk_pre,
                                                     const uint32_t *__restrict__ chunk_out, uint64_t n_lines, const __grid_constant__ PeerDst pd,
                                                     ChromSeed *seeds, unsigned long long *stat_len, unsigned long long *stat_uniq, uint32_t stat_slots,
                                                     uint64_t diag_chroms, unsigned long long *sc, DumpArrays da)
{
    __shared__ WarpSh Ws[FWARPS];
    __shared__ __align__(16) uint8_t s_out[FWARPS][FOB + 16];
    const unsigned wid = threadIdx.x >> 5;
    const uint64_t chunk = (uint64_t)blockIdx.x * FWARPS + wid;
    WarpSh &W = Ws[wid];
    ChunkGeom G;
    chunk_setup(W, G, bed, n, skip, chunk);
    if (G.k == 0 || G.k > FMAXL) return;
    const FAgg ex0 = fagg_op(fagg_op(span_pre[blockIdx.x / FSCAN_T], tile_pre[blockIdx.x]), chunk_pre[chunk]);
    write_chunk<DUMP>(W, G, bed, halo, carry_max, ex0, chunk_out[chunk], n_lines, pd, s_out[wid], blockIdx.x, seeds, stat_len, stat_uniq, stat_slots,
                      diag_chroms, sc, da);
}

// ---- both passes in one kernel ------------------------------------------------------------------------------------
// k_front_fused measures the 8 chunks of a tile, publishes the tile's aggregate, gets the aggregate of everything before the
// tile by a decoupled look-back over the tiles (its first warp reads the status of 32 predecessors at a time and folds what
// they have published: own aggregates up to the nearest one that already knows its inclusive prefix), and every warp writes
// its chunk's transformed lines -- with the masks and the line list it already has.  No second read of the input from HBM, no second chunk setup, no scan kernels.
// Tiles (8 chunks, one per warp) are taken from a ticket, so a chunk's predecessors have always started.  The totals are not
// known before the launch: the destination and the chromosome tables have a capacity, a chunk that would exceed it raises
// SC_OVERFLOW and writes nothing (the host then takes the two-pass form).  For the one-shot entries only: the ranges of the
// N-GPU path and of the chained entry need their offsets from outside (k_front_measure / k_front_write).
#ifndef S3G_FUSED_OCC
#define S3G_FUSED_OCC 4
#endif
struct FusedTables {
    FAgg *agg, *incl;             // per tile of 8 chunks: its own aggregate, the aggregate of everything up to and including it
    uint32_t *flag;               // per tile: generation << 2 | state (1: aggregate published, 2: inclusive prefix published)
    uint32_t *ticket;
    uint32_t gen;
    uint64_t n_chunks;
    uint64_t tf_cap, chrom_cap;
};
enum { SC_OVERFLOW = 6, SC_TICKET = 7 };

__device__ __forceinline__ FAgg fagg_load_cg(const FAgg *p)
{
    const unsigned long long *q = reinterpret_cast<const unsigned long long *>(p);
    FAgg a;
    a.lines = __ldcg(q); a.out = __ldcg(q + 1); a.chroms = __ldcg(q + 2); a.v = (int64_t)__ldcg(q + 3);
    const unsigned long long t = __ldcg(q + 4);
    a.seg = (uint32_t)t; a.pad = 0;
    return a;
}
__device__ __forceinline__ FAgg fagg_shfl_down(const FAgg &a, int d)
{
    FAgg r;
    r.lines = __shfl_down_sync(0xffffffffu, a.lines, d); r.out = __shfl_down_sync(0xffffffffu, a.out, d);
    r.chroms = __shfl_down_sync(0xffffffffu, a.chroms, d); r.v = __shfl_down_sync(0xffffffffu, a.v, d);
    r.seg = __shfl_down_sync(0xffffffffu, a.seg, d); r.pad = 0;
    return r;
}
__device__ __forceinline__ FAgg fagg_shfl(const FAgg &a, int src)
{
    FAgg r;
    r.lines = __shfl_sync(0xffffffffu, a.lines, src); r.out = __shfl_sync(0xffffffffu, a.out, src);
    r.chroms = __shfl_sync(0xffffffffu, a.chroms, src); r.v = __shfl_sync(0xffffffffu, a.v, src);
    r.seg = __shfl_sync(0xffffffffu, a.seg, src); r.pad = 0;
    return r;
}

__global__ void __launch_bounds__(FTH, S3G_FUSED_OCC) k_front_fused(const uint8_t *__restrict__ bed, uint64_t n, uint32_t skip, FusedTables T,
                                                     const __grid_constant__ PeerDst pd, ChromSeed *seeds, unsigned long long *stat_len,
                                                     unsigned long long *stat_uniq, uint32_t stat_slots, unsigned long long *sc)
{
    __shared__ WarpSh Ws[FWARPS];
    __shared__ __align__(16) uint8_t s_out[FWARPS][FOB + 16];
    __shared__ FAgg s_agg[FWARPS];             // the chunks' own aggregates
    __shared__ FAgg s_ex;                      // what everything before the tile leaves
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(T.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const unsigned l = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t chunk = (uint64_t)tile * FWARPS + wid;
    WarpSh &W = Ws[wid];
    ChunkGeom G;
    chunk_setup(W, G, bed, n, skip, chunk);
    FAgg A;
    uint32_t malformed;
    measure_chunk(W, G, bed, 0, sc, A, malformed);
    const unsigned long long last = chunk_last_newline(W, G, bed, n);
    if (l == 0) {
        if (malformed) atomicAdd(&sc[SC_MALFORMED], 1ull);
        if (last) atomicMax(&sc[SC_LASTNL], last);
        s_agg[wid] = A;
    }
    __syncthreads();
    // ---- the tile's first warp: publish the tile's aggregate, look back over the tiles before it ----
    // (the status tables are per TILE: with a warp's chunk as the unit every warp of a wave that runs in step would have to
    // fold the whole wave behind it, 3500 aggregates; 444 tiles are in flight at most, 14 steps of 32)
    const uint32_t gtag = T.gen << 2;
    if (wid == 0) {
        FAgg TA = fagg_identity();
#pragma unroll
        for (int w = 0; w < FWARPS; w++) TA = fagg_op(TA, s_agg[w]);
        if (l == 0) {
            T.agg[tile] = TA;
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(&T.flag[tile]) = gtag | 1u;
        }
        FAgg ex = fagg_identity();
        for (int64_t j = (int64_t)tile - 1; j >= 0; j -= 32) {
            const int64_t idx = j - (int64_t)l;
            uint32_t st;
            uint32_t spins = 0;
            for (;;) {
                st = idx >= 0 ? *reinterpret_cast<volatile const uint32_t *>(&T.flag[idx]) : (gtag | 2u);
                const bool ready = (st >> 2) == T.gen && (st & 3u) != 0;
                if (__all_sync(0xffffffffu, ready)) break;
                if (++spins > (1u << 24)) __trap();              // a predecessor never published: fail loudly instead of hanging
                __nanosleep(40);
            }
            __threadfence();
            const unsigned incl_m = __ballot_sync(0xffffffffu, (st & 3u) == 2u);
            const int m = incl_m ? __ffs((int)incl_m) - 1 : 32;  // the nearest predecessor that knows its inclusive prefix
            FAgg V = fagg_identity();
            if (idx >= 0) {
                if ((int)l < m) V = fagg_load_cg(&T.agg[idx]);
                else if ((int)l == m) V = fagg_load_cg(&T.incl[idx]);
            }
            // lane 0 <- V[31] (+) .. (+) V[1] (+) V[0]: higher lanes are earlier tiles
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const FAgg O = fagg_shfl_down(V, d);
                if (l + d < 32) V = fagg_op(O, V);
            }
            ex = fagg_op(fagg_shfl(V, 0), ex);
            if (m < 32) break;
        }
        if (l == 0) {
            const FAgg incl = fagg_op(ex, TA);
            T.incl[tile] = incl;
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(&T.flag[tile]) = gtag | 2u;
            s_ex = ex;
            if ((uint64_t)tile * FWARPS + FWARPS == T.n_chunks) {
                sc[SC_TOTAL + 0] = incl.lines; sc[SC_TOTAL + 1] = incl.out; sc[SC_TOTAL + 2] = incl.chroms;
                sc[SC_TOTAL + 3] = (unsigned long long)incl.v; sc[SC_TOTAL + 4] = incl.seg;
            }
        }
    }
    __syncthreads();
    if (G.k == 0 || G.k > FMAXL) return;
    FAgg ex = s_ex;
    for (unsigned w = 0; w < wid; w++) ex = fagg_op(ex, s_agg[w]);
    const FAgg incl = fagg_op(ex, A);
    if (incl.out > T.tf_cap || incl.chroms > T.chrom_cap) { if (l == 0) sc[SC_OVERFLOW] = 1; return; }
    DumpArrays da = {nullptr, nullptr, nullptr, nullptr, nullptr};
    write_chunk<false>(W, G, bed, 0, INT64_MIN, ex, (uint32_t)A.out, 0, pd, s_out[wid], tile, seeds, stat_len, stat_uniq, stat_slots, ~0ull, sc, da);
}

__global__ void k_chrom_finish(const ChromSeed *seeds, const unsigned long long *stat_len, const unsigned long long *stat_uniq, uint32_t stat_slots,
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
    unsigned long long sl = 0, su = 0;
    for (uint32_t k = 0; k < stat_slots; k++) { sl += stat_len[c * stat_slots + k]; su += stat_uniq[c * stat_slots + k]; }
    r.bases_nonunique = (int64_t)sl;
    r.bases_unique = (int64_t)su;
    r.bz_off = 0; r.bz_len = 0;
    out[c] = r;
}

// kernel (1): measures the range d_bed[0, n).  `halo` = 1: the range's first line is the last line BEFORE it (multi-GPU
// ranges, shard.cu): it hands its stop, length and chromosome to the second line and is itself neither written nor counted.
int run_tokenize(Ctx *ctx, const uint8_t *d_bed, uint64_t n, uint32_t skip, TfResult *out, uint32_t halo)
{
    *out = TfResult();
    uint64_t ntiles = (n + (uint64_t)FCH * FWARPS - 1) / ((uint64_t)FCH * FWARPS);
    if (ntiles == 0) ntiles = 1;
    if (ntiles > 0x7fffffffull) { set_error("input too large"); return S3G_E_LIMIT; }
    S3G_TRY(ctx->tile_cnt.ensure(ntiles * FWARPS * 4));                  // output bytes per chunk
    S3G_TRY(ctx->scan_a.ensure(ntiles * sizeof(FAgg)));                  // per tile: aggregate, then exclusive prefix
    S3G_TRY(ctx->scan_b.ensure(ntiles * FWARPS * sizeof(FAgg)));         // per chunk: prefix inside its tile
    S3G_TRY(ctx->scalars.ensure(64 * 8));
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    S3G_CUDA(cudaMemsetAsync(d_sc, 0, 64 * 8, ctx->stream));
    S3G_BYTES(ctx, n);
    S3G_LAUNCH(ctx, k_front_measure, (unsigned)ntiles, FTH, 0, d_bed, n, skip, halo, ctx->scan_a.as<FAgg>(), ctx->scan_b.as<FAgg>(),
               ctx->tile_cnt.as<uint32_t>(), (unsigned long long *)d_sc);
    const uint64_t nspans = (ntiles + FSCAN_T - 1) / FSCAN_T;
    S3G_TRY(ctx->scan_c.ensure(nspans * sizeof(FAgg)));
    S3G_LAUNCH(ctx, k_front_scan_tiles, (unsigned)nspans, FSCAN_T, 0, ctx->scan_a.as<FAgg>(), ntiles, ctx->scan_c.as<FAgg>());
    S3G_LAUNCH(ctx, k_front_scan_spans, 1, FSCAN_T, 0, ctx->scan_c.as<FAgg>(), nspans, (unsigned long long *)d_sc);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_sc, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    S3G_TRY(check_launch("front measure"));
    const uint64_t *h = ctx->h_scalars;
    out->n_lines = h[SC_TOTAL + 0];
    out->tf_len = h[SC_TOTAL + 1];
    out->n_chroms = h[SC_TOTAL + 2];
    out->dropped = out->n_lines ? n - h[SC_LASTNL] : n - skip;      // unterminated tail (hpp:181-190); the skip bytes are not ours
    ctx->front_halo = halo; ctx->front_skip = skip;
    ctx->front_tail_max = (int64_t)h[SC_TOTAL + 3];
    ctx->front_line1_flag = (uint32_t)h[SC_LINE1];
    if (out->n_lines && h[SC_MALFORMED]) {
        set_error("malformed BED line(s): fewer than three fields (in %llu 16 KiB tile(s) of the input)", (unsigned long long)h[SC_MALFORMED]);
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
int run_transform_rest(Ctx *ctx, const uint8_t *d_bed, uint64_t n, TfResult *out, uint32_t halo, int64_t carry_max, bool dump, bool last_part,
                       const uint64_t *peer_bufs, uint32_t n_peers, uint64_t peer_off, uint64_t multicast_buf)
{
    const uint64_t n_lines = out->n_lines, n_chroms = out->n_chroms;
    // a range that is not the last hands its last chromosome to the next range, which counts that chromosome's lines
    const uint64_t diag_chroms = last_part ? n_chroms : (n_chroms ? n_chroms - 1 : 0);
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    uint64_t ntiles = (n + (uint64_t)FCH * FWARPS - 1) / ((uint64_t)FCH * FWARPS);
    if (ntiles == 0) ntiles = 1;
    const uint32_t slots = n_chroms <= 4096 ? 32u : 1u;            // same-address atomics serialise: spread a chromosome's sums
    PeerDst pd;
    memset(&pd, 0, sizeof pd);
    if (n_peers) {
        if (n_peers > 8) { set_error("at most 8 destination buffers"); return S3G_E_PARAM; }
        for (uint32_t k = 0; k < n_peers; k++) {
            if (peer_bufs[k] & 15) { set_error("destination buffers must be 16-byte aligned"); return S3G_E_PARAM; }
            pd.ptr[k] = reinterpret_cast<uint8_t *>(peer_bufs[k]);
        }
        pd.n = n_peers; pd.off = peer_off;
        if (multicast_buf & 15) { set_error("multicast address must be 16-byte aligned"); return S3G_E_PARAM; }
        pd.mc = reinterpret_cast<uint8_t *>(multicast_buf);
    } else {
        S3G_TRY(ctx->tf.ensure(out->tf_len + 64));
        pd.ptr[0] = ctx->tf.as<uint8_t>(); pd.n = 1; pd.off = 0;
    }
    S3G_TRY(ctx->chrom_first.ensure((n_chroms + 1) * sizeof(ChromSeed)));
    S3G_TRY(ctx->stat_b.ensure((n_chroms + 1) * slots * 16));
    S3G_TRY(ctx->chroms.ensure((n_chroms + 1) * sizeof(s3g_chrom)));
    unsigned long long *stat_len = ctx->stat_b.as<unsigned long long>(), *stat_uniq = stat_len + (n_chroms + 1) * slots;
    S3G_CUDA(cudaMemsetAsync(stat_len, 0, (n_chroms + 1) * slots * 16, ctx->stream));
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
        S3G_LAUNCH(ctx, k_front_write<true>, (unsigned)ntiles, FTH, 0, d_bed, n, ctx->front_skip, halo, carry_max, ctx->scan_c.as<FAgg>(), ctx->scan_a.as<FAgg>(),
                   ctx->scan_b.as<FAgg>(), ctx->tile_cnt.as<uint32_t>(), n_lines, pd, ctx->chrom_first.as<ChromSeed>(),
                   stat_len, stat_uniq, slots, diag_chroms, (unsigned long long *)d_sc, da);
    } else {
        S3G_LAUNCH(ctx, k_front_write<false>, (unsigned)ntiles, FTH, 0, d_bed, n, ctx->front_skip, halo, carry_max, ctx->scan_c.as<FAgg>(), ctx->scan_a.as<FAgg>(),
                   ctx->scan_b.as<FAgg>(), ctx->tile_cnt.as<uint32_t>(), n_lines, pd, ctx->chrom_first.as<ChromSeed>(),
                   stat_len, stat_uniq, slots, diag_chroms, (unsigned long long *)d_sc, da);
    }
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + SC_UNSORTED, d_sc + SC_UNSORTED, 16, cudaMemcpyDeviceToHost, ctx->stream));   // read after the caller's next synchronise
    S3G_LAUNCH(ctx, k_chrom_finish, (unsigned)((n_chroms + 127) / 128), 128, 0, ctx->chrom_first.as<ChromSeed>(), stat_len, stat_uniq, slots,
               n_chroms, n_lines, out->tf_len, ctx->chroms.as<s3g_chrom>(), halo);
    return check_launch("transform write");
}

// both kernels of the front end as one launch (k_front_fused); S3G_E_CAPACITY: the destination or the chromosome tables were
// too small for this input -- the caller takes the two-pass form
constexpr uint64_t FUSED_CHROM_CAP = 4096;
static int run_transform_fused(Ctx *ctx, const uint8_t *d_bed, uint64_t n, uint32_t skip, TfResult *out)
{
    *out = TfResult();
    uint64_t ntiles = (n + (uint64_t)FCH * FWARPS - 1) / ((uint64_t)FCH * FWARPS);
    if (ntiles == 0) ntiles = 1;
    if (ntiles > 0x7fffffffull) { set_error("input too large"); return S3G_E_LIMIT; }
    const uint64_t n_chunks = ntiles * FWARPS;
    S3G_TRY(ctx->scan_b.ensure(ntiles * sizeof(FAgg)));
    S3G_TRY(ctx->front_incl.ensure(ntiles * sizeof(FAgg)));
    const size_t flag_cap = ctx->front_flag.cap;
    S3G_TRY(ctx->front_flag.ensure(ntiles * 4));
    if (ctx->front_flag.cap != flag_cap || ctx->front_gen >= (1u << 29)) {       // new table (or the tag wraps): start from zero
        S3G_CUDA(cudaMemsetAsync(ctx->front_flag.p, 0, ctx->front_flag.cap, ctx->stream));
        ctx->front_gen = 0;
    }
    S3G_TRY(ctx->scalars.ensure(64 * 8));
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    S3G_CUDA(cudaMemsetAsync(d_sc, 0, 64 * 8, ctx->stream));
    const uint64_t tf_cap = n + n / 8 + 4096;
    S3G_TRY(ctx->tf.ensure(tf_cap + 64));
    const uint32_t slots = 32;
    S3G_TRY(ctx->chrom_first.ensure((FUSED_CHROM_CAP + 1) * sizeof(ChromSeed)));
    S3G_TRY(ctx->stat_b.ensure((FUSED_CHROM_CAP + 1) * slots * 16));
    S3G_TRY(ctx->chroms.ensure((FUSED_CHROM_CAP + 1) * sizeof(s3g_chrom)));
    unsigned long long *stat_len = ctx->stat_b.as<unsigned long long>(), *stat_uniq = stat_len + (FUSED_CHROM_CAP + 1) * slots;
    S3G_CUDA(cudaMemsetAsync(stat_len, 0, (FUSED_CHROM_CAP + 1) * slots * 16, ctx->stream));
    FusedTables T;
    T.agg = ctx->scan_b.as<FAgg>(); T.incl = ctx->front_incl.as<FAgg>(); T.flag = ctx->front_flag.as<uint32_t>();
    T.ticket = reinterpret_cast<uint32_t *>(d_sc + SC_TICKET);
    T.gen = ++ctx->front_gen; T.n_chunks = n_chunks; T.tf_cap = tf_cap; T.chrom_cap = FUSED_CHROM_CAP;
    PeerDst pd;
    memset(&pd, 0, sizeof pd);
    pd.ptr[0] = ctx->tf.as<uint8_t>(); pd.n = 1; pd.off = 0;
    S3G_BYTES(ctx, 2.0 * (double)n);
    S3G_LAUNCH(ctx, k_front_fused, (unsigned)ntiles, FTH, 0, d_bed, n, skip, T, pd, ctx->chrom_first.as<ChromSeed>(), stat_len, stat_uniq, slots,
               (unsigned long long *)d_sc);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_sc, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    S3G_TRY(check_launch("front fused"));
    const uint64_t *h = ctx->h_scalars;
    out->n_lines = h[SC_TOTAL + 0];
    out->tf_len = h[SC_TOTAL + 1];
    out->n_chroms = h[SC_TOTAL + 2];
    out->dropped = out->n_lines ? n - h[SC_LASTNL] : n - skip;
    ctx->front_halo = 0; ctx->front_skip = skip;
    ctx->front_tail_max = (int64_t)h[SC_TOTAL + 3];
    if (out->n_lines && h[SC_MALFORMED]) {
        set_error("malformed BED line(s): fewer than three fields (in %llu 2 KiB chunk(s) of the input)", (unsigned long long)h[SC_MALFORMED]);
        return S3G_E_MALFORMED;
    }
    if (h[SC_OVERFLOW]) return S3G_E_CAPACITY;
    if (out->n_lines)
        S3G_LAUNCH(ctx, k_chrom_finish, (unsigned)((out->n_chroms + 127) / 128), 128, 0, ctx->chrom_first.as<ChromSeed>(), stat_len, stat_uniq, slots,
                   out->n_chroms, out->n_lines, out->tf_len, ctx->chroms.as<s3g_chrom>(), 0u);
    return check_launch("front fused");
}

int run_transform(Ctx *ctx, const uint8_t *d_bed, uint64_t n, TfResult *out, bool tokenize_only, uint32_t skip, bool last_part)
{
    // S3G_FRONT=one: one launch (k_front_fused) when nothing outside the range is needed and every chromosome of it is
    // counted.  Not the default: measured on B200 it is 4-12 % slower than the two passes (cfg3, 30 M lines: 3.50 ms against
    // 1.38 + 1.85 ms; cfg2: 1.63 against 1.57 ms) -- the measuring half runs at the writing half's register count, and what
    // the second chunk setup costs the two barriers and the look-back take back.  Kept for the tests and the record.
    const char *fe = getenv("S3G_FRONT");
    if (!tokenize_only && last_part && fe && !strcmp(fe, "one")) {
        int rc = run_transform_fused(ctx, d_bed, n, skip, out);
        if (rc != S3G_E_CAPACITY) return rc;
    }
    S3G_TRY(run_tokenize(ctx, d_bed, n, skip, out, 0));
    ctx->h_scalars[SC_UNSORTED] = 0; ctx->h_scalars[SC_CRLF] = 0;
    if (out->n_lines == 0) return S3G_OK;
    S3G_TRY(run_transform_rest(ctx, d_bed, n, out, 0, INT64_MIN, tokenize_only, last_part));
    if (tokenize_only) {
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        return check_launch("tokenize");
    }
    return S3G_OK;
}

}  // namespace s3g
