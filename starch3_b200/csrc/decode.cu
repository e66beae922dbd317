// decode.cu -- the decoder path: archive -> BED.
//
// The reference has no decoder; what is restated here is the reference's vendored libbz2
// decompressor -- BZ2_bzDecompress / BZ2_decompress (bz/bzlib.c:551-900, bz/decompress.c:106-646),
// BZ2_hbCreateDecodeTables (bz/huffman.c:172-205) -- and the inverse of update_transformation_state
// (hpp:428-504), as ARCHIVE_FORMAT.md states it.  bz/ = third-party/bzip2-1.0.6.tar.gz of the reference,
// hpp = /root/reference/include/starch3api.hpp.
//
//   k_find_magic      every bit offset of the payload is tested for the 48-bit block / end-of-stream magic
//                     (blocks are not byte aligned, bz/compress.c:609); the host chains the hits stream by stream
//                     using the end position each block decode reports
//   k_bz_decode       per block: header, Huffman tables, then the serial symbol loop (a symbol's bit position depends
//                     on all symbols before it): canonical decode through a 9-bit look-up table, RUNA/RUNB runs,
//                     inverse move-to-front -> the BWT last column.  One thread per block; blocks are what runs in parallel.
//   k_ibwt_build      the T vector of the inverse BWT (bz/decompress.c:501-515) as a stable counting sort, CTA per block
//   k_ibwt_walk_a/b   the n-step pointer chase (bz/bzlib.c:606-640) cut at 4096 marked positions: every piece is walked
//                     by its own thread, k_ibwt_order ranks the pieces along the cycle, the second walk writes bytes
//   k_unrle           undoes the initial run-length coding (bz/bzlib.c:592-640): a warp per block, 32 bytes per step,
//                     serial only where four equal bytes meet
//   k_inv_*           inverse transform: a thread per line; the current length and the running stop are one
//                     segmented scan, the output offsets another
#include <algorithm>
#include <string>
#include "common.cuh"
#include "scan.cuh"

namespace s3g {

constexpr uint64_t BLOCK_MAGIC = 0x314159265359ull;    // bz/compress.c:632-633
constexpr uint64_t END_MAGIC = 0x177245385090ull;      // bz/compress.c:657-658
constexpr uint32_t IBWT_K = 4096;                       // marked positions per block in the inverse BWT walk
constexpr uint32_t TT_MARK = 0x80000000u;
constexpr uint32_t NODE_NONE = 0xffffffffu;
// zero bytes after an uploaded buffer: a bit reader that decodes garbage (a look-alike of the block magic inside another
// block's data, a corrupt stream) stops inside them -- every loop of the block decoder ends on zero bits or at its
// 4096-symbol position check, 20 bits per symbol at most
constexpr uint64_t DEC_PAD = 32768;

struct DecBlock {
    uint64_t bit_pos;        // where the block magic starts (bits from the payload start)
    uint64_t end_bit;        // first bit after the block's last symbol
    uint64_t limit_bit;      // end of the stream that holds the block
    uint32_t nblock_max;     // 100000 * level (bz/decompress.c:201)
    uint32_t nblock;         // bytes in the block (BWT last column)
    uint32_t orig_ptr;
    uint32_t crc;            // stored block CRC
    int32_t status;          // 0 = decoded, < 0 = what went wrong
    uint32_t out_len;        // bytes after undoing RLE1
    uint64_t out_off;        // where they go in the output buffer
    uint32_t cycle;          // length of the cycle the inverse BWT walk runs through (== nblock unless the block is periodic)
    uint32_t start_pos;      // T[origPtr]
};

__device__ __forceinline__ uint32_t be32(const uint8_t *z, uint64_t word)
{
    return __byte_perm(reinterpret_cast<const uint32_t *>(z)[word], 0, 0x0123);
}

// candidates: (bit position << 1) | (1 = end-of-stream magic)
__global__ void k_find_magic(const uint8_t *__restrict__ z, uint64_t n_words, uint64_t n_bits, unsigned long long *cand, uint32_t *n_cand, uint32_t cap)
{
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint64_t x = ((uint64_t)be32(z, w) << 32) | be32(z, w + 1);      // the buffer is padded with zero words
    const uint32_t w2 = be32(z, w + 2);
#pragma unroll 4
    for (uint32_t s = 0; s < 32; s++) {
        const uint64_t v = (s ? (x << s) | ((uint64_t)w2 >> (32 - s)) : x) >> 16;
        if (v == BLOCK_MAGIC || v == END_MAGIC) {
            const uint64_t p = w * 32 + s;
            if (p + 48 <= n_bits) {
                uint32_t k = atomicAdd(n_cand, 1u);
                if (k < cap) cand[k] = (p << 1) | (v == END_MAGIC ? 1ull : 0ull);
            }
        }
    }
}

// ---- block decode --------------------------------------------------------------------------------------
enum { DE_OK = 0, DE_RANDOMISED = -1, DE_HEADER = -2, DE_SELECTOR = -3, DE_CODELEN = -4, DE_SYMBOL = -5, DE_OVERRUN = -6, DE_ORIGPTR = -7,
       DE_TRUNCATED = -8 };

struct BitReader {
    const uint8_t *z;
    uint64_t wi;           // next word to take
    uint32_t nextw;        // ... (as loaded, little endian) already on its way: the load is issued one word ahead, so its latency hides behind the ~8 symbols
                           // decoded from the word before it
    uint64_t buf;          // unread bits, left aligned
    int cnt;               // how many
    __device__ __forceinline__ void init(const uint8_t *zz, uint64_t bit)
    {
        z = zz; wi = bit >> 5; buf = 0; cnt = 0;
        nextw = reinterpret_cast<const uint32_t *>(z)[wi];
        if (bit & 31) get((int)(bit & 31));
    }
    __device__ __forceinline__ void fill()
    {
        // the byte swap happens here, at the use: swapping where the load is issued would wait for it on the spot
        buf |= (uint64_t)__byte_perm(nextw, 0, 0x0123) << (32 - cnt);
        cnt += 32;
        wi++;
        nextw = reinterpret_cast<const uint32_t *>(z)[wi];
    }
    __device__ __forceinline__ uint32_t get(int k)            // 1 <= k <= 32
    {
        if (cnt < k) fill();
        uint32_t v = (uint32_t)(buf >> (64 - k));
        buf <<= k; cnt -= k;
        return v;
    }
    __device__ __forceinline__ uint32_t peek(int k)
    {
        if (cnt < k) fill();
        return (uint32_t)(buf >> (64 - k));
    }
    __device__ __forceinline__ void skip(int k) { buf <<= k; cnt -= k; }
    __device__ __forceinline__ uint64_t pos() const { return wi * 32 - (uint64_t)cnt; }
};

constexpr int LUT_BITS = 9;
struct DecSh {
    uint8_t selector[18004];
    uint8_t len[6][258];
    uint16_t perm[6][258];
    int32_t limit[6][24], base[6][24];
    uint16_t lut[6][1 << LUT_BITS];      // (code length << 9) | symbol; 0 = longer than LUT_BITS
    uint8_t yy[256], seq2unseq[256];
    int32_t min_len[6];
};

__global__ void __launch_bounds__(32) k_bz_decode(const uint8_t *__restrict__ z, DecBlock *blocks, uint8_t *lcol)
{
    __shared__ DecSh S;
    DecBlock &B = blocks[blockIdx.x];
    if (threadIdx.x != 0) return;
    uint8_t *L = lcol + (uint64_t)blockIdx.x * BLK_STRIDE;
    BitReader br;
    br.init(z, B.bit_pos + 48);
    B.nblock = 0; B.end_bit = 0;
#define DEC_FAIL(code) do { B.status = (code); return; } while (0)
    B.crc = br.get(32);
    if (br.get(1)) DEC_FAIL(DE_RANDOMISED);                        // bz/decompress.c:226: never written by 1.0.x compressors
    const uint32_t orig = br.get(24);
    // symbols in use (bz/decompress.c:243-262)
    int n_in_use = 0;
    {
        const uint32_t in16 = br.get(16);
        for (int i = 0; i < 16; i++) {
            if (in16 & (0x8000u >> i)) {
                const uint32_t bits = br.get(16);
                for (int j = 0; j < 16; j++) if (bits & (0x8000u >> j)) S.seq2unseq[n_in_use++] = (uint8_t)(i * 16 + j);
            }
        }
    }
    if (n_in_use == 0) DEC_FAIL(DE_HEADER);
    const int alpha = n_in_use + 2;
    const int n_groups = (int)br.get(3);
    if (n_groups < 2 || n_groups > 6) DEC_FAIL(DE_HEADER);
    const int n_sel = (int)br.get(15);
    if (n_sel < 1 || n_sel > 18002) DEC_FAIL(DE_HEADER);
    // selectors: unary, then inverse move-to-front (bz/decompress.c:270-297)
    {
        uint8_t pos[6];
        for (int v = 0; v < n_groups; v++) pos[v] = (uint8_t)v;
        for (int i = 0; i < n_sel; i++) {
            int j = 0;
            while (br.get(1)) { j++; if (j >= n_groups) DEC_FAIL(DE_SELECTOR); }
            uint8_t tmp = pos[j];
            for (; j > 0; j--) pos[j] = pos[j - 1];
            pos[0] = tmp;
            S.selector[i] = tmp;
        }
    }
    // code lengths (bz/decompress.c:300-316)
    for (int t = 0; t < n_groups; t++) {
        int curr = (int)br.get(5);
        for (int i = 0; i < alpha; i++) {
            for (;;) {
                if (curr < 1 || curr > 20) DEC_FAIL(DE_CODELEN);
                if (!br.get(1)) break;
                curr += br.get(1) ? -1 : 1;
            }
            S.len[t][i] = (uint8_t)curr;
        }
    }
    // decode tables (bz/huffman.c:172-205) and the look-up table of the codes of up to LUT_BITS bits
    for (int t = 0; t < n_groups; t++) {
        int mn = 32, mx = 0;
        for (int i = 0; i < alpha; i++) { int l = S.len[t][i]; if (l > mx) mx = l; if (l < mn) mn = l; }
        S.min_len[t] = mn;
        int pp = 0;
        for (int i = mn; i <= mx; i++) for (int j = 0; j < alpha; j++) if (S.len[t][j] == i) S.perm[t][pp++] = (uint16_t)j;
        int32_t *base = S.base[t], *limit = S.limit[t];
        for (int i = 0; i < 24; i++) { base[i] = 0; limit[i] = 0; }
        for (int i = 0; i < alpha; i++) base[S.len[t][i] + 1]++;
        for (int i = 1; i < 23; i++) base[i] += base[i - 1];
        int32_t vec = 0;
        for (int i = mn; i <= mx; i++) { vec += base[i + 1] - base[i]; limit[i] = vec - 1; vec <<= 1; }
        for (int i = mn + 1; i <= mx; i++) base[i] = ((limit[i - 1] + 1) << 1) - base[i];
        for (int i = 0; i < (1 << LUT_BITS); i++) S.lut[t][i] = 0;
        // canonical codes in (length, symbol) order = the order of perm[]
        uint32_t code = 0; int k = 0;
        for (int l = mn; l <= mx && l <= LUT_BITS; l++) {
            for (; k < pp && S.len[t][S.perm[t][k]] == l; k++) {
                const uint32_t first = code << (LUT_BITS - l), cnt = 1u << (LUT_BITS - l);
                if (first + cnt > (1u << LUT_BITS)) DEC_FAIL(DE_CODELEN);           // over-subscribed lengths (corrupt table)
                for (uint32_t q = 0; q < cnt; q++) S.lut[t][first + q] = (uint16_t)((l << 9) | S.perm[t][k]);
                code++;
            }
            code <<= 1;
        }
    }
    // the symbols (bz/decompress.c:330-455)
    for (int i = 0; i < 256; i++) S.yy[i] = (uint8_t)i;
    // alphabets of up to 32 symbols (sorted BED: 12..20) keep the move-to-front list in four registers, one byte per entry
    // holding the byte VALUE (seqToUnseq folded in): a move to the front is shifts and masks, no memory
    const bool small = n_in_use <= 32;
    uint64_t l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    if (small)
        for (int i = n_in_use - 1; i >= 0; i--) {
            const uint64_t v = S.seq2unseq[i];
            l3 = (l3 << 8) | (l2 >> 56); l2 = (l2 << 8) | (l1 >> 56); l1 = (l1 << 8) | (l0 >> 56); l0 = (l0 << 8) | v;
        }
    const int EOB = n_in_use + 1;
    const uint32_t nmax = B.nblock_max;
    uint32_t nblock = 0;
    int group_no = -1, group_pos = 0, g = 0;
    auto next_sym = [&](int &sym) -> int {
        if (group_pos == 0) {
            group_no++;
            if (group_no >= n_sel) return DE_SELECTOR;
            group_pos = 50;
            g = S.selector[group_no];
        }
        group_pos--;
        const uint32_t e = S.lut[g][br.peek(LUT_BITS)];
        if (e) { br.skip((int)(e >> 9)); sym = (int)(e & 511u); return DE_OK; }
        int zn = S.min_len[g];
        int32_t zvec = (int32_t)br.get(zn);
        for (;;) {
            if (zn > 20) return DE_SYMBOL;
            if (zvec <= S.limit[g][zn]) break;
            zn++;
            zvec = (zvec << 1) | (int32_t)br.get(1);
        }
        const int32_t idx = zvec - S.base[g][zn];
        if (idx < 0 || idx >= 258) return DE_SYMBOL;
        sym = S.perm[g][idx];
        return DE_OK;
    };
    int sym = 0, rc;
    if ((rc = next_sym(sym)) != DE_OK) DEC_FAIL(rc);
    for (;;) {
        if (sym == EOB) break;
        if (sym <= 1) {                                         // RUNA / RUNB: a run of the symbol at the front of the list
            int es = -1, N = 1;
            do {
                if (N >= 2 * 1024 * 1024) DEC_FAIL(DE_SYMBOL);
                es += (sym == 0 ? 1 : 2) * N;
                N <<= 1;
                if ((rc = next_sym(sym)) != DE_OK) DEC_FAIL(rc);
            } while (sym <= 1);
            es++;
            const uint8_t uc = small ? (uint8_t)l0 : S.seq2unseq[S.yy[0]];
            if ((uint64_t)nblock + (uint64_t)es > nmax) DEC_FAIL(DE_OVERRUN);
            for (int q = 0; q < es; q++) L[nblock + q] = uc;
            nblock += (uint32_t)es;
            continue;
        }
        if (nblock >= nmax) DEC_FAIL(DE_OVERRUN);
        const int nn = sym - 1;
        if (nn >= n_in_use) DEC_FAIL(DE_SYMBOL);
        if (small) {
            const int sh = (nn & 7) * 8;
            uint64_t v;
            if (nn < 8) {
                v = (l0 >> sh) & 0xffull;
                l0 = (l0 & ~((2ull << (sh + 7)) - 1ull)) | ((l0 & ((1ull << sh) - 1ull)) << 8) | v;
            } else {
                const int w = nn >> 3;
                const uint64_t cur = w == 1 ? l1 : (w == 2 ? l2 : l3);
                v = (cur >> sh) & 0xffull;
                const uint64_t upd = (cur & ~((2ull << (sh + 7)) - 1ull)) | ((cur & ((1ull << sh) - 1ull)) << 8);
                // the words below w move up by one entry; w takes the entry that falls out of w - 1
                const uint64_t c0 = l0 >> 56, c1 = l1 >> 56, c2 = l2 >> 56;
                l0 = (l0 << 8) | v;
                if (w == 1) l1 = upd | c0;
                else {
                    l1 = (l1 << 8) | c0;
                    if (w == 2) l2 = upd | c1;
                    else { l2 = (l2 << 8) | c1; l3 = upd | c2; }
                }
            }
            L[nblock++] = (uint8_t)v;
        } else {
            const uint8_t uc = S.yy[nn];
            for (int q = nn; q > 0; q--) S.yy[q] = S.yy[q - 1];
            S.yy[0] = uc;
            L[nblock++] = S.seq2unseq[uc];
        }
        if ((rc = next_sym(sym)) != DE_OK) DEC_FAIL(rc);
        if ((nblock & 4095u) == 0 && br.pos() > B.limit_bit) DEC_FAIL(DE_TRUNCATED);
    }
    if (orig >= nblock) DEC_FAIL(DE_ORIGPTR);                   // bz/decompress.c:461
    if (br.pos() > B.limit_bit) DEC_FAIL(DE_TRUNCATED);
    B.nblock = nblock; B.orig_ptr = orig; B.end_bit = br.pos(); B.status = DE_OK;
#undef DEC_FAIL
}

// ---- inverse BWT ---------------------------------------------------------------------------------------
// tt[k] = (T[k] << 8) | L[k] with T from the stable counting sort of L (bz/decompress.c:501-515); bit 31 marks the
// positions where the walk is cut.  scratch: 256 * 1024 words per block.
constexpr int IB_T = 1024;
__global__ void __launch_bounds__(IB_T) k_ibwt_build(DecBlock *blocks, const uint8_t *lcol, uint32_t *tt_all, uint32_t *scratch)
{
    DecBlock &B = blocks[blockIdx.x];
    if (B.status != DE_OK) return;
    const int n = (int)B.nblock;
    const uint8_t *L = lcol + (uint64_t)blockIdx.x * BLK_STRIDE;
    uint32_t *tt = tt_all + (uint64_t)blockIdx.x * BLK_STRIDE;
    uint32_t *cnt = scratch + (uint64_t)blockIdx.x * 256 * IB_T;        // [256][IB_T]
    __shared__ uint32_t s_tot[256], s_start[256];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int seg = (n + IB_T - 1) / IB_T;
    const int i0 = t * seg < n ? t * seg : n, i1 = i0 + seg < n ? i0 + seg : n;
    for (int c = 0; c < 256; c++) cnt[c * IB_T + t] = 0;
    for (int i = i0; i < i1; i++) cnt[L[i] * IB_T + t]++;
    __syncthreads();
    for (int c = warp; c < 256; c += IB_T / 32) {
        uint32_t run = 0;
        for (int k = 0; k < IB_T; k += 32) {
            uint32_t v = cnt[c * IB_T + k + lane];
            uint32_t inc = warp_incl_sum<uint32_t>(v);
            cnt[c * IB_T + k + lane] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_tot[c] = run;
    }
    __syncthreads();
    if (t == 0) { uint32_t acc = 0; for (int c = 0; c < 256; c++) { s_start[c] = acc; acc += s_tot[c]; } }
    __syncthreads();
    for (int i = i0; i < i1; i++) {
        uint32_t c = L[i];
        uint32_t k = s_start[c] + cnt[c * IB_T + t]++;
        tt[k] = (uint32_t)i << 8;
    }
    __syncthreads();
    for (int i = i0; i < i1; i++) tt[i] |= L[i];
    __syncthreads();
    const uint32_t stride = ((uint32_t)n + IBWT_K - 1) / IBWT_K;
    const uint32_t start = (tt[B.orig_ptr] >> 8) & 0xfffffu;          // bz/decompress.c:517 (the mark bit may already be set)
    for (uint32_t j = t; j * stride < (uint32_t)n; j += IB_T) tt[j * stride] |= TT_MARK;
    __syncthreads();
    if (t == 0) { tt[start] |= TT_MARK; B.start_pos = start; }
}

// node j < ks: the piece that starts at position j * stride; node ks: the piece that starts at start_pos when that is
// not a multiple of stride.  nodes: [nb][IBWT_K + 2] of (next, len).
__device__ __forceinline__ void ibwt_geom(const DecBlock &B, uint32_t *stride, uint32_t *ks, uint32_t *sid)
{
    const uint32_t n = B.nblock;
    *stride = (n + IBWT_K - 1) / IBWT_K;
    *ks = (n + *stride - 1) / *stride;
    *sid = (B.start_pos % *stride == 0) ? B.start_pos / *stride : *ks;
}

__global__ void __launch_bounds__(256) k_ibwt_walk_a(const DecBlock *blocks, const uint32_t *tt_all, uint2 *nodes)
{
    const DecBlock &B = blocks[blockIdx.y];
    if (B.status != DE_OK) return;
    uint32_t stride, ks, sid;
    ibwt_geom(B, &stride, &ks, &sid);
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > ks || (j == ks && sid != ks)) return;
    const uint32_t *tt = tt_all + (uint64_t)blockIdx.y * BLK_STRIDE;
    uint32_t p = j < ks ? j * stride : B.start_pos;
    uint32_t w = tt[p], len = 0;
    const uint32_t n = B.nblock;
    for (;;) {
        len++;
        p = (w >> 8) & 0xfffffu;
        w = tt[p];
        if ((w & TT_MARK) || len >= n) break;
    }
    const uint32_t nid = (p == B.start_pos && sid == ks) ? ks : p / stride;
    nodes[(uint64_t)blockIdx.y * (IBWT_K + 2) + j] = make_uint2(nid, len);
}

// ranks the pieces along the cycle that starts at start_pos: offs[j] = output offset of piece j (NODE_NONE: never reached)
__global__ void __launch_bounds__(256) k_ibwt_order(DecBlock *blocks, const uint2 *nodes, uint32_t *offs)
{
    DecBlock &B = blocks[blockIdx.x];
    if (B.status != DE_OK) return;
    __shared__ uint2 s_node[IBWT_K + 2];       // (next, len); a visited node becomes (offset, NODE_NONE)
    uint32_t stride, ks, sid;
    ibwt_geom(B, &stride, &ks, &sid);
    for (uint32_t j = threadIdx.x; j <= ks; j += blockDim.x) s_node[j] = nodes[(uint64_t)blockIdx.x * (IBWT_K + 2) + j];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t cur = sid, off = 0, steps = 0;
        while (cur <= ks && s_node[cur].y != NODE_NONE && steps <= ks + 1) {
            const uint2 nd = s_node[cur];
            s_node[cur] = make_uint2(off, NODE_NONE);
            off += nd.y;
            cur = nd.x;
            steps++;
        }
        B.cycle = off;
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j <= ks; j += blockDim.x)
        offs[(uint64_t)blockIdx.x * (IBWT_K + 2) + j] = s_node[j].y == NODE_NONE ? s_node[j].x : NODE_NONE;
}

__global__ void __launch_bounds__(256) k_ibwt_walk_b(const DecBlock *blocks, const uint32_t *tt_all, const uint32_t *offs, uint8_t *out_all)
{
    const DecBlock &B = blocks[blockIdx.y];
    if (B.status != DE_OK) return;
    uint32_t stride, ks, sid;
    ibwt_geom(B, &stride, &ks, &sid);
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > ks || (j == ks && sid != ks)) return;
    uint32_t o = offs[(uint64_t)blockIdx.y * (IBWT_K + 2) + j];
    if (o == NODE_NONE) return;                                   // a piece of another cycle (periodic block): never visited
    const uint32_t *tt = tt_all + (uint64_t)blockIdx.y * BLK_STRIDE;
    uint8_t *out = out_all + (uint64_t)blockIdx.y * BLK_STRIDE;
    const uint32_t n = B.nblock, cyc = B.cycle;
    uint32_t p = j < ks ? j * stride : B.start_pos;
    // output k of the walk is the low byte of tt[pos_k], pos_0 = T[origPtr], pos_k+1 = T[pos_k] (bz/bzlib.c:606-640, BZ_GET_FAST)
    uint32_t w = tt[p], len = 0;
    for (;;) {
        const uint8_t ch = (uint8_t)w;
        if (cyc == n) out[o] = ch;
        else for (uint32_t q = o; q < n; q += cyc) out[q] = ch;   // the walk goes round a shorter cycle n / cyc times
        o++; len++;
        p = (w >> 8) & 0xfffffu;
        w = tt[p];
        if ((w & TT_MARK) || len >= n) break;
    }
}

// ---- undo RLE1 -----------------------------------------------------------------------------------------
// four equal bytes are followed by a count of further copies (bz/bzlib.c:592-640).  A warp takes 32 bytes per step; a step
// in which no run can reach four is a plain copy.
template <bool WRITE> __device__ __forceinline__ uint32_t unrle_walk(const uint8_t *__restrict__ in, uint32_t n, uint8_t *__restrict__ out)
{
    const unsigned l = threadIdx.x & 31;
    uint32_t o = 0, run = 0, last = 0, pending = 0;             // run: equal bytes ending at the previous position (0: fresh)
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + l, nv = n - base < 32 ? n - base : 32;
        const bool valid = i < n;
        const uint32_t c = valid ? in[i] : 0x100u;
        uint32_t pc = __shfl_up_sync(0xffffffffu, c, 1);
        if (l == 0) pc = run ? last : 0x200u;
        const unsigned eq = __ballot_sync(0xffffffffu, valid && c == pc);
        const uint32_t had = run ? run - 1 : 0;                  // equalities already in hand (0..2)
        const uint64_t E = ((uint64_t)eq << 3) | (uint64_t)((0x7u << (3 - had)) & 0x7u);
        const bool has4 = (E & (E >> 1) & (E >> 2)) != 0;
        if (!pending && !has4) {
            if (WRITE && valid) out[o + l] = (uint8_t)c;
            o += nv;
            const unsigned vm = nv == 32 ? 0xffffffffu : ((1u << nv) - 1u);
            if ((eq & vm) == vm) run += nv;                      // (nv <= 2 here, or a run would have reached four)
            else run = 1 + (uint32_t)__clz((int)~((eq & vm) << (32 - nv)));
            last = __shfl_sync(0xffffffffu, c, (int)nv - 1);
        } else {
            if (l == 0) {
                for (uint32_t k = 0; k < nv; k++) {
                    const uint32_t b = in[base + k];
                    if (pending) {
                        if (WRITE) for (uint32_t q = 0; q < b; q++) out[o + q] = (uint8_t)last;
                        o += b; pending = 0; run = 0;
                        continue;
                    }
                    if (run && b == last) run++; else { run = 1; last = b; }
                    if (WRITE) out[o] = (uint8_t)b;
                    o++;
                    if (run == 4) pending = 1;
                }
            }
            o = __shfl_sync(0xffffffffu, o, 0); run = __shfl_sync(0xffffffffu, run, 0);
            last = __shfl_sync(0xffffffffu, last, 0); pending = __shfl_sync(0xffffffffu, pending, 0);
            if (pending) run = 4;
        }
    }
    return o;
}

// A block is cut into UNRLE_S segments, a warp each.  A segment starts at a position q with byte[q] != byte[q-1] and
// byte[q-1] != byte[q-2]: q cannot be a count byte (that needs four equal bytes before it) and it opens a new run whatever
// came before, so the walk can start there with a fresh state.  seg_len: [blocks][UNRLE_S] decoded bytes per segment.
constexpr uint32_t UNRLE_S = 64;
__device__ __forceinline__ uint32_t unrle_sync_point(const uint8_t *__restrict__ in, uint32_t n, uint32_t from)
{
    const unsigned l = threadIdx.x & 31;
    if (from == 0) return 0;
    for (uint32_t base = from; base < n; base += 32) {
        const uint32_t i = base + l;
        bool ok = false;
        if (i < n && i >= 2) { const uint8_t b0 = in[i], b1 = in[i - 1], b2 = in[i - 2]; ok = b0 != b1 && b1 != b2; }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) return base + (uint32_t)__ffs((int)m) - 1;
    }
    return n;
}
template <bool WRITE> __global__ void __launch_bounds__(128) k_unrle(DecBlock *blocks, uint32_t nb, const uint8_t *in_all, uint8_t *out, uint32_t *seg_len)
{
    const uint32_t w = blockIdx.x * 4 + (threadIdx.x >> 5);
    const uint32_t b = w / UNRLE_S, sg = w % UNRLE_S;
    const unsigned l = threadIdx.x & 31;
    if (b >= nb) return;
    DecBlock &B = blocks[b];
    if (B.status != DE_OK) return;
    const uint8_t *in = in_all + (uint64_t)b * BLK_STRIDE;
    const uint32_t n = B.nblock;
    const uint32_t per = ((n + UNRLE_S - 1) / UNRLE_S + 31) & ~31u;
    const uint32_t nom0 = min(sg * per, n), nom1 = min((sg + 1) * per, n);
    const uint32_t s0 = unrle_sync_point(in, n, nom0);
    const uint32_t s1 = sg + 1 == UNRLE_S ? n : unrle_sync_point(in, n, nom1);
    uint8_t *dst = nullptr;
    if (WRITE) {
        uint32_t before = 0;
        for (uint32_t k = l; k < sg; k += 32) before += seg_len[(uint64_t)b * UNRLE_S + k];
        before = __reduce_add_sync(0xffffffffu, before);
        dst = out + B.out_off + before;
    }
    const uint32_t len = s1 > s0 ? unrle_walk<WRITE>(in + s0, s1 - s0, dst) : 0;
    if (!WRITE && l == 0) { seg_len[(uint64_t)b * UNRLE_S + sg] = len; atomicAdd(&B.out_len, len); }
}

// ---- inverse transform ---------------------------------------------------------------------------------
constexpr int IV_T = 256;
constexpr int IV_TILE = IV_T * 16;

__global__ void __launch_bounds__(IV_T) k_inv_count(const uint8_t *__restrict__ tf, uint64_t n, uint64_t *tile_cnt)
{
    __shared__ uint32_t sm[33];
    const uint64_t pos = (uint64_t)blockIdx.x * IV_TILE + (uint64_t)threadIdx.x * 16;
    uint32_t c = 0;
    for (int k = 0; k < 16; k++) if (pos + k < n && tf[pos + k] == '\n') c++;
    uint32_t tot;
    block_excl_sum<uint32_t>(c, sm, &tot);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
}
struct SumU64 {
    typedef uint64_t T;
    __host__ __device__ static T identity() { return 0; }
    __host__ __device__ static T op(T a, T b) { return a + b; }
};
__global__ void __launch_bounds__(IV_T) k_inv_starts(const uint8_t *__restrict__ tf, uint64_t n, const uint64_t *tile_base, uint64_t *line_start)
{
    __shared__ uint32_t sm[33];
    const uint64_t pos = (uint64_t)blockIdx.x * IV_TILE + (uint64_t)threadIdx.x * 16;
    uint32_t c = 0;
    for (int k = 0; k < 16; k++) if (pos + k < n && tf[pos + k] == '\n') c++;
    uint32_t tot;
    uint32_t ex = block_excl_sum<uint32_t>(c, sm, &tot);
    uint64_t idx = tile_base[blockIdx.x] + ex + 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
    for (int k = 0; k < 16; k++) if (pos + k < n && tf[pos + k] == '\n') line_start[idx++] = pos + k + 1;
}

// what the lines [a, b) do to (current length, running stop); associative, a stream start forgets what came before
struct InvAgg {
    int64_t sum;          // sum of the deltas, plus the lengths of the elements after the range's first 'p' line
    int64_t last_p;       // value of the range's last 'p' line
    uint32_t before_p;    // elements before the range's first 'p' line (they take the length in force before the range)
    uint32_t flags;       // 1 = has a 'p' line, 2 = a stream starts inside
};
struct InvScan {
    typedef InvAgg T;
    const uint8_t *tf; const uint64_t *line_start; const uint64_t *soff; uint64_t n_streams;
    int64_t *start, *stop; uint32_t *out_len, *sid_out; const uint32_t *name_len;
    __host__ __device__ static T identity() { T t; t.sum = 0; t.last_p = 0; t.before_p = 0; t.flags = 0; return t; }
    __host__ __device__ static T op(const T &a, const T &b)
    {
        if (b.flags & 2u) return b;
        T r;
        r.before_p = a.before_p + ((a.flags & 1u) ? 0u : b.before_p);
        r.sum = a.sum + b.sum + ((a.flags & 1u) ? (int64_t)b.before_p * a.last_p : 0);
        r.last_p = (b.flags & 1u) ? b.last_p : a.last_p;
        r.flags = (a.flags & 2u) | ((a.flags | b.flags) & 1u);
        return r;
    }
    __device__ uint32_t stream_of(uint64_t s) const       // the stream that holds position s
    {
        uint64_t lo = 0, hi = n_streams;                   // largest k with soff[k] <= s
        while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (soff[mid] <= s) lo = mid; else hi = mid; }
        return (uint32_t)lo;
    }
    __device__ int64_t number(uint64_t p, uint64_t e, uint64_t *endp) const
    {
        int neg = 0; uint64_t acc = 0;
        if (p < e && (tf[p] == '-' || tf[p] == '+')) { neg = tf[p] == '-'; p++; }
        while (p < e) { uint32_t d = (uint32_t)tf[p] - '0'; if (d > 9u) break; acc = acc * 10 + d; p++; }
        *endp = p;
        return neg ? (int64_t)(0 - acc) : (int64_t)acc;
    }
    __device__ T load(uint64_t i) const
    {
        const uint64_t s = line_start[i], e = line_start[i + 1] - 1;
        T t = identity();
        uint64_t endp;
        if (s < e && tf[s] == 'p') { t.last_p = number(s + 1, e, &endp); t.flags = 1; }
        else { t.sum = number(s, e, &endp); t.before_p = 1; }
        const uint32_t k = stream_of(s);
        if (soff[k] == s) t.flags |= 2u;
        return t;
    }
    __device__ void store(uint64_t i, T excl, T val) const
    {
        const uint64_t s = line_start[i], e = line_start[i + 1] - 1;
        const uint32_t k = stream_of(s);
        sid_out[i] = k;
        if (val.flags & 1u) { out_len[i] = 0; return; }                            // a 'p' line writes nothing
        const bool head = (val.flags & 2u) != 0;
        const int64_t len_in = head ? 0 : ((excl.flags & 1u) ? excl.last_p : 0);   // streams begin with length 0 and stop 0
        const int64_t stop_in = head ? 0 : excl.sum;
        const int64_t st = (int64_t)((uint64_t)stop_in + (uint64_t)val.sum), sp = (int64_t)((uint64_t)st + (uint64_t)len_in);
        start[i] = st; stop[i] = sp;
        uint64_t q = s;
        while (q < e && tf[q] != '\t') q++;
        const uint32_t rem = q < e ? (uint32_t)(e - q) : 0;                          // the tab and what follows it
        auto dl = [](int64_t v) { uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v; int d = 1; while (a >= 10) { a /= 10; d++; } return (uint32_t)(d + (v < 0)); };
        out_len[i] = name_len[k] + 1 + dl(st) + 1 + dl(sp) + rem + 1;
    }
};
struct OutOffScan {
    typedef uint64_t T;
    const uint32_t *out_len; uint64_t *out_off;
    __host__ __device__ static T identity() { return 0; }
    __host__ __device__ static T op(T a, T b) { return a + b; }
    __device__ T load(uint64_t i) const { return out_len[i]; }
    __device__ void store(uint64_t i, T excl, T) const { out_off[i] = excl; }
};

__device__ __forceinline__ uint32_t put_dec64(uint8_t *dst, int64_t v)
{
    uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    uint32_t k = 0;
    if (v < 0) dst[k++] = '-';
    uint8_t tmp[20]; int nd = 0;
    do { tmp[nd++] = (uint8_t)('0' + a % 10); a /= 10; } while (a);
    for (int j = nd - 1; j >= 0; j--) dst[k++] = tmp[j];
    return k;
}

__global__ void __launch_bounds__(256) k_inv_write(const uint8_t *__restrict__ tf, const uint64_t *__restrict__ line_start, uint64_t n_lines,
                                                   const int64_t *start, const int64_t *stop, const uint32_t *out_len, const uint64_t *out_off,
                                                   const uint32_t *sid, const uint8_t *names, const uint64_t *name_off, const uint32_t *name_len,
                                                   uint8_t *bed)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines || out_len[i] == 0) return;
    const uint64_t s = line_start[i], e = line_start[i + 1] - 1;
    uint8_t *w = bed + out_off[i];
    const uint32_t k = sid[i], nl = name_len[k];
    const uint8_t *nm = names + name_off[k];
    for (uint32_t q = 0; q < nl; q++) *w++ = nm[q];
    *w++ = '\t';
    w += put_dec64(w, start[i]);
    *w++ = '\t';
    w += put_dec64(w, stop[i]);
    uint64_t q = s;
    while (q < e && tf[q] != '\t') q++;
    for (; q < e; q++) *w++ = tf[q];
    *w = '\n';
}

// ---- host side -----------------------------------------------------------------------------------------
struct StreamIn { uint64_t off, size; };          // inside the payload

// bzip2 streams -> their decoded bytes back to back in ctx->tf; out_off[s] = where stream s starts (n_streams + 1 entries)
static int decode_streams(Ctx *ctx, const uint8_t *d_z, uint64_t nz, const std::vector<StreamIn> &streams, const uint8_t *h_z,
                          std::vector<uint64_t> &out_off)
{
    const uint64_t n_streams = streams.size();
    out_off.assign(n_streams + 1, 0);
    if (n_streams == 0) return S3G_OK;
    // ---- block boundaries ----
    const uint64_t n_words = (nz + 3) / 4;
    uint32_t cap = (uint32_t)std::min<uint64_t>(nz / 32 + 4 * n_streams + 64, 1u << 26);
    S3G_TRY(ctx->io_c.ensure((uint64_t)cap * 8 + 64));
    S3G_TRY(ctx->scalars.ensure(64 * 8));
    uint32_t *d_ncand = reinterpret_cast<uint32_t *>(ctx->scalars.as<uint64_t>() + 20);
    S3G_CUDA(cudaMemsetAsync(d_ncand, 0, 8, ctx->stream));
    S3G_BYTES(ctx, nz);
    S3G_LAUNCH(ctx, k_find_magic, (unsigned)((n_words + 255) / 256), 256, 0, d_z, n_words, nz * 8, ctx->io_c.as<unsigned long long>(), d_ncand, cap);
    uint32_t n_cand = 0;
    S3G_CUDA(cudaMemcpyAsync(&n_cand, d_ncand, 4, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_cand > cap) { set_error("too many block-magic candidates"); return S3G_E_LIMIT; }
    std::vector<unsigned long long> cand(n_cand);
    if (n_cand) S3G_CUDA(cudaMemcpyAsync(cand.data(), ctx->io_c.p, (size_t)n_cand * 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    std::sort(cand.begin(), cand.end());
    // candidate blocks, each with the stream that holds it
    std::vector<DecBlock> blocks;
    std::vector<uint32_t> blk_stream;
    {
        size_t ci = 0;
        for (uint64_t s = 0; s < n_streams; s++) {
            const StreamIn &st = streams[s];
            if (st.size < 14 || st.off + st.size > nz) { set_error("stream %llu: bad offset / size", (unsigned long long)s); return S3G_E_PARAM; }
            const uint8_t *h = h_z + st.off;
            if (h[0] != 'B' || h[1] != 'Z' || h[2] != 'h' || h[3] < '1' || h[3] > '9') { set_error("stream %llu: not a bzip2 stream", (unsigned long long)s); return S3G_E_PARAM; }
            const uint64_t lo = st.off * 8, hi = (st.off + st.size) * 8;
            while (ci < cand.size() && (cand[ci] >> 1) < lo) ci++;
            for (; ci < cand.size() && (cand[ci] >> 1) < hi; ci++) {
                if (cand[ci] & 1) continue;
                DecBlock B;
                memset(&B, 0, sizeof B);
                B.bit_pos = cand[ci] >> 1; B.limit_bit = hi; B.nblock_max = 100000u * (uint32_t)(h[3] - '0'); B.status = DE_TRUNCATED;
                blocks.push_back(B); blk_stream.push_back((uint32_t)s);
            }
        }
    }
    const uint64_t nb = blocks.size();
    std::vector<uint32_t> chain;                 // the blocks that really are blocks, in stream order
    std::vector<uint64_t> first_of(n_streams + 1, 0);
    if (nb) {
        if (nb > 65535) { set_error("more than 65535 bzip2 blocks in one decode call"); return S3G_E_LIMIT; }
        S3G_TRY(ctx->blocks.ensure(nb * sizeof(DecBlock)));
        S3G_TRY(ctx->lcol.ensure(nb * (uint64_t)BLK_STRIDE));
        DecBlock *d_blocks = ctx->blocks.as<DecBlock>();
        S3G_CUDA(cudaMemcpyAsync(d_blocks, blocks.data(), nb * sizeof(DecBlock), cudaMemcpyHostToDevice, ctx->stream));
        S3G_BYTES(ctx, nz);
        S3G_LAUNCH(ctx, k_bz_decode, (unsigned)nb, 32, 0, d_z, d_blocks, ctx->lcol.as<uint8_t>());
        S3G_CUDA(cudaMemcpyAsync(blocks.data(), d_blocks, nb * sizeof(DecBlock), cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        S3G_TRY(check_launch("bz decode"));
    }
    // ---- chain the blocks of every stream: a block starts where the one before it ended ----
    auto bits_at = [&](uint64_t bit, int k) -> uint64_t {          // k <= 48 bits of the payload, host copy
        uint64_t v = 0;
        for (int i = 0; i < k; i++) { uint64_t p = bit + i; v = (v << 1) | ((h_z[p >> 3] >> (7 - (p & 7))) & 1u); }
        return v;
    };
    {
        size_t bi = 0;
        for (uint64_t s = 0; s < n_streams; s++) {
            first_of[s] = chain.size();
            const StreamIn &st = streams[s];
            uint64_t pos = st.off * 8 + 32;
            const uint64_t hi = (st.off + st.size) * 8;
            uint32_t combined = 0;
            for (;;) {
                if (pos + 48 > hi) { set_error("stream %llu: truncated", (unsigned long long)s); return S3G_E_PARAM; }
                const uint64_t magic = bits_at(pos, 48);
                if (magic == END_MAGIC) {
                    if (pos + 80 > hi) { set_error("stream %llu: truncated trailer", (unsigned long long)s); return S3G_E_PARAM; }
                    if ((uint32_t)bits_at(pos + 48, 32) != combined) { set_error("stream %llu: combined CRC mismatch", (unsigned long long)s); return S3G_E_PARAM; }
                    break;
                }
                if (magic != BLOCK_MAGIC) { set_error("stream %llu: no block magic at bit %llu", (unsigned long long)s, (unsigned long long)pos); return S3G_E_PARAM; }
                while (bi < nb && (blk_stream[bi] < s || (blk_stream[bi] == s && blocks[bi].bit_pos < pos))) bi++;
                if (bi >= nb || blk_stream[bi] != s || blocks[bi].bit_pos != pos) { set_error("stream %llu: block at bit %llu was not found by the scan", (unsigned long long)s, (unsigned long long)pos); return S3G_E_CUDA; }
                if (blocks[bi].status != DE_OK) { set_error("stream %llu: corrupt block at bit %llu (code %d)", (unsigned long long)s, (unsigned long long)pos, blocks[bi].status); return S3G_E_PARAM; }
                combined = ((combined << 1) | (combined >> 31)) ^ blocks[bi].crc;     // bz/decompress.c / bz/bzlib.c:607
                chain.push_back((uint32_t)bi);
                pos = blocks[bi].end_bit;
                bi++;
            }
        }
        first_of[n_streams] = chain.size();
    }
    // blocks that were only look-alikes of the magic inside other blocks' data do nothing from here on
    {
        std::vector<uint8_t> real(nb, 0);
        for (uint32_t b : chain) real[b] = 1;
        for (uint64_t b = 0; b < nb; b++) if (!real[b]) blocks[b].status = DE_TRUNCATED;
    }
    if (nb) {
        DecBlock *d_blocks = ctx->blocks.as<DecBlock>();
        S3G_CUDA(cudaMemcpyAsync(d_blocks, blocks.data(), nb * sizeof(DecBlock), cudaMemcpyHostToDevice, ctx->stream));
        // ---- inverse BWT ----
        S3G_TRY(ctx->sa.ensure(nb * (uint64_t)BLK_STRIDE * 4));
        S3G_TRY(ctx->kv0.ensure(nb * (uint64_t)256 * IB_T * 4));
        S3G_TRY(ctx->mtf0.ensure(nb * (uint64_t)BLK_STRIDE));
        S3G_TRY(ctx->io_d.ensure(nb * (uint64_t)(IBWT_K + 2) * 8));
        S3G_TRY(ctx->io_e.ensure(nb * (uint64_t)(IBWT_K + 2) * 4));
        double N = 0;
        for (uint32_t b : chain) N += blocks[b].nblock;
        S3G_BYTES(ctx, 10 * N);
        S3G_LAUNCH(ctx, k_ibwt_build, (unsigned)nb, IB_T, 0, d_blocks, ctx->lcol.as<uint8_t>(), ctx->sa.as<uint32_t>(), ctx->kv0.as<uint32_t>());
        dim3 wg((IBWT_K + 1 + 255) / 256, (unsigned)nb);
        S3G_BYTES(ctx, 4 * N);
        S3G_LAUNCH(ctx, k_ibwt_walk_a, wg, 256, 0, d_blocks, ctx->sa.as<uint32_t>(), ctx->io_d.as<uint2>());
        S3G_LAUNCH(ctx, k_ibwt_order, (unsigned)nb, 256, 0, d_blocks, ctx->io_d.as<uint2>(), ctx->io_e.as<uint32_t>());
        S3G_BYTES(ctx, 5 * N);
        S3G_LAUNCH(ctx, k_ibwt_walk_b, wg, 256, 0, d_blocks, ctx->sa.as<uint32_t>(), ctx->io_e.as<uint32_t>(), ctx->mtf0.as<uint8_t>());
        // ---- undo RLE1: sizes, then bytes ----
        S3G_BYTES(ctx, N);
        S3G_TRY(ctx->ztiles.ensure(nb * (uint64_t)UNRLE_S * 4));
        S3G_LAUNCH(ctx, k_unrle<false>, (unsigned)((nb * UNRLE_S + 3) / 4), 128, 0, d_blocks, (uint32_t)nb, ctx->mtf0.as<uint8_t>(), (uint8_t *)nullptr,
                   ctx->ztiles.as<uint32_t>());
        S3G_CUDA(cudaMemcpyAsync(blocks.data(), d_blocks, nb * sizeof(DecBlock), cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        S3G_TRY(check_launch("inverse bwt"));
    }
    uint64_t total = 0;
    for (uint64_t s = 0; s < n_streams; s++) {
        out_off[s] = total;
        for (uint64_t c = first_of[s]; c < first_of[s + 1]; c++) {
            DecBlock &B = blocks[chain[c]];
            if (B.cycle == 0 || B.nblock % B.cycle != 0) { set_error("stream %llu: corrupt block (inverse BWT cycle %u of %u)", (unsigned long long)s, B.cycle, B.nblock); return S3G_E_PARAM; }
            B.out_off = total; total += B.out_len;
        }
    }
    out_off[n_streams] = total;
    S3G_TRY(ctx->tf.ensure(total + 64));
    if (nb) {
        DecBlock *d_blocks = ctx->blocks.as<DecBlock>();
        S3G_CUDA(cudaMemcpyAsync(d_blocks, blocks.data(), nb * sizeof(DecBlock), cudaMemcpyHostToDevice, ctx->stream));
        S3G_BYTES(ctx, 2.0 * (double)total);
        S3G_LAUNCH(ctx, k_unrle<true>, (unsigned)((nb * UNRLE_S + 3) / 4), 128, 0, d_blocks, (uint32_t)nb, ctx->mtf0.as<uint8_t>(), ctx->tf.as<uint8_t>(),
                   ctx->ztiles.as<uint32_t>());
        // ---- block CRCs over the decoded bytes (bz/bzlib.c:843-846) ----
        std::vector<BlockInfo> bi(chain.size());
        for (size_t c = 0; c < chain.size(); c++) { memset(&bi[c], 0, sizeof(BlockInfo)); bi[c].in_start = blocks[chain[c]].out_off; bi[c].in_end = bi[c].in_start + blocks[chain[c]].out_len; }
        S3G_TRY(ctx->blk_prov.ensure(bi.size() * sizeof(BlockInfo) + 64));
        if (!bi.empty()) {
            S3G_CUDA(cudaMemcpyAsync(ctx->blk_prov.p, bi.data(), bi.size() * sizeof(BlockInfo), cudaMemcpyHostToDevice, ctx->stream));
            S3G_TRY(run_block_crc(ctx, ctx->tf.as<uint8_t>(), ctx->blk_prov.as<BlockInfo>(), bi.size(), (double)total));
            S3G_CUDA(cudaMemcpyAsync(bi.data(), ctx->blk_prov.p, bi.size() * sizeof(BlockInfo), cudaMemcpyDeviceToHost, ctx->stream));
        }
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        S3G_TRY(check_launch("undo rle1"));
        for (size_t c = 0; c < chain.size(); c++)
            if (bi[c].crc != blocks[chain[c]].crc) { set_error("block %zu: CRC mismatch (stored %08x, decoded %08x)", c, blocks[chain[c]].crc, bi[c].crc); return S3G_E_PARAM; }
    }
    ctx->last_dec_blocks = chain.size();
    return S3G_OK;
}

// transformed streams (in ctx->tf or any device buffer) -> BED text in ctx->streams; names on the host
static int inverse_transform(Ctx *ctx, const uint8_t *d_tf, uint64_t n, const std::vector<uint64_t> &soff, const std::vector<std::string> &names,
                             uint64_t *bed_len)
{
    *bed_len = 0;
    const uint64_t n_streams = names.size();
    if (n == 0 || n_streams == 0) return S3G_OK;
    const uint64_t ntiles = (n + IV_TILE - 1) / IV_TILE;
    S3G_TRY(ctx->tile_cnt.ensure((ntiles + 1) * 8));
    S3G_TRY(ctx->scalars.ensure(64 * 8));
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    uint64_t *tile_cnt = ctx->tile_cnt.as<uint64_t>();
    S3G_BYTES(ctx, n);
    S3G_LAUNCH(ctx, k_inv_count, (unsigned)ntiles, IV_T, 0, d_tf, n, tile_cnt);
    S3G_LAUNCH(ctx, k_scan_agg<SumU64>, 1, AGG_THREADS, 0, tile_cnt, ntiles, d_sc + 24);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 24, d_sc + 24, 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t n_lines = ctx->h_scalars[24];
    if (n_lines == 0) return S3G_OK;
    S3G_TRY(ctx->line_start.ensure((n_lines + 1) * 8));
    S3G_TRY(ctx->start.ensure(n_lines * 8));
    S3G_TRY(ctx->stop.ensure(n_lines * 8));
    S3G_TRY(ctx->rem_off.ensure(n_lines * 4));          // output length per line
    S3G_TRY(ctx->io_a.ensure(n_lines * 4));             // stream of the line
    S3G_TRY(ctx->io_b.ensure(n_lines * 8));             // output offset per line
    S3G_BYTES(ctx, n + 8 * n_lines);
    S3G_LAUNCH(ctx, k_inv_starts, (unsigned)ntiles, IV_T, 0, d_tf, n, tile_cnt, ctx->line_start.as<uint64_t>());
    // stream table and names on the device
    std::vector<uint64_t> name_off(n_streams + 1, 0);
    std::vector<uint32_t> name_len(n_streams);
    std::string all;
    for (uint64_t s = 0; s < n_streams; s++) { name_off[s] = all.size(); name_len[s] = (uint32_t)names[s].size(); all += names[s]; }
    name_off[n_streams] = all.size();
    S3G_TRY(ctx->soff.ensure((n_streams + 2) * 8));
    S3G_TRY(ctx->stream_tab.ensure((n_streams + 2) * 12 + all.size() + 64));
    uint64_t *d_name_off = ctx->stream_tab.as<uint64_t>();
    uint32_t *d_name_len = reinterpret_cast<uint32_t *>(d_name_off + n_streams + 1);
    uint8_t *d_names = reinterpret_cast<uint8_t *>(d_name_len + n_streams + 1);
    S3G_CUDA(cudaMemcpyAsync(ctx->soff.p, soff.data(), (n_streams + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaMemcpyAsync(d_name_off, name_off.data(), (n_streams + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaMemcpyAsync(d_name_len, name_len.data(), n_streams * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!all.empty()) S3G_CUDA(cudaMemcpyAsync(d_names, all.data(), all.size(), cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));      // the host vectors above go out of scope
    const uint64_t stiles = (n_lines + SCAN_TILE - 1) / SCAN_TILE + 1;
    S3G_TRY(ctx->scan_a.ensure(stiles * sizeof(InvAgg)));
    S3G_TRY(ctx->scan_b.ensure(stiles * 8));
    InvScan f1;
    f1.tf = d_tf; f1.line_start = ctx->line_start.as<uint64_t>(); f1.soff = ctx->soff.as<uint64_t>(); f1.n_streams = n_streams;
    f1.start = ctx->start.as<int64_t>(); f1.stop = ctx->stop.as<int64_t>(); f1.out_len = ctx->rem_off.as<uint32_t>();
    f1.sid_out = ctx->io_a.as<uint32_t>(); f1.name_len = d_name_len;
    S3G_TRY(device_scan(ctx, f1, n_lines, ctx->scan_a.as<InvAgg>(), (InvAgg *)nullptr));
    OutOffScan f2; f2.out_len = ctx->rem_off.as<uint32_t>(); f2.out_off = ctx->io_b.as<uint64_t>();
    S3G_TRY(device_scan(ctx, f2, n_lines, ctx->scan_b.as<uint64_t>(), d_sc + 25));
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 25, d_sc + 25, 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t total = ctx->h_scalars[25];
    S3G_TRY(ctx->streams.ensure(total + 64));
    S3G_BYTES(ctx, n + total);
    S3G_LAUNCH(ctx, k_inv_write, (unsigned)((n_lines + 255) / 256), 256, 0, d_tf, ctx->line_start.as<uint64_t>(), n_lines, ctx->start.as<int64_t>(),
               ctx->stop.as<int64_t>(), ctx->rem_off.as<uint32_t>(), ctx->io_b.as<uint64_t>(), ctx->io_a.as<uint32_t>(), d_names, d_name_off,
               d_name_len, ctx->streams.as<uint8_t>());
    *bed_len = total;
    return check_launch("inverse transform");
}

// upload a host buffer, zero-padded so that the bit readers may run a few words past the end
static int stage_padded(Ctx *ctx, DevBuf &buf, const void *src, uint64_t n)
{
    S3G_TRY(buf.ensure(n + DEC_PAD));
    if (n) S3G_CUDA(cudaMemcpyAsync(buf.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaMemsetAsync((uint8_t *)buf.p + n, 0, DEC_PAD, ctx->stream));
    return S3G_OK;
}

// ---- the archive's metadata line (ARCHIVE_FORMAT.md): the few fields the decoder needs -------------------
static bool json_string_at(const std::string &m, size_t p, std::string *out, size_t *endp)
{
    if (p >= m.size() || m[p] != '"') return false;
    out->clear();
    for (p++; p < m.size(); p++) {
        char c = m[p];
        if (c == '"') { *endp = p + 1; return true; }
        if (c != '\\') { out->push_back(c); continue; }
        if (++p >= m.size()) return false;
        switch (m[p]) {
            case 'b': out->push_back('\b'); break; case 'f': out->push_back('\f'); break; case 'n': out->push_back('\n'); break;
            case 'r': out->push_back('\r'); break; case 't': out->push_back('\t'); break;
            case 'u': {
                if (p + 4 >= m.size()) return false;
                unsigned v = 0;
                for (int k = 1; k <= 4; k++) { char h = m[p + k]; v = v * 16 + (h <= '9' ? h - '0' : (h | 32) - 'a' + 10); }
                out->push_back((char)v);          // the writer only escapes control characters this way
                p += 4; break;
            }
            default: out->push_back(m[p]);
        }
    }
    return false;
}
static bool json_uint_after(const std::string &m, size_t from, size_t to, const char *key, uint64_t *v)
{
    size_t p = m.find(key, from);
    if (p == std::string::npos || p >= to) return false;
    p += strlen(key);
    uint64_t a = 0; bool any = false;
    while (p < m.size() && m[p] >= '0' && m[p] <= '9') { a = a * 10 + (uint64_t)(m[p] - '0'); p++; any = true; }
    *v = a;
    return any;
}

struct ArcStream { std::string name; uint64_t off, size, lines, tf_bytes; };
static int parse_metadata(const uint8_t *arc, uint64_t n, std::vector<ArcStream> &out, uint64_t *payload_off)
{
    static const uint8_t magic[4] = {0xca, 0x5c, 0xad, 0x1a};      // hpp:907-910
    if (n < 5 || memcmp(arc, magic, 4) != 0) { set_error("not a starch3 archive (magic)"); return S3G_E_PARAM; }
    const void *nl = memchr(arc + 4, '\n', n - 4);
    if (!nl) { set_error("archive metadata is not terminated"); return S3G_E_PARAM; }
    const size_t hlen = (const uint8_t *)nl - (arc + 4);
    *payload_off = 4 + hlen + 1;
    std::string m(reinterpret_cast<const char *>(arc + 4), hlen);
    size_t p = m.find("\"streams\":[");
    if (p == std::string::npos) { set_error("archive metadata has no stream table"); return S3G_E_PARAM; }
    p += 11;
    while (p < m.size() && m[p] == '{') {
        const char *k = "\"chromosome\":";
        if (m.compare(p + 1, strlen(k), k) != 0) { set_error("archive metadata: unexpected stream entry"); return S3G_E_PARAM; }
        ArcStream s;
        size_t q;
        if (!json_string_at(m, p + 1 + strlen(k), &s.name, &q)) { set_error("archive metadata: bad chromosome name"); return S3G_E_PARAM; }
        size_t end = m.find('}', q);
        if (end == std::string::npos) { set_error("archive metadata: unterminated stream entry"); return S3G_E_PARAM; }
        if (!json_uint_after(m, q, end, "\"offset\":", &s.off) || !json_uint_after(m, q, end, "\"size\":", &s.size) ||
            !json_uint_after(m, q, end, "\"lines\":", &s.lines) || !json_uint_after(m, q, end, "\"transformedBytes\":", &s.tf_bytes)) {
            set_error("archive metadata: stream entry lacks offset / size / lines / transformedBytes"); return S3G_E_PARAM;
        }
        out.push_back(s);
        p = end + 1;
        if (p < m.size() && m[p] == ',') p++;
    }
    return S3G_OK;
}

}  // namespace s3g

using namespace s3g;

extern "C" {

int s3g_bz_decompress(s3g_ctx *ctx, const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_len)
{
    if (!ctx || !in || !out_len) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_TRY(stage_padded(ctx, ctx->io_a, in, n));
    std::vector<StreamIn> st(1);
    st[0].off = 0; st[0].size = n;
    std::vector<uint64_t> off;
    S3G_TRY(decode_streams(ctx, ctx->io_a.as<uint8_t>(), n, st, in, off));
    *out_len = off[1];
    if (off[1] > out_cap) { set_error("out_cap too small: need %llu", (unsigned long long)off[1]); return S3G_E_CAPACITY; }
    if (out && off[1]) S3G_CUDA(cudaMemcpyAsync(out, ctx->tf.p, off[1], cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

int s3g_inverse_transform(s3g_ctx *ctx, const uint8_t *tf, uint64_t n, const uint8_t *name, uint32_t name_len, uint8_t *bed, uint64_t bed_cap,
                          uint64_t *bed_len)
{
    if (!ctx || (!tf && n) || !bed_len || (!name && name_len)) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    if (n && tf[n - 1] != '\n') { set_error("a transformed stream ends with a line feed"); return S3G_E_PARAM; }
    S3G_TRY(stage_padded(ctx, ctx->io_c, tf, n));
    std::vector<uint64_t> soff = {0, n};
    std::vector<std::string> names(1, std::string(reinterpret_cast<const char *>(name), name_len));
    S3G_TRY(inverse_transform(ctx, ctx->io_c.as<uint8_t>(), n, soff, names, bed_len));
    if (*bed_len > bed_cap) { set_error("bed_cap too small: need %llu", (unsigned long long)*bed_len); return S3G_E_CAPACITY; }
    if (bed && *bed_len) S3G_CUDA(cudaMemcpyAsync(bed, ctx->streams.p, *bed_len, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

int s3g_decompress_archive(s3g_ctx *ctx, const uint8_t *archive, uint64_t n, uint8_t *bed, uint64_t bed_cap, uint64_t *bed_len, s3g_decode_info *info)
{
    if (!ctx || !archive || !bed_len) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    *bed_len = 0;
    if (info) memset(info, 0, sizeof *info);
    std::vector<ArcStream> as;
    uint64_t payload = 0;
    S3G_TRY(parse_metadata(archive, n, as, &payload));
    const uint64_t nz = n - payload;
    S3G_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    S3G_TRY(stage_padded(ctx, ctx->io_a, archive + payload, nz));
    std::vector<StreamIn> st(as.size());
    std::vector<std::string> names(as.size());
    for (size_t s = 0; s < as.size(); s++) { st[s].off = as[s].off; st[s].size = as[s].size; names[s] = as[s].name; }
    std::vector<uint64_t> off;
    S3G_TRY(decode_streams(ctx, ctx->io_a.as<uint8_t>(), nz, st, archive + payload, off));
    for (size_t s = 0; s < as.size(); s++)
        if (off[s + 1] - off[s] != as[s].tf_bytes) {
            set_error("stream %zu decodes to %llu bytes, the metadata says %llu", s, (unsigned long long)(off[s + 1] - off[s]), (unsigned long long)as[s].tf_bytes);
            return S3G_E_PARAM;
        }
    S3G_TRY(inverse_transform(ctx, ctx->tf.as<uint8_t>(), off.empty() ? 0 : off.back(), off, names, bed_len));
    S3G_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if (info) { info->n_streams = as.size(); info->n_blocks = ctx->last_dec_blocks; info->tf_bytes = off.empty() ? 0 : off.back(); info->d_bed = ctx->streams.p; }
    if (*bed_len > bed_cap && bed) { set_error("bed_cap too small: need %llu", (unsigned long long)*bed_len); return S3G_E_CAPACITY; }
    if (bed && *bed_len) S3G_CUDA(cudaMemcpyAsync(bed, ctx->streams.p, *bed_len, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    if (info) { float ms = 0; if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) info->device_ms = ms; else cudaGetLastError(); }
    return S3G_OK;
}

}  // extern "C"
