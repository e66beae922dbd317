// mtf_huff.cu -- kernels (3c) and (3d): move-to-front + RUNA/RUNB zero-run coding
// (generateMTFValues, bz/compress.c:120-231) and bzip2's iterative Huffman table
// selection and emission (sendMTFValues, bz/compress.c:239-598; BZ2_hbMakeCodeLengths
// / BZ2_hbAssignCodes, bz/huffman.c:63-166).  One CTA per bzip2 block when a batch holds enough blocks to fill the
// SMs; a batch of fewer blocks spreads every block over a thread-block CLUSTER of 2, 4 or 8 CTAs that exchange
// what they need through distributed shared memory (the per-symbol last occurrences of the MTF chunks; the symbol
// frequencies, selector history and bit counts of the Huffman groups).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace s3g {

// =============================================================================
// (3c) MTF.  The MTF position of symbol s at index i equals the number of
// symbols whose most recent occurrence is later than the most recent occurrence
// of s ("recency rank").  Symbols not seen yet get the virtual position -(c+1),
// which reproduces the initial list 0,1,2,... (bz/compress.c:161).  With that,
// a block splits into 32 chunks that only need the last occurrence of every
// symbol before the chunk start.
// =============================================================================
constexpr int MT = 1024;                 // threads per CTA (32 warps = 32 chunks)
constexpr int UNSET = INT32_MIN;

template <int NS>
__device__ __forceinline__ void mtf_chunk(const uint8_t *L, uint8_t *M, int beg, int end, const int *init_last, int a)
{
    const unsigned l = threadIdx.x & 31;
    int lastv[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        int c = s * 32 + (int)l;
        lastv[s] = c < a ? init_last[c] : UNSET;
    }
    for (int base = beg; base < end; base += 32) {
        int pos = base + (int)l;
        int mysym = pos < end ? L[pos] : 0;
        int mym = 0;
        int lim = end - base < 32 ? end - base : 32;
        for (int t = 0; t < lim; t++) {
            int s = __shfl_sync(0xffffffffu, mysym, t);
            int slot = s >> 5, ln = s & 31;
            int v = lastv[0];
#pragma unroll
            for (int q = 1; q < NS; q++) if (slot == q) v = lastv[q];
            int ls = __shfl_sync(0xffffffffu, v, ln);
            int cnt = 0;
#pragma unroll
            for (int q = 0; q < NS; q++) cnt += __popc(__ballot_sync(0xffffffffu, lastv[q] > ls));
            if ((int)l == t) mym = cnt;
            if ((int)l == ln) {
#pragma unroll
                for (int q = 0; q < NS; q++) if (slot == q) lastv[q] = base + t;
            }
        }
        if (pos < end) M[pos] = (uint8_t)mym;
    }
}

// zero runs in bijective base 2 (RUNA/RUNB, bz/compress.c:174-188), other ranks +1, then EOB; mtfFreq.
// Called by all THREADS threads of the CTA after the ranks M[0..n) are visible.
template <int THREADS>
__device__ __forceinline__ void mtf_zero_runs(const uint8_t *M, uint16_t *mtfv, int n, int a, int *s_freq, uint32_t *s_scan,
                                              BlockInfo *blocks, uint32_t lb, int32_t *freq_all)
{
    constexpr int DI = 8;
    uint32_t out_base = 0;       // symbols written so far
    uint32_t nz_carry = 0;       // (position + 1) of the last non-zero rank so far
    int fa = 0, fb = 0;          // private RUNA / RUNB counts
    // the ranks of the next tile are requested while this one goes through its two block scans
    uint64_t v_next = 0; uint8_t e_next = 1;
    {
        int q0 = (int)threadIdx.x * DI;
        if (q0 + DI < n) { v_next = *reinterpret_cast<const uint64_t *>(M + q0); e_next = M[q0 + DI]; }
    }
    for (int tile0 = 0; tile0 < n; tile0 += THREADS * DI) {
        int p0 = tile0 + (int)threadIdx.x * DI;
        uint8_t m[DI + 1];
        if (p0 + DI < n) {
            uint64_t v = v_next;                                           // slots are 128-byte aligned, p0 % 8 == 0
#pragma unroll
            for (int k = 0; k < DI; k++) m[k] = (uint8_t)(v >> (8 * k));
            m[DI] = e_next;
        } else {
#pragma unroll
            for (int k = 0; k <= DI; k++) m[k] = (p0 + k < n) ? M[p0 + k] : (uint8_t)1;   // past the end acts as non-zero
        }
        {
            int q0 = p0 + THREADS * DI;
            if (q0 + DI < n) { v_next = *reinterpret_cast<const uint64_t *>(M + q0); e_next = M[q0 + DI]; }
        }
        uint32_t lastnz = 0;
#pragma unroll
        for (int k = 0; k < DI; k++) if (p0 + k < n && m[k]) lastnz = (uint32_t)(p0 + k + 1);
        uint32_t tmax;
        uint32_t ex = block_excl_max<uint32_t>(lastnz, s_scan, &tmax);
        uint32_t rs = ex > nz_carry ? ex : nz_carry;     // the zero run in effect starts at position rs
        uint32_t emit = 0;
        uint32_t rs_k = rs;
#pragma unroll
        for (int k = 0; k < DI; k++) {
            if (p0 + k < n) {
                if (m[k]) { emit++; rs_k = (uint32_t)(p0 + k + 1); }
                else if (m[k + 1]) { uint32_t r = (uint32_t)(p0 + k + 1) - rs_k; emit += 31 - __clz(r + 1); }
            }
        }
        uint32_t ttot;
        uint32_t eo = block_excl_sum<uint32_t>(emit, s_scan, &ttot);
        uint32_t o = out_base + eo;
        rs_k = rs;
#pragma unroll
        for (int k = 0; k < DI; k++) {
            if (p0 + k < n) {
                if (m[k]) { mtfv[o++] = (uint16_t)(m[k] + 1); atomicAdd(&s_freq[m[k] + 1], 1); rs_k = (uint32_t)(p0 + k + 1); }
                else if (m[k + 1]) {
                    uint32_t zp = (uint32_t)(p0 + k + 1) - rs_k - 1;
                    for (;;) {
                        if (zp & 1) { mtfv[o++] = 1; fb++; } else { mtfv[o++] = 0; fa++; }
                        if (zp < 2) break;
                        zp = (zp - 2) >> 1;
                    }
                }
            }
        }
        out_base += ttot;
        if (tmax > nz_carry) nz_carry = tmax;
    }
    if (fa) atomicAdd(&s_freq[0], fa);
    if (fb) atomicAdd(&s_freq[1], fb);
    __syncthreads();
    if (threadIdx.x == 0) {
        mtfv[out_base] = (uint16_t)(a + 1);      // EOB
        s_freq[a + 1] += 1;
        blocks[lb].n_mtf = out_base + 1;
    }
    __syncthreads();
    int32_t *fq = freq_all + (uint64_t)lb * 258;
    for (int i = threadIdx.x; i < 258; i += THREADS) fq[i] = s_freq[i];
}

// The same coding, tile-parallel, for batches that cannot fill the SMs with one CTA per block (k_mtf* with
// zero_runs = 0 leave the ranks in M[]): a run's digits are written where the run ENDS, with its start = the position
// after the last non-zero rank before it.  Inside a tile that start is known except for the first run end when no
// non-zero rank precedes it in the tile; so
//   k_zrun_tiles    per tile: position of its last non-zero rank, its symbol count without that one open run,
//                   and where that run ends;
//   k_zrun_offsets  per block: exclusive max-scan (carry of the last non-zero position) and sum-scan (output
//                   offsets) over its tiles, EOB, nMTF;
//   k_zrun_emit     per tile: the symbols (put together in shared memory, written as one piece) and the frequencies.
constexpr int ZT = 256;                  // threads per tile
constexpr int ZDI = 16;                  // ranks per thread
constexpr int ZTILE = ZT * ZDI;          // ranks per tile
constexpr int ZNT = (BLK_STRIDE + ZTILE - 1) / ZTILE;

struct ZTileInfo { uint32_t lastnz, base, open_end, off; };       // per block and tile; lastnz becomes the carry, off is filled by k_zrun_offsets

// ranks of one thread: m[0..ZDI) and the rank after them (positions past the block end act as non-zero)
__device__ __forceinline__ void zrun_load(const uint8_t *M, int p0, int n, uint8_t (&m)[ZDI + 1])
{
    if (p0 + ZDI < n) {
#pragma unroll
        for (int q = 0; q < ZDI / 16; q++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(M + p0 + 16 * q);      // slots are 128-byte aligned, p0 % 16 == 0
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; k++) m[16 * q + k] = (uint8_t)(w4[k >> 2] >> (8 * (k & 3)));
        }
        m[ZDI] = M[p0 + ZDI];
    } else {
#pragma unroll
        for (int k = 0; k <= ZDI; k++) m[k] = (p0 + k < n) ? M[p0 + k] : (uint8_t)1;
    }
}

__global__ void __launch_bounds__(ZT) k_zrun_tiles(const uint8_t *mtf0, const BlockInfo *blocks, ZTileInfo *tiles)
{
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_open;
    const uint32_t lb = blockIdx.y, tile = blockIdx.x;
    const int n = (int)blocks[lb].nblock;
    const int tile0 = (int)tile * ZTILE;
    if (tile0 >= n) return;
    const uint8_t *M = mtf0 + (uint64_t)lb * BLK_STRIDE;
    if (threadIdx.x == 0) s_open = 0;
    const int p0 = tile0 + (int)threadIdx.x * ZDI;
    uint8_t m[ZDI + 1];
    zrun_load(M, p0, n, m);
    uint32_t lastnz = 0;
#pragma unroll
    for (int k = 0; k < ZDI; k++) if (p0 + k < n && m[k]) lastnz = (uint32_t)(p0 + k + 1);
    uint32_t tmax;
    uint32_t ex = block_excl_max<uint32_t>(lastnz, s_scan, &tmax);      // position + 1 of the last non-zero rank before mine, in this tile
    uint32_t emit = 0, rs_k = ex;
#pragma unroll
    for (int k = 0; k < ZDI; k++) {
        if (p0 + k < n) {
            if (m[k]) { emit++; rs_k = (uint32_t)(p0 + k + 1); }
            else if (m[k + 1]) {
                if (rs_k) { uint32_t r = (uint32_t)(p0 + k + 1) - rs_k; emit += 31 - __clz(r + 1); }
                else s_open = (uint32_t)(p0 + k + 1);          // the one run end whose start lies before the tile
            }
        }
    }
    uint32_t ttot;
    block_excl_sum<uint32_t>(emit, s_scan, &ttot);
    if (threadIdx.x == 0) {
        ZTileInfo t; t.lastnz = tmax; t.base = ttot; t.open_end = s_open; t.off = 0;
        tiles[(uint64_t)lb * ZNT + tile] = t;
    }
}

__global__ void __launch_bounds__(256) k_zrun_offsets(ZTileInfo *tiles_all, uint16_t *mtfv_all, int32_t *freq_all, BlockInfo *blocks)
{
    __shared__ uint32_t s_scan[33];
    static_assert(ZNT <= 256, "one thread per tile");
    const uint32_t lb = blockIdx.x, t = threadIdx.x;
    const int n = (int)blocks[lb].nblock;
    const int a = (int)blocks[lb].n_in_use;
    const uint32_t nt = (uint32_t)((n + ZTILE - 1) / ZTILE);
    ZTileInfo *tiles = tiles_all + (uint64_t)lb * ZNT;
    ZTileInfo ti; ti.lastnz = 0; ti.base = 0; ti.open_end = 0; ti.off = 0;
    if (t < nt) ti = tiles[t];
    uint32_t tmax, ttot;
    uint32_t carry = block_excl_max<uint32_t>(ti.lastnz, s_scan, &tmax);     // last non-zero position + 1 in the earlier tiles
    uint32_t cnt = ti.base;
    if (ti.open_end) { uint32_t r = ti.open_end - carry; cnt += 31 - __clz(r + 1); }
    uint32_t off = block_excl_sum<uint32_t>(cnt, s_scan, &ttot);
    if (t < nt) { tiles[t].off = off; tiles[t].lastnz = carry; }
    if (t == 0) {
        mtfv_all[(uint64_t)lb * BLK_STRIDE + ttot] = (uint16_t)(a + 1);       // EOB
        atomicAdd(&freq_all[(uint64_t)lb * 258 + a + 1], 1);
        blocks[lb].n_mtf = ttot + 1;
    }
}

__global__ void __launch_bounds__(ZT) k_zrun_emit(const uint8_t *mtf0, const BlockInfo *blocks, const ZTileInfo *tiles, uint16_t *mtfv_all,
                                                 int32_t *freq_all)
{
    __shared__ uint32_t s_scan[33];
    __shared__ int s_freq[258];
    __shared__ uint16_t s_out[ZTILE + 32];      // at most one symbol per rank, except that the one open run gives up to 20 digits
    const uint32_t lb = blockIdx.y, tile = blockIdx.x;
    const int n = (int)blocks[lb].nblock;
    const int tile0 = (int)tile * ZTILE;
    if (tile0 >= n) return;
    const uint8_t *M = mtf0 + (uint64_t)lb * BLK_STRIDE;
    uint16_t *mtfv = mtfv_all + (uint64_t)lb * BLK_STRIDE;
    for (int i = threadIdx.x; i < 258; i += ZT) s_freq[i] = 0;
    const ZTileInfo ti = tiles[(uint64_t)lb * ZNT + tile];
    const int p0 = tile0 + (int)threadIdx.x * ZDI;
    uint8_t m[ZDI + 1];
    zrun_load(M, p0, n, m);
    uint32_t lastnz = 0;
#pragma unroll
    for (int k = 0; k < ZDI; k++) if (p0 + k < n && m[k]) lastnz = (uint32_t)(p0 + k + 1);
    uint32_t tmax;
    uint32_t ex = block_excl_max<uint32_t>(lastnz, s_scan, &tmax);
    const uint32_t rs = ex ? ex : ti.lastnz;                   // the zero run in effect starts at position rs
    uint32_t emit = 0, rs_k = rs;
#pragma unroll
    for (int k = 0; k < ZDI; k++) {
        if (p0 + k < n) {
            if (m[k]) { emit++; rs_k = (uint32_t)(p0 + k + 1); }
            else if (m[k + 1]) { uint32_t r = (uint32_t)(p0 + k + 1) - rs_k; emit += 31 - __clz(r + 1); }
        }
    }
    uint32_t ttot;
    uint32_t o = block_excl_sum<uint32_t>(emit, s_scan, &ttot);          // also orders the s_freq zeroing before the atomics
    int fa = 0, fb = 0;
    rs_k = rs;
#pragma unroll
    for (int k = 0; k < ZDI; k++) {
        if (p0 + k < n) {
            if (m[k]) { s_out[o++] = (uint16_t)(m[k] + 1); atomicAdd(&s_freq[m[k] + 1], 1); rs_k = (uint32_t)(p0 + k + 1); }
            else if (m[k + 1]) {
                uint32_t zp = (uint32_t)(p0 + k + 1) - rs_k - 1;
                for (;;) {
                    if (zp & 1) { s_out[o++] = 1; fb++; } else { s_out[o++] = 0; fa++; }
                    if (zp < 2) break;
                    zp = (zp - 2) >> 1;
                }
            }
        }
    }
    if (fa) atomicAdd(&s_freq[0], fa);
    if (fb) atomicAdd(&s_freq[1], fb);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < ttot; i += ZT) mtfv[ti.off + i] = s_out[i];
    int32_t *fq = freq_all + (uint64_t)lb * 258;
    for (int i = threadIdx.x; i < 258; i += ZT) { int v = s_freq[i]; if (v) atomicAdd(&fq[i], v); }
}

// =============================================================================
// (3c, alphabets of at most 32 symbols -- every BED-derived stream) MTF with the list held in
// registers: 8 entries per 64-bit word, position found with a zero-byte test, move-to-front as a
// masked one-byte shift.  A block is cut into 512 chunks, one per thread; the list a chunk
// starts with is the symbols ordered by their last occurrence before the chunk.
// =============================================================================
constexpr int MS_BIG = 512;             // chunks (threads) per block, lists of up to 12 words
constexpr int MS_SMALL = 768;           // alphabets <= 24: a list of one or two words needs few registers, so two such CTAs fit an SM
constexpr int MTF_REG_MAX = 96;         // largest alphabet of the register-list kernel (12 list words; 96 x 512 x 4 B of shared memory)

// List entries are FB-bit fields, FPW = 64 / FB per word (FB = 4 for alphabets <= 16: the whole list in one word;
// FB = 5 for alphabets <= 24: two words; bytes otherwise).  Unused fields hold all ones, which no symbol equals.
template <int FB> struct MtfPack {
    static constexpr int FPW = 64 / FB;
    static constexpr uint64_t WMASK = FPW * FB == 64 ? ~0ull : (1ull << (FPW * FB % 64)) - 1;
    static constexpr uint64_t FMASK = (1ull << FB) - 1;
    static constexpr uint64_t ones() { uint64_t k = 0; for (int i = 0; i < FPW; i++) k |= 1ull << (FB * i); return k; }
    static constexpr uint64_t K1 = ones();
    static constexpr uint64_t KH = K1 << (FB - 1);
};

template <int FB, int NW>
__device__ __forceinline__ uint32_t mtf_step(uint64_t (&lst)[NW], uint32_t s)
{
    typedef MtfPack<FB> PK;
    uint64_t pat = (uint64_t)s * PK::K1;
    int qh = 0; uint64_t zh = 0;
#pragma unroll
    for (int q = NW - 1; q >= 0; q--) {
        uint64_t x = lst[q] ^ pat;
        uint64_t z = (x - PK::K1) & ~x & PK::KH;      // lowest set bit marks the first zero field of x
        if (z) { qh = q; zh = z; }
    }
    uint32_t bit = (uint32_t)(__ffsll((long long)zh) - 1);
    uint32_t bpos = FB == 8 ? bit >> 3 : FB == 4 ? bit >> 2 : (bit * 13u) >> 6;        // bit / FB (bit < 64)
    uint64_t inmask = (bpos + 1) * FB >= 64 ? ~0ull : ((1ull << ((bpos + 1) * FB)) - 1);
    uint64_t carry = s;
#pragma unroll
    for (int q = 0; q < NW; q++) {
        uint64_t old = lst[q];
        uint64_t sh = ((old << FB) | carry) & PK::WMASK;
        carry = (old >> (FB * (PK::FPW - 1))) & PK::FMASK;
        uint64_t m = q < qh ? PK::WMASK : (q == qh ? inmask : 0ull);
        lst[q] = (old & ~m) | (sh & m);
    }
    return (uint32_t)qh * PK::FPW + bpos;
}

template <int FB, int NW, int MS>
__device__ __forceinline__ void mtf_thread_chunk(const uint8_t *L, uint8_t *M, int beg, int end, const int *s_last, int a)
{
    typedef MtfPack<FB> PK;
    // initial list: symbols by decreasing last occurrence (virtual negative positions keep 0,1,2,.. for unseen ones)
    uint64_t lst[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) lst[q] = PK::WMASK;
    for (int c = 0; c < a; c++) {
        int v = s_last[c * MS + threadIdx.x];
        int rank = 0;
        for (int c2 = 0; c2 < a; c2++) rank += s_last[c2 * MS + threadIdx.x] > v ? 1 : 0;
        const int rq = rank / PK::FPW, rf = rank % PK::FPW;
#pragma unroll
        for (int q = 0; q < NW; q++)
            if (rq == q) lst[q] = (lst[q] & ~(PK::FMASK << (FB * rf))) | ((uint64_t)c << (FB * rf));
    }
    for (int p = beg; p < end; p += 8) {
        uint64_t in8 = *reinterpret_cast<const uint64_t *>(L + p);          // beg % 8 == 0, slots padded
        uint64_t out8 = 0;
        int lim = end - p < 8 ? end - p : 8;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < lim) {
                uint32_t pos = mtf_step<FB, NW>(lst, (uint32_t)(in8 >> (8 * k)) & 0xffu);
                out8 |= (uint64_t)pos << (8 * k);
            }
        }
        *reinterpret_cast<uint64_t *>(M + p) = out8;
    }
}

template <int MS, bool SMALL>
__global__ void __launch_bounds__(MS, 2) k_mtf_small(const uint8_t *lcol, uint8_t *mtf0, uint16_t *mtfv_all, int32_t *freq_all,
                                                     BlockInfo *blocks, int rows, int zero_runs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *s_last = reinterpret_cast<int *>(smem_raw);            // [rows][MS], rows = largest alphabet (<= MTF_REG_MAX) in the batch
    int *s_freq = s_last + rows * MS;                           // [258]
    uint32_t *s_scan = reinterpret_cast<uint32_t *>(s_freq + 260);
    int *s_pub = reinterpret_cast<int *>(s_scan + 40);          // [256] cluster form: last occurrence of every symbol in this CTA's chunks
    // grid = (cs, blocks): the cs CTAs of a cluster share one bzip2 block, CTA `rank` takes chunks rank * MS .. of its MS * cs
    const uint32_t lb = blockIdx.y;
    const int rank = (int)blockIdx.x, cs = (int)gridDim.x;
    const int n = (int)blocks[lb].nblock;
    const int a = (int)blocks[lb].n_in_use;
    if (a > MTF_REG_MAX || (a <= 24) != SMALL) return;          // handled by k_mtf or by the other instantiation
    if (a > rows) __trap();                                     // host mirror out of date: fail loudly
    const uint8_t *L = lcol + (uint64_t)lb * BLK_STRIDE;
    uint8_t *M = mtf0 + (uint64_t)lb * BLK_STRIDE;
    uint16_t *mtfv = mtfv_all + (uint64_t)lb * BLK_STRIDE;
    const int tid = threadIdx.x;
    for (int i = tid; i < a * MS; i += MS) s_last[i] = 0;
    for (int i = tid; i < 258; i += MS) s_freq[i] = 0;
    __syncthreads();
    int chunk = (((n + MS * cs - 1) / (MS * cs)) + 7) & ~7;
    long long beg_ll = (long long)(rank * MS + tid) * chunk;
    int beg = beg_ll > n ? n : (int)beg_ll, end = beg_ll + chunk > n ? n : (int)(beg_ll + chunk);
    // phase A: last occurrence (position + 1) of every symbol inside my chunk: a forward sweep of
    // fire-and-forget shared-memory stores (bank = tid, conflict-free), 8 bytes per load
    for (int p = beg; p < end; p += 8) {
        uint64_t in8 = *reinterpret_cast<const uint64_t *>(L + p);          // beg % 8 == 0, slots padded
        int lim = end - p < 8 ? end - p : 8;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < lim) s_last[((uint32_t)(in8 >> (8 * k)) & 0xffu) * MS + tid] = p + k + 1;
    }
    __syncthreads();
    // phase B: exclusive "latest occurrence" over the chunks; in the cluster form the chunks of the CTAs before this one
    // come first: every CTA publishes where it last saw each symbol, and reads the nearest earlier CTA that saw it
    if (cs > 1) {
        if (tid < a) {
            int last = 0;
            for (int ch = 0; ch < MS; ch++) { int t = s_last[tid * MS + ch]; if (t) last = t; }
            s_pub[tid] = last;
        }
        cg::this_cluster().sync();
    }
    if (tid < a) {
        int run = -(tid + 1);
        if (cs > 1)
            for (int r = rank - 1; r >= 0; r--) {
                int v = cg::this_cluster().map_shared_rank(s_pub, r)[tid];
                if (v) { run = v; break; }
            }
        for (int ch = 0; ch < MS; ch++) {
            int t = s_last[tid * MS + ch];
            s_last[tid * MS + ch] = run;
            if (t) run = t;
        }
    }
    if (cs > 1) cg::this_cluster().sync();      // no CTA leaves while its table may still be read
    __syncthreads();
    // phase C: ranks
    if (beg < end) {
        if (SMALL) {
            if (a <= 16) mtf_thread_chunk<4, 1, MS>(L, M, beg, end, s_last, a);
            else mtf_thread_chunk<5, 2, MS>(L, M, beg, end, s_last, a);
        } else switch ((a + 7) >> 3) {
            case 4: mtf_thread_chunk<8, 4, MS>(L, M, beg, end, s_last, a); break;
            case 5: mtf_thread_chunk<8, 5, MS>(L, M, beg, end, s_last, a); break;
            case 6: mtf_thread_chunk<8, 6, MS>(L, M, beg, end, s_last, a); break;
            case 7: mtf_thread_chunk<8, 7, MS>(L, M, beg, end, s_last, a); break;
            case 8: mtf_thread_chunk<8, 8, MS>(L, M, beg, end, s_last, a); break;
            case 9: mtf_thread_chunk<8, 9, MS>(L, M, beg, end, s_last, a); break;
            case 10: mtf_thread_chunk<8, 10, MS>(L, M, beg, end, s_last, a); break;
            case 11: mtf_thread_chunk<8, 11, MS>(L, M, beg, end, s_last, a); break;
            default: mtf_thread_chunk<8, 12, MS>(L, M, beg, end, s_last, a); break;
        }
    }
    if (!zero_runs) return;             // coded by the k_zrun_* kernels
    __threadfence_block();
    __syncthreads();
    mtf_zero_runs<MS>(M, mtfv, n, a, s_freq, s_scan, blocks, lb, freq_all);
}

__global__ void __launch_bounds__(MT) k_mtf(const uint8_t *lcol, uint8_t *mtf0, uint16_t *mtfv_all, int32_t *freq_all,
                                            BlockInfo *blocks, int zero_runs)
{
    __shared__ int s_last[32 * 256];
    __shared__ int s_freq[258];
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_carry[2];
    const uint32_t lb = blockIdx.x;
    const int n = (int)blocks[lb].nblock;
    const int a = (int)blocks[lb].n_in_use;
    if (a <= MTF_REG_MAX) return;                               // handled by k_mtf_small
    const uint8_t *L = lcol + (uint64_t)lb * BLK_STRIDE;
    uint8_t *M = mtf0 + (uint64_t)lb * BLK_STRIDE;
    uint16_t *mtfv = mtfv_all + (uint64_t)lb * BLK_STRIDE;
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 256; i += MT) s_last[i] = UNSET;
    for (int i = threadIdx.x; i < 258; i += MT) s_freq[i] = 0;
    __syncthreads();
    int chunk = ((n + 31) / 32 + 31) & ~31;
    int beg = (int)w * chunk, end = beg + chunk;
    if (beg > n) beg = n;
    if (end > n) end = n;
    // phase A: last occurrence of every symbol inside my chunk (scan backwards until all found)
    {
        int *mine = s_last + w * 256;
        int found = 0;
        for (int top = end; top > beg && found < a; top -= 32) {
            int pos = top - 32 + (int)l;
            int sym = pos >= beg ? L[pos] : 256 + (int)l;        // distinct dummies
            unsigned peers = __match_any_sync(0xffffffffu, sym);
            bool isnew = false;
            if (pos >= beg && (int)l == 31 - __clz(peers) && mine[sym] == UNSET) { mine[sym] = pos; isnew = true; }
            found += __popc(__ballot_sync(0xffffffffu, isnew));
            __syncwarp();
        }
    }
    __syncthreads();
    // phase B: exclusive "latest" over chunks, virtual positions for symbols never seen
    if (threadIdx.x < 256) {
        int c = threadIdx.x;
        int run = -(c + 1);
        for (int ww = 0; ww < 32; ww++) {
            int t = s_last[ww * 256 + c];
            s_last[ww * 256 + c] = run;
            if (t != UNSET) run = t;
        }
    }
    __syncthreads();
    // phase C: recency ranks
    if (beg < end) {
        const int *il = s_last + w * 256;
        switch ((a + 31) >> 5) {
            case 1: mtf_chunk<1>(L, M, beg, end, il, a); break;
            case 2: mtf_chunk<2>(L, M, beg, end, il, a); break;
            case 3: mtf_chunk<3>(L, M, beg, end, il, a); break;
            case 4: mtf_chunk<4>(L, M, beg, end, il, a); break;
            case 5: mtf_chunk<5>(L, M, beg, end, il, a); break;
            case 6: mtf_chunk<6>(L, M, beg, end, il, a); break;
            case 7: mtf_chunk<7>(L, M, beg, end, il, a); break;
            default: mtf_chunk<8>(L, M, beg, end, il, a); break;
        }
    }
    if (!zero_runs) return;
    __threadfence_block();
    __syncthreads();
    mtf_zero_runs<MT>(M, mtfv, n, a, s_freq, s_scan, blocks, lb, freq_all);
}

// the register-list kernel in its two forms, under the names the per-kernel report uses
static const auto k_mtf_list_small = k_mtf_small<MS_SMALL, true>;
static const auto k_mtf_list_big = k_mtf_small<MS_BIG, false>;

// CTAs per bzip2 block for a batch of nb blocks when `slots` CTAs of the kernel fit the GPU at once: 1, 2, 4 or 8
// (S3G_CLUSTER=n forces a size, tests run both forms)
static unsigned cluster_size(uint64_t nb, uint64_t slots)
{
    unsigned cs = 1;
    while (cs < 8 && (uint64_t)(cs * 2) * nb <= slots) cs *= 2;
    if (const char *e = getenv("S3G_CLUSTER")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) cs = (unsigned)v; }
    return cs;
}

// A launch with one CTA per block runs in waves of `W` CTAs (W = resident CTAs of the kernel on the GPU); a batch of
// 300 blocks over W = 296 would hold the GPU for two waves, the second one nearly empty, and a batch of 178 for a whole wave
// at 60 % of the slots.  So a batch is cut into chunks: whole waves of one CTA per block, then the remainder as clusters of
// 2, 4 and 8 CTAs per block -- a chunk of m blocks with c CTAs each (c m <= W) takes about 1 / c of a wave.
// plan: (first block of the chunk relative to the batch, blocks, CTAs per block); single = the batch as ONE launch with the
// largest cluster size that fits (the form the tests force with S3G_CLUSTER / S3G_ZRUN).
struct Chunk { uint64_t s0, n; unsigned cs; };
static void plan_chunks(uint64_t nb, uint64_t W, bool single, std::vector<Chunk> &out)
{
    out.clear();
    if (single) { out.push_back({0, nb, nb > W / 2 ? 1u : cluster_size(nb, W)}); return; }
    uint64_t s = 0;
    if (nb >= W) { const uint64_t m = nb / W * W; out.push_back({0, m, 1}); s = m; }
    uint64_t rem = nb - s;
    // what one launch would take against the chunks below (a cluster costs about a fifth more per block)
    if (rem == 0) return;
    const unsigned c1 = rem > W / 2 ? 1u : cluster_size(rem, W);
    const double one = (c1 > 1 ? 1.2 : 1.0) / c1;
    std::vector<Chunk> cut;
    double t = 0;
    uint64_t r = rem, at = s;
    for (unsigned c = 2; c <= 8 && r; c *= 2) {
        const uint64_t cap = W / c;
        if (r >= cap) { cut.push_back({at, cap, c}); at += cap; r -= cap; t += 1.2 / c; }
    }
    if (r) { cut.push_back({at, r, 8}); t += 1.2 / 8; }
    if (t < 0.85 * one) out.insert(out.end(), cut.begin(), cut.end());
    else out.push_back({s, rem, c1});
}

int run_mtf(Ctx *ctx, uint64_t b0, uint64_t nb)
{
    if (nb == 0) return S3G_OK;
    size_t slots = (size_t)nb * BLK_STRIDE;
    S3G_TRY(ctx->mtf0.ensure(slots));                 // MTF ranks before zero-run coding
    S3G_TRY(ctx->mtfv16.ensure(slots * 2));
    S3G_TRY(ctx->mtf_freq.ensure((size_t)nb * 258 * 4));
    S3G_TRY(ctx->ztiles.ensure((size_t)nb * ZNT * sizeof(ZTileInfo)));
    if (!ctx->attr_mtf) {
        const size_t tail_smem0 = 260 * 4 + 40 * 4 + 256 * 4;
        S3G_CUDA(cudaFuncSetAttribute(k_mtf_list_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)24 * MS_SMALL * 4 + tail_smem0)));
        S3G_CUDA(cudaFuncSetAttribute(k_mtf_list_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)MTF_REG_MAX * MS_BIG * 4 + tail_smem0)));
        ctx->attr_mtf = true;
    }
    // zero-run coding inside the MTF kernels (one CTA per block) for the chunks of whole waves, as tile-parallel kernels of
    // its own for the chunks that spread a block over a cluster (S3G_ZRUN=fused|split, S3G_CLUSTER=n: one launch in the
    // form asked for -- the tests run every form)
    const char *ez = getenv("S3G_ZRUN");
    const bool forced = ez || getenv("S3G_CLUSTER");
    std::vector<Chunk> chunks;
    plan_chunks(nb, 2 * SM_COUNT, forced, chunks);
    for (const Chunk &ck : chunks) {
        const uint64_t s0 = ck.s0, n = ck.n;
        // alphabets of <= MTF_REG_MAX symbols take the register-list kernel, larger ones the warp-cooperative one
        double N = 0;
        int rows = 1;
        bool any_small = false, any_big = false, any_huge = false; // which forms of the MTF kernel this chunk needs
        for (uint64_t b = s0; b < s0 + n; b++) {
            int a = b0 + b < ctx->h_blocks.size() ? (int)ctx->h_blocks[b0 + b].n_in_use : 0;
            if (b0 + b < ctx->h_blocks.size()) N += ctx->h_blocks[b0 + b].nblock;
            if (a <= MTF_REG_MAX && a > rows) rows = a;
            if (a == 0) { rows = MTF_REG_MAX; any_small = any_big = any_huge = true; }     // alphabet not mirrored on the host: size for the worst case
            else if (a <= 24) any_small = true;
            else if (a <= MTF_REG_MAX) any_big = true;
            else any_huge = true;
        }
        const size_t tail_smem = 260 * 4 + 40 * 4 + 256 * 4;
        int fused = ck.cs == 1 && n > (uint64_t)SM_COUNT ? 1 : 0;
        if (ez) fused = !strcmp(ez, "split") ? 0 : !strcmp(ez, "fused") ? 1 : fused;
        const unsigned cs = fused ? 1u : (forced ? cluster_size(n, 2 * SM_COUNT) : ck.cs);
        const uint8_t *lcol = ctx->lcol.as<uint8_t>() + s0 * BLK_STRIDE;
        uint8_t *mtf0 = ctx->mtf0.as<uint8_t>() + s0 * BLK_STRIDE;
        uint16_t *mtfv = ctx->mtfv16.as<uint16_t>() + s0 * BLK_STRIDE;
        int32_t *freq = ctx->mtf_freq.as<int32_t>() + s0 * 258;
        BlockInfo *blocks = ctx->blocks.as<BlockInfo>() + b0 + s0;
        S3G_BYTES(ctx, 3 * N + 2 * 0.67 * N);            // L in, ranks out and in, uint16 symbols out (~0.67 per byte)
        if (any_small)
            S3G_LAUNCH_CLUSTER(ctx, k_mtf_list_small, dim3(cs, (unsigned)n), MS_SMALL, (size_t)std::min(rows, 24) * MS_SMALL * 4 + tail_smem, cs, lcol,
                       mtf0, mtfv, freq, blocks, std::min(rows, 24), fused);
        if (any_big)
            S3G_LAUNCH_CLUSTER(ctx, k_mtf_list_big, dim3(cs, (unsigned)n), MS_BIG, (size_t)rows * MS_BIG * 4 + tail_smem, cs, lcol,
                       mtf0, mtfv, freq, blocks, rows, fused);
        if (any_huge) S3G_LAUNCH(ctx, k_mtf, (unsigned)n, MT, 0, lcol, mtf0, mtfv, freq, blocks, fused);
        if (!fused) {
            // zero-run coding, tile-parallel over the ranks
            ZTileInfo *zt = ctx->ztiles.as<ZTileInfo>() + s0 * ZNT;
            S3G_CUDA(cudaMemsetAsync(freq, 0, (size_t)n * 258 * 4, ctx->stream));
            S3G_LAUNCH(ctx, k_zrun_tiles, dim3(ZNT, (unsigned)n), ZT, 0, mtf0, blocks, zt);
            S3G_LAUNCH(ctx, k_zrun_offsets, (unsigned)n, 256, 0, zt, mtfv, freq, blocks);
            S3G_LAUNCH(ctx, k_zrun_emit, dim3(ZNT, (unsigned)n), ZT, 0, mtf0, blocks, zt, mtfv, freq);
        }
    }
    return check_launch("mtf");
}

// =============================================================================
// (3d) Huffman table selection and emission
// =============================================================================
// k_huff runs with HT = 1024 threads (one CTA per SM: shortest time for one block) or, when a batch holds more
// blocks than SMs, with 512 (two CTAs per SM: one block's serial phases overlap the other's parallel ones)
constexpr int G_SIZE = 50;               // BZ_G_SIZE
constexpr int N_ITERS = 4;               // BZ_N_ITERS
constexpr int MAX_SEL = 18002;           // BZ_MAX_SELECTORS
constexpr int ALPHA_MAX = 258;

// bz/huffman.c:63-148, run verbatim by one thread per table
struct HeapScratch {
    int32_t heap[ALPHA_MAX + 2], weight[ALPHA_MAX * 2], parent[ALPHA_MAX * 2];
};
__device__ void hb_make_lengths(uint8_t *len, const int32_t *freq, int alpha, int max_len, HeapScratch &hs)
{
    int32_t *heap = hs.heap, *weight = hs.weight, *parent = hs.parent;
    for (int i = 0; i < alpha; i++) weight[i + 1] = (freq[i] == 0 ? 1 : freq[i]) << 8;
    for (;;) {
        int nnodes = alpha, nheap = 0;
        heap[0] = 0; weight[0] = 0; parent[0] = -2;
        for (int i = 1; i <= alpha; i++) {
            parent[i] = -1;
            nheap++;
            int z = nheap, t = i;
            while (weight[t] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
            heap[z] = t;
        }
        while (nheap > 1) {
            int n12[2];
#pragma unroll 1
            for (int q = 0; q < 2; q++) {
                n12[q] = heap[1]; heap[1] = heap[nheap]; nheap--;
                int z = 1, t = heap[1];
                for (;;) {
                    int y = z << 1;
                    if (y > nheap) break;
                    if (y < nheap && weight[heap[y + 1]] < weight[heap[y]]) y++;
                    if (weight[t] < weight[heap[y]]) break;
                    heap[z] = heap[y]; z = y;
                }
                heap[z] = t;
            }
            nnodes++;
            parent[n12[0]] = parent[n12[1]] = nnodes;
            uint32_t w1 = (uint32_t)weight[n12[0]], w2 = (uint32_t)weight[n12[1]];
            uint32_t d1 = w1 & 0xffu, d2 = w2 & 0xffu;
            weight[nnodes] = (int32_t)(((w1 & 0xffffff00u) + (w2 & 0xffffff00u)) | (1u + (d1 > d2 ? d1 : d2)));
            parent[nnodes] = -1;
            nheap++;
            int z = nheap, t = nnodes;
            while (weight[t] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
            heap[z] = t;
        }
        bool too_long = false;
        for (int i = 1; i <= alpha; i++) {
            int j = 0, k = i;
            while (parent[k] >= 0) { k = parent[k]; j++; }
            len[i - 1] = (uint8_t)j;
            if (j > max_len) too_long = true;
        }
        if (!too_long) break;
        for (int i = 1; i <= alpha; i++) { int j = weight[i] >> 8; j = 1 + (j / 2); weight[i] = j << 8; }
    }
}

// MSB-first bit writer over 32-bit words.  Words wholly inside the writer's bit
// range are stored; the first and last (possibly shared) words are OR-ed in.
struct BitW {
    uint32_t *words;
    uint64_t widx;       // next word to flush
    uint64_t acc;        // pending bits, right-aligned
    int nacc;            // number of pending bits (< 32 after put)
    bool first;
    __device__ void begin(uint32_t *w, uint64_t bitpos)
    {
        words = w; widx = bitpos >> 5; nacc = (int)(bitpos & 31); acc = 0; first = true;
    }
    __device__ __forceinline__ void put(int n, uint32_t v)
    {
        acc = (acc << n) | v; nacc += n;
        if (nacc >= 32) {
            uint32_t out = (uint32_t)(acc >> (nacc - 32));
            if (first) { atomicOr(&words[widx], out); first = false; } else words[widx] = out;
            widx++; nacc -= 32;
            acc &= (1ull << nacc) - 1;
        }
    }
    __device__ void end()
    {
        if (nacc > 0) atomicOr(&words[widx], (uint32_t)(acc << (32 - nacc)));
    }
};

struct HuffSmem {
    uint64_t len_pack[ALPHA_MAX + 2];
    HeapScratch heaps[6];
    uint8_t  len[6][ALPHA_MAX + 2];
    int32_t  rfreq[6][ALPHA_MAX];
    int32_t  code[6][ALPHA_MAX];
    uint8_t  selector[MAX_SEL + 2];
    uint8_t  sel_mtf[MAX_SEL + 2];
    uint32_t scan[33];
    uint32_t tab_bits[8];
    uint32_t hdr_bits;
    // cluster form: what the CTAs of a block tell each other
    int32_t  rtot[6][ALPHA_MAX];       // symbol frequencies summed over the cluster
    uint32_t pub_last[6];              // last group (+1) of this CTA's range that chose table t
    uint32_t pub_bits[2];              // selector bits / symbol bits of this CTA's range
};

// first_block_flags: bit0 set -> this block also carries nothing extra; stream headers are added by the assembler
template <int HT>
__global__ void __launch_bounds__(HT, 1024 / HT) k_huff(const uint16_t *mtfv_all, const int32_t *freq_all, const uint8_t *in_use_all,
                                             BlockInfo *blocks, uint32_t *bits_all, uint8_t *sel_out, uint8_t *len_out,
                                             int with_block_header)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HuffSmem &S = *reinterpret_cast<HuffSmem *>(smem_raw);
    // grid = (cs, blocks): the cs CTAs of a cluster share one bzip2 block; CTA `rank` owns the groups [gr0, gr1)
    const uint32_t lb = blockIdx.y;
    const int rank = (int)blockIdx.x, cs = (int)gridDim.x;
    BlockInfo &B = blocks[lb];
    const int nmtf = (int)B.n_mtf;
    const int alpha = (int)B.n_in_use + 2;
    const uint16_t *mtfv = mtfv_all + (uint64_t)lb * BLK_STRIDE;
    const int32_t *mfreq = freq_all + (uint64_t)lb * 258;
    const uint8_t *in_use = in_use_all + (uint64_t)lb * 256;
    uint32_t *words = bits_all + (uint64_t)lb * BITS_WORDS;
    const int ng = nmtf < 200 ? 2 : nmtf < 600 ? 3 : nmtf < 1200 ? 4 : nmtf < 2400 ? 5 : 6;   // bz/compress.c:273-277
    const int nsel = (nmtf + G_SIZE - 1) / G_SIZE;
    const int tid = threadIdx.x;
    const int gr0 = (int)((long long)nsel * rank / cs), gr1 = (int)((long long)nsel * (rank + 1) / cs);

    for (int i = tid; i < 6 * (ALPHA_MAX + 2); i += HT) (&S.len[0][0])[i] = 15;   // BZ_GREATER_ICOST
    __syncthreads();
    if (tid == 0) {     // initial partition, bz/compress.c:280-317
        int npart = ng, remf = nmtf, gs = 0;
        while (npart > 0) {
            int tfreq = remf / npart, ge = gs - 1, afreq = 0;
            while (afreq < tfreq && ge < alpha - 1) { ge++; afreq += mfreq[ge]; }
            if (ge > gs && npart != ng && npart != 1 && ((ng - npart) % 2 == 1)) { afreq -= mfreq[ge]; ge--; }
            for (int v = 0; v < alpha; v++) S.len[npart - 1][v] = (v >= gs && v <= ge) ? 0 : 15;
            npart--; gs = ge + 1; remf -= afreq;
        }
    }
    __syncthreads();
    for (int iter = 0; iter < N_ITERS; iter++) {
        for (int i = tid; i < 6 * ALPHA_MAX; i += HT) (&S.rfreq[0][0])[i] = 0;
        // six 10-bit cost lanes per symbol (50 symbols x length <= 20 stays below 1024): the same
        // arithmetic as the reference's len_pack fast path (bz/compress.c:334-393), one add per symbol
        for (int v = tid; v < alpha; v += HT) {
            uint64_t pk = 0;
#pragma unroll
            for (int t = 0; t < 6; t++) pk |= (uint64_t)S.len[t][v] << (10 * t);
            S.len_pack[v] = pk;
        }
        __syncthreads();
        for (int g = gr0 + tid; g < gr1; g += HT) {
            const int gs = g * G_SIZE, cnt = min(G_SIZE, nmtf - gs);
            uint32_t w[G_SIZE / 2];
            const uint32_t *src = reinterpret_cast<const uint32_t *>(mtfv + gs);     // gs * 2 bytes is 4-byte aligned
            uint64_t acc = 0;
            if (cnt == G_SIZE) {
#pragma unroll
                for (int k = 0; k < G_SIZE / 2; k++) {
                    w[k] = src[k];
                    acc += S.len_pack[w[k] & 0xffffu] + S.len_pack[w[k] >> 16];
                }
            } else {
                for (int i = 0; i < cnt; i++) acc += S.len_pack[mtfv[gs + i]];
            }
            int bt = 0; uint32_t bc = (uint32_t)(acc & 1023u);
#pragma unroll
            for (int t = 1; t < 6; t++) {
                uint32_t c = (uint32_t)(acc >> (10 * t)) & 1023u;
                if (t < ng && c < bc) { bc = c; bt = t; }          // first minimum, :399-401
            }
            S.selector[g] = (uint8_t)bt;
            // symbol frequencies of the chosen table (:410-432); the eight smallest symbol values
            // (RUNA, RUNB and the nearest MTF ranks -- most of the mass) are counted in a register first
            int32_t *rf = S.rfreq[bt];
            uint64_t small = 0;
            if (cnt == G_SIZE) {
#pragma unroll
                for (int k = 0; k < G_SIZE / 2; k++) {
                    uint32_t v0 = w[k] & 0xffffu, v1 = w[k] >> 16;
                    if (v0 < 8) small += 1ull << (8 * v0); else atomicAdd(&rf[v0], 1);
                    if (v1 < 8) small += 1ull << (8 * v1); else atomicAdd(&rf[v1], 1);
                }
            } else {
                for (int i = 0; i < cnt; i++) {
                    uint32_t v0 = mtfv[gs + i];
                    if (v0 < 8) small += 1ull << (8 * v0); else atomicAdd(&rf[v0], 1);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int c = (int)((small >> (8 * k)) & 255u);
                if (c) atomicAdd(&rf[k], c);
            }
        }
        __syncthreads();
        if (cs > 1) {
            // every CTA sums the frequencies of all CTAs and builds the same lengths from them
            cg::cluster_group cl = cg::this_cluster();
            cl.sync();
            for (int i = tid; i < 6 * ALPHA_MAX; i += HT) {
                int sum = 0;
                for (int r = 0; r < cs; r++) sum += (&cl.map_shared_rank(&S, r)->rfreq[0][0])[i];
                (&S.rtot[0][0])[i] = sum;
            }
            cl.sync();                          // all have read: the next pass may clear rfreq
            if (tid < ng) hb_make_lengths(S.len[tid], S.rtot[tid], alpha, 17, S.heaps[tid]);
        } else if (tid < ng) hb_make_lengths(S.len[tid], S.rfreq[tid], alpha, 17, S.heaps[tid]);
        __syncthreads();
    }
    // codes (bz/huffman.c:152-166) and per-table header sizes
    if (tid < ng) {
        int mn = 32, mx = 0;
        for (int i = 0; i < alpha; i++) { int L = S.len[tid][i]; mx = max(mx, L); mn = min(mn, L); }
        int vec = 0;
        for (int nlen = mn; nlen <= mx; nlen++) {
            for (int i = 0; i < alpha; i++) if (S.len[tid][i] == nlen) S.code[tid][i] = vec++;
            vec <<= 1;
        }
        uint32_t bits = 5; int cur = S.len[tid][0];
        for (int i = 0; i < alpha; i++) { int L = S.len[tid][i]; bits += 2u * (uint32_t)abs(L - cur) + 1; cur = L; }
        S.tab_bits[tid] = bits;
    }
    if (tid == 64) {   // size of the fixed part
        uint32_t hb = with_block_header ? 105u : 0u;
        hb += 16;
        for (int i = 0; i < 16; i++) { bool u = false; for (int j = 0; j < 16; j++) u |= in_use[i * 16 + j] != 0; if (u) hb += 16; }
        hb += 3 + 15;
        S.hdr_bits = hb;
    }
    __syncthreads();
    // ---- selector MTF (bz/compress.c:462-478) in parallel: the MTF position of table v at selector i
    // is the number of tables used more recently than v; a table never used yet sits at its initial
    // index (virtual last use -(t+1)).  Each thread owns a run of selectors and needs, per table,
    // the last use before its run: six exclusive block-wide max-scans.
    {
        const int spt0 = (gr1 - gr0 + HT - 1) / HT;
        const int a0 = min(gr0 + tid * spt0, gr1), a1 = min(a0 + spt0, gr1);
        int last[6];
#pragma unroll
        for (int t = 0; t < 6; t++) last[t] = 0;                 // (position + 1), 0 = not used in my run
        for (int i = a0; i < a1; i++) {
            int v = S.selector[i];
#pragma unroll
            for (int t = 0; t < 6; t++) if (v == t) last[t] = i + 1;
        }
        int before[6];
        uint32_t exv[6];
#pragma unroll
        for (int t = 0; t < 6; t++) {
            uint32_t tot;
            exv[t] = block_excl_max<uint32_t>((uint32_t)last[t], S.scan, &tot);
            if (tid == 0) S.pub_last[t] = tot;
        }
        if (cs > 1) {
            // the groups of the CTAs before this one come first
            cg::cluster_group cl = cg::this_cluster();
            cl.sync();
#pragma unroll
            for (int t = 0; t < 6; t++)
                for (int r = 0; r < rank; r++) { uint32_t v = cl.map_shared_rank(&S, r)->pub_last[t]; if (v > exv[t]) exv[t] = v; }
        }
#pragma unroll
        for (int t = 0; t < 6; t++) before[t] = exv[t] ? (int)exv[t] : -t;   // -t orders never-used tables 0,1,2,... (ties impossible)
        for (int i = a0; i < a1; i++) {
            int v = S.selector[i];
            int lv = 0;
#pragma unroll
            for (int t = 0; t < 6; t++) if (v == t) lv = before[t];
            int j = 0;
#pragma unroll
            for (int t = 0; t < 6; t++) if (t < ng && before[t] > lv) j++;
            S.sel_mtf[i] = (uint8_t)j;
#pragma unroll
            for (int t = 0; t < 6; t++) if (v == t) before[t] = i + 1;
        }
    }
    __syncthreads();
    // ---- layout: [fixed part][selectors][tables][symbols] ----
    const int spt = (gr1 - gr0 + HT - 1) / HT;     // selectors (= groups) per thread
    const int g0 = min(gr0 + tid * spt, gr1), g1 = min(g0 + spt, gr1);
    uint32_t my_sel_bits = 0, my_sym_bits = 0;
    for (int g = g0; g < g1; g++) {
        my_sel_bits += S.sel_mtf[g] + 1u;
        int gs = g * G_SIZE, ge = min(gs + G_SIZE, nmtf);
        const uint8_t *ln = S.len[S.selector[g]];
        if (ge - gs == G_SIZE) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(mtfv + gs);     // gs * 2 bytes is 4-byte aligned
            uint32_t w[G_SIZE / 2];
#pragma unroll
            for (int k = 0; k < G_SIZE / 2; k++) w[k] = src[k];
#pragma unroll
            for (int k = 0; k < G_SIZE / 2; k++) my_sym_bits += (uint32_t)ln[w[k] & 0xffffu] + ln[w[k] >> 16];
        } else
            for (int i = gs; i < ge; i++) my_sym_bits += ln[mtfv[i]];
    }
    uint32_t sel_total, sym_total;
    uint32_t sel_ex = block_excl_sum<uint32_t>(my_sel_bits, S.scan, &sel_total);
    uint32_t sym_ex = block_excl_sum<uint32_t>(my_sym_bits, S.scan, &sym_total);
    if (cs > 1) {
        cg::cluster_group cl = cg::this_cluster();
        if (tid == 0) { S.pub_bits[0] = sel_total; S.pub_bits[1] = sym_total; }
        cl.sync();
        uint32_t sa = 0, sb = 0, ta = 0, tb = 0;
        for (int r = 0; r < cs; r++) {
            const HuffSmem *R = cl.map_shared_rank(&S, r);
            uint32_t x = R->pub_bits[0], y = R->pub_bits[1];
            if (r < rank) { sa += x; sb += y; }
            ta += x; tb += y;
        }
        sel_ex += sa; sym_ex += sb; sel_total = ta; sym_total = tb;
    }
    uint32_t tab_total = 0, tab_off[6];
    for (int t = 0; t < ng; t++) { tab_off[t] = tab_total; tab_total += S.tab_bits[t]; }
    const uint64_t sel_base = S.hdr_bits, tab_base = sel_base + sel_total, sym_base = tab_base + tab_total;
    const uint64_t total_bits = sym_base + sym_total;
    const uint32_t total_words = (uint32_t)((total_bits + 31) >> 5) + 1;
    for (uint32_t i = rank * HT + tid; i < total_words; i += HT * cs) words[i] = 0;
    __syncthreads();
    if (cs > 1) { __threadfence(); cg::this_cluster().sync(); }      // every word is cleared before anyone ORs into it; last remote read is behind us
    if (rank == 0 && tid == 0) {
        BitW bw; bw.begin(words, 0);
        if (with_block_header) {     // bz/compress.c:632-650
            bw.put(24, 0x314159u); bw.put(24, 0x265359u);
            bw.put(16, B.crc >> 16); bw.put(16, B.crc & 0xffffu);
            bw.put(1, 0);
            bw.put(24, (uint32_t)B.orig_ptr & 0xffffffu);
        }
        uint32_t used16 = 0;
        for (int i = 0; i < 16; i++) { bool u = false; for (int j = 0; j < 16; j++) u |= in_use[i * 16 + j] != 0; if (u) used16 |= 1u << (15 - i); }
        bw.put(16, used16);
        for (int i = 0; i < 16; i++) if (used16 & (1u << (15 - i))) {
            uint32_t m = 0;
            for (int j = 0; j < 16; j++) if (in_use[i * 16 + j]) m |= 1u << (15 - j);
            bw.put(16, m);
        }
        bw.put(3, (uint32_t)ng); bw.put(15, (uint32_t)nsel);
        bw.end();
    }
    if (g0 < g1) {   // selectors in unary (:519-524)
        BitW bw; bw.begin(words, sel_base + sel_ex);
        for (int g = g0; g < g1; g++) { int j = S.sel_mtf[g]; bw.put(j + 1, (1u << (j + 1)) - 2u); }
        bw.end();
    }
    if (rank == 0 && tid >= 128 && tid < 128 + ng) {   // delta-coded tables (:531-539)
        int t = tid - 128;
        BitW bw; bw.begin(words, tab_base + tab_off[t]);
        int cur = S.len[t][0];
        bw.put(5, (uint32_t)cur);
        for (int i = 0; i < alpha; i++) {
            int L = S.len[t][i];
            while (cur < L) { bw.put(2, 2); cur++; }
            while (cur > L) { bw.put(2, 3); cur--; }
            bw.put(1, 0);
        }
        bw.end();
    }
    if (g0 < g1) {   // the symbols (:545-594)
        BitW bw; bw.begin(words, sym_base + sym_ex);
        for (int g = g0; g < g1; g++) {
            int gs = g * G_SIZE, ge = min(gs + G_SIZE, nmtf);
            const uint8_t *ln = S.len[S.selector[g]];
            const int32_t *cd = S.code[S.selector[g]];
            if (ge - gs == G_SIZE) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(mtfv + gs);
                uint32_t w[G_SIZE / 2];
#pragma unroll
                for (int k = 0; k < G_SIZE / 2; k++) w[k] = src[k];
#pragma unroll
                for (int k = 0; k < G_SIZE / 2; k++) {
                    uint32_t v0 = w[k] & 0xffffu, v1 = w[k] >> 16;
                    bw.put(ln[v0], (uint32_t)cd[v0]);
                    bw.put(ln[v1], (uint32_t)cd[v1]);
                }
            } else
                for (int i = gs; i < ge; i++) { int v = mtfv[i]; bw.put(ln[v], (uint32_t)cd[v]); }
        }
        bw.end();
    }
    if (rank == 0 && tid == 0) B.n_bits = total_bits;
    if (sel_out) {
        for (int i = gr0 + tid; i < gr1; i += HT) sel_out[(uint64_t)lb * (MAX_SEL + 2) + i] = S.selector[i];
        if (rank == 0)
            for (int i = tid; i < 6 * ALPHA_MAX; i += HT) len_out[(uint64_t)lb * 6 * ALPHA_MAX + i] = S.len[i / ALPHA_MAX][i % ALPHA_MAX];
    }
}

// the Huffman stage's chunks (see run_huff); cs 0: the 512-thread form, one CTA per block
static void plan_huff_chunks(uint64_t nb, bool forced, std::vector<Chunk> &chunks)
{
    chunks.clear();
    if (nb <= (uint64_t)SM_COUNT) chunks.push_back({0, nb, cluster_size(nb, SM_COUNT)});
    else if (forced || nb > (uint64_t)SM_COUNT * 3 / 2) chunks.push_back({0, nb, 0u});
    else plan_chunks(nb, SM_COUNT, false, chunks);
}

int run_huff(Ctx *ctx, uint64_t b0, uint64_t nb, int with_block_header, uint8_t *d_sel_out, uint8_t *d_len_out)
{
    if (nb == 0) return S3G_OK;
    S3G_TRY(ctx->bits.ensure((size_t)nb * BITS_WORDS * 4));
    if (!ctx->attr_huff) {
        S3G_CUDA(cudaFuncSetAttribute(k_huff<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HuffSmem)));
        S3G_CUDA(cudaFuncSetAttribute(k_huff<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HuffSmem)));
        ctx->attr_huff = true;
    }
    // One launch of 512-thread CTAs (two per SM) for a batch of more than 1.5 blocks per SM: the CTAs of this kernel take
    // very different times, so the GPU stays full until the end whatever the count (cutting cfg4's 893 blocks into 888 + 5
    // made it 9.6 ms instead of 9.15).  Up to one block per SM: one 1024-thread CTA per SM, a block spread over a cluster of
    // 1, 2, 4 or 8 of them.  In between (149 .. 222 blocks, 60-75 % of the slots of the two-per-SM form): the first 148 as
    // one CTA per SM, the rest over clusters.
    std::vector<Chunk> chunks;
    plan_huff_chunks(nb, getenv("S3G_CLUSTER") != nullptr, chunks);
    for (const Chunk &ck : chunks) {
        const uint64_t s0 = ck.s0, n = ck.n;
        double N = 0;
        for (uint64_t b = s0; b < s0 + n && b0 + b < ctx->h_blocks.size(); b++) N += ctx->h_blocks[b0 + b].nblock;
        S3G_BYTES(ctx, 6 * 2 * 0.67 * N + 0.25 * N);      // 4 selection passes + size + emit over uint16 symbols, bits out
        const uint16_t *mtfv = ctx->mtfv16.as<uint16_t>() + s0 * BLK_STRIDE;
        const int32_t *freq = ctx->mtf_freq.as<int32_t>() + s0 * 258;
        const uint8_t *in_use = ctx->in_use.as<uint8_t>() + (b0 + s0) * 256;
        BlockInfo *blocks = ctx->blocks.as<BlockInfo>() + b0 + s0;
        uint32_t *bits = ctx->bits.as<uint32_t>() + s0 * BITS_WORDS;
        uint8_t *sel = d_sel_out ? d_sel_out + s0 * (MAX_SEL + 2) : nullptr, *len = d_len_out ? d_len_out + s0 * 6 * ALPHA_MAX : nullptr;
        if (ck.cs == 0)
            S3G_LAUNCH(ctx, k_huff<512>, dim3(1, (unsigned)n), 512, sizeof(HuffSmem), mtfv, freq, in_use, blocks, bits, sel, len, with_block_header);
        else        // one CTA of 1024 threads per SM, every block over a cluster of cs CTAs
            S3G_LAUNCH_CLUSTER(ctx, k_huff<1024>, dim3(ck.cs, (unsigned)n), 1024, sizeof(HuffSmem), ck.cs, mtfv, freq, in_use, blocks, bits, sel, len,
                               with_block_header);
    }
    return check_launch("huff");
}

}  // namespace s3g

extern "C" int s3g_batch_chunks(uint64_t n_blocks, int stage, uint64_t *first, uint64_t *count, uint32_t *ctas, uint64_t cap, uint64_t *n_chunks)
{
    using namespace s3g;
    if (!n_chunks || (stage != 3 && stage != 4)) { set_error("bad argument (stage 3 = MTF, 4 = Huffman)"); return S3G_E_PARAM; }
    std::vector<Chunk> ch;
    if (n_blocks) { if (stage == 3) plan_chunks(n_blocks, 2 * SM_COUNT, false, ch); else plan_huff_chunks(n_blocks, false, ch); }
    *n_chunks = ch.size();
    if (ch.size() > cap) { set_error("cap too small: need %llu", (unsigned long long)ch.size()); return S3G_E_CAPACITY; }
    for (size_t i = 0; i < ch.size(); i++) {
        if (first) first[i] = ch[i].s0;
        if (count) count[i] = ch[i].n;
        if (ctas) ctas[i] = ch[i].cs;
    }
    return S3G_OK;
}
