// bwt_bucket.cu -- kernel (3b), bucket form: BZ2_blockSort (bz/blocksort.c:1031-1089) as bucket-then-finish, the shape
// of bzip2's own mainSort (bz/blocksort.c:751-1011: bucket on a prefix, finish every bucket by comparing strings),
// with buckets that fit shared memory so that a rotation crosses HBM twice instead of once per radix pass.
//
// Same contract as bwt.cu: ptr[] = rotation starts in ascending order of the cyclic rotations (bz/blocksort.c:347-469),
// origPtr = where rotation 0 lands (:1083-1086), the last column (bz/compress.c:166-167) on the side.
//
//   k_bs_sample   per block: 16 sample keys per bucket (the 44-bit key of bwt.cu), sorted in shared memory; every 16th
//                 is a splitter, so the buckets hold about the same number of rotations whatever the text looks like
//                 (fixed key bits cannot do that: half the rotations of a BED block share their first three symbols)
//   k_bs_count    per tile of 4096 positions: rolling key, bucket by binary search over the splitters, bucket sizes
//   k_bs_scan     bucket starts; a block with a bucket beyond the shared-memory capacity is left to the radix form
//   k_bs_scatter  per tile: the 64-bit records (key << 20 | start) staged in bucket order, written out in pieces
//   k_bs_sort     per bucket, in shared memory: 32 sub-buckets from a sorted sample of the bucket itself, each sorted by a
//                 warp in registers (bitonic network over shuffles), groups of equal keys ranked on deeper symbols read
//                 from the block bytes (all-pairs counting inside the group, up to FLEVELS levels), then ptr[], the last
//                 column and origPtr written once.  What is still tied goes out flagged for the prefix-doubling rounds
//                 of bwt.cu, exactly as its group finisher leaves it.
// HBM bytes per rotation: 1 + 2 (count) + 3 + 8 (scatter) + 8 + 5 (sort) = 27, against 111 for five radix passes + finisher.
#include "bwt.cuh"

namespace s3g {

constexpr int BS_A = 16;                         // samples per bucket
constexpr int BS_BMAX = 1024;                    // buckets of a full block
constexpr int BS_SMAX = BS_A * BS_BMAX;          // samples of a full block
constexpr int BS_AVG = 880;                      // rotations per bucket aimed at (a full block: 899 981 / 1024)
constexpr int BS_CAP = 3584;                     // rotations a bucket may hold: 4 x the average (16 samples per bucket: never reached by chance)
constexpr int BS_T = 256;                        // threads of the tile kernels and of the bucket kernel
constexpr int BS_RPT = BS_CAP / BS_T;            // records per thread in the bucket kernel
constexpr int BS_NSUB = 32;                      // sub-buckets of a bucket
constexpr int BS_SUBCAP = 256;                   // rotations a sub-bucket may hold (8 per lane)
constexpr int BS_TAB = BS_BMAX + 1;

__host__ __device__ inline uint32_t bs_buckets(uint32_t n)
{
    uint32_t want = (n + BS_AVG - 1) / BS_AVG, b = 1;
    while (b < want && b < (uint32_t)BS_BMAX) b <<= 1;
    return b;
}

struct BsP {
    BwtP P;
    uint32_t *sp32;            // [nb][BS_BMAX] splitters: the top 32 bits of the key (ascending, padded with ~0)
    uint32_t *gcount;          // [nb][BS_BMAX] bucket sizes
    uint32_t *bstart;          // [nb][BS_TAB]  bucket starts
    uint32_t *cursor;          // [nb][BS_BMAX] next free slot of a bucket during the scatter
    uint16_t *bid;             // [nb][BLK_STRIDE] bucket of every position
    uint64_t *kv;              // [nb][BLK_STRIDE] records in bucket order
    uint32_t *flags;           // [0]: a block was handed to the radix form
    uint8_t *lcol;
    BlockInfo *blocks;         // writable view of P.blocks
    unsigned long long *g_left;
};

// the 44-bit key of rotation p: mixed-radix value of its first k symbols, times f, plus the class of the next symbol
// (the same value k_sweep_first builds with a rolling update)
__device__ __forceinline__ uint64_t bs_key_at(const uint8_t *b, const uint8_t *sq, uint32_t p, uint32_t n, uint32_t k, uint32_t a, uint32_t f)
{
    uint64_t key = 0;
    uint32_t q = p;
    for (uint32_t t = 0; t < k; t++) { key = key * a + sq[b[q]]; if (++q == n) q = 0; }
    return key * f + (uint32_t)sq[b[q]] * f / a;
}

// ---- samples and splitters -----------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_bs_sample(BsP B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *smp = reinterpret_cast<uint64_t *>(smem_raw);
    __shared__ uint8_t sq[256];
    const uint32_t lb = blockIdx.x, tid = threadIdx.x;
    if (B.P.mode[lb] != 1) return;
    const uint32_t n = B.P.cnt_n[lb], nbk = bs_buckets(n), S = nbk * BS_A;
    const uint8_t *b = B.P.blk + B.P.blocks[lb].blk_off;
    const uint32_t k = B.P.init_k[lb], a = B.P.init_a[lb], f = B.P.init_f[lb];
    if (tid < 256) sq[tid] = B.P.seq[(uint64_t)lb * 256 + tid];
    __syncthreads();
    for (uint32_t j = tid; j < S; j += 1024) smp[j] = bs_key_at(b, sq, (uint32_t)((uint64_t)j * n / S), n, k, a, f);
    __syncthreads();
    for (uint32_t kk = 2; kk <= S; kk <<= 1) {
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t t = tid; t < S / 2; t += 1024) {
                uint32_t i = 2 * t - (t & (j - 1)), ix = i + j;          // the pair (i, i + j), i without bit j
                uint64_t x = smp[i], y = smp[ix];
                bool up = (i & kk) == 0;
                if ((x > y) == up) { smp[i] = y; smp[ix] = x; }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < (uint32_t)BS_BMAX; i += 1024) {
        B.sp32[(uint64_t)lb * BS_BMAX + i] = i + 1 < nbk ? (uint32_t)(smp[BS_A * (i + 1) - 1] >> 12) : 0xffffffffu;
        B.gcount[(uint64_t)lb * BS_BMAX + i] = 0;
    }
}

// ---- the tile kernels: keys of 4096 consecutive positions ------------------------------------------------
struct BsTileSmem {
    __align__(16) uint8_t sym[STILE + 64];
    uint8_t seq[256], frac[256];
    uint32_t sp[BS_BMAX];
    uint32_t cnt[BS_BMAX];
    uint32_t scan[33];
};
struct BsScatSmem {
    uint64_t stage[STILE];
    uint16_t sbid[STILE];
    uint32_t tbase[BS_BMAX], goff[BS_BMAX];
};

// symbols of the tile (and the k + 1 that follow it, cyclic) as ranks, into S.sym
__device__ __forceinline__ void bs_load_tile(BsTileSmem &S, const uint8_t *b, uint32_t tbase0, uint32_t cntT, uint32_t n, uint32_t k, uint32_t tid)
{
    for (uint32_t i0 = tid * 16; i0 < (uint32_t)STILE; i0 += BS_T * 16) {
        if (tbase0 + i0 + 16 <= n) {
            uint4 v = *reinterpret_cast<const uint4 *>(b + tbase0 + i0);
            uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; j++)
                wv[j] = (uint32_t)S.seq[wv[j] & 255] | (uint32_t)S.seq[(wv[j] >> 8) & 255] << 8 | (uint32_t)S.seq[(wv[j] >> 16) & 255] << 16 |
                        (uint32_t)S.seq[wv[j] >> 24] << 24;
            *reinterpret_cast<uint4 *>(S.sym + i0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        } else {
            for (uint32_t i = i0; i < i0 + 16 && i < cntT + k + 1; i++) {
                uint32_t q = tbase0 + i;
                if (q >= n) { q -= n; if (q >= n) q %= n; }
                S.sym[i] = S.seq[b[q]];
            }
        }
    }
    for (uint32_t i = STILE + tid; i < cntT + k + 1; i += BS_T) {
        uint32_t q = tbase0 + i;
        if (q >= n) { q -= n; if (q >= n) q %= n; }
        S.sym[i] = S.seq[b[q]];
    }
}

// SCATTER = false: bucket of every position (bid) and the bucket sizes; true: the records, in bucket order
template <bool SCATTER>
__global__ void __launch_bounds__(BS_T) k_bs_tile(BsP B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BsTileSmem &S = *reinterpret_cast<BsTileSmem *>(smem_raw);
    BsScatSmem &X = *reinterpret_cast<BsScatSmem *>(smem_raw + ((sizeof(BsTileSmem) + 15) & ~(size_t)15));
    const uint32_t lb = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (B.P.mode[lb] != 1) return;
    const uint32_t n = B.P.cnt_n[lb];
    const uint32_t tbase0 = tile * STILE;
    if (tbase0 >= n) return;
    const uint32_t cntT = min((uint32_t)STILE, n - tbase0), nbk = bs_buckets(n);
    const uint8_t *b = B.P.blk + B.P.blocks[lb].blk_off;
    const uint32_t k = B.P.init_k[lb], a = B.P.init_a[lb], f = B.P.init_f[lb];
    S.seq[tid] = B.P.seq[(uint64_t)lb * 256 + tid];
    S.frac[tid] = (uint8_t)(tid < a ? tid * f / a : 0);
    for (uint32_t i = tid; i < (uint32_t)BS_BMAX; i += BS_T) { S.cnt[i] = 0; if (!SCATTER) S.sp[i] = B.sp32[(uint64_t)lb * BS_BMAX + i]; }
    __syncthreads();
    bs_load_tile(S, b, tbase0, cntT, n, k, tid);
    __syncthreads();
    uint64_t pw = 1;
    for (uint32_t i = 1; i < k; i++) pw *= a;
    const uint32_t p0 = tid * SI;
    uint64_t key = 0;
    if (p0 < cntT) for (uint32_t j = 0; j < k; j++) key = key * a + S.sym[p0 + j];
    uint16_t *bidp = B.bid + (uint64_t)lb * BLK_STRIDE + tbase0 + p0;
    if (!SCATTER) {
        uint32_t ids[SI];
#pragma unroll
        for (int r = 0; r < SI; r++) {
            uint32_t p = p0 + r;
            ids[r] = 0;
            if (p < cntT) {
                uint32_t x = (uint32_t)((key * f + S.frac[S.sym[p + k]]) >> 12);
                uint32_t pos = 0;                                 // number of splitters below x (nbk - 1 of them, ascending)
                for (uint32_t step = nbk >> 1; step; step >>= 1) if (S.sp[pos + step - 1] < x) pos += step;
                ids[r] = pos;
                atomicAdd(&S.cnt[pos], 1u);
                key = (key - S.sym[p] * pw) * a + S.sym[p + k];
            }
        }
        if (p0 + SI <= cntT) {
            uint4 lo = make_uint4(ids[0] | ids[1] << 16, ids[2] | ids[3] << 16, ids[4] | ids[5] << 16, ids[6] | ids[7] << 16);
            uint4 hi = make_uint4(ids[8] | ids[9] << 16, ids[10] | ids[11] << 16, ids[12] | ids[13] << 16, ids[14] | ids[15] << 16);
            reinterpret_cast<uint4 *>(bidp)[0] = lo; reinterpret_cast<uint4 *>(bidp)[1] = hi;
        } else {
#pragma unroll
            for (int r = 0; r < SI; r++) if (p0 + r < cntT) bidp[r] = (uint16_t)ids[r];
        }
        __syncthreads();
        for (uint32_t i = tid; i < nbk; i += BS_T) { uint32_t c = S.cnt[i]; if (c) atomicAdd(&B.gcount[(uint64_t)lb * BS_BMAX + i], c); }
        return;
    }
    // ---- scatter ----
    uint32_t ids[SI];
    if (p0 + SI <= cntT) {
        uint4 lo = reinterpret_cast<const uint4 *>(bidp)[0], hi = reinterpret_cast<const uint4 *>(bidp)[1];
        uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int r = 0; r < SI; r++) ids[r] = (w[r >> 1] >> (16 * (r & 1))) & 0xffffu;
    } else {
#pragma unroll
        for (int r = 0; r < SI; r++) ids[r] = p0 + r < cntT ? bidp[r] : 0u;
    }
    uint64_t rec[SI];
    uint16_t rnk[SI];
#pragma unroll
    for (int r = 0; r < SI; r++) {
        uint32_t p = p0 + r;
        rec[r] = 0; rnk[r] = 0;
        if (p < cntT) {
            rec[r] = ((key * f + S.frac[S.sym[p + k]]) << VAL_BITS) | (tbase0 + p);
            rnk[r] = (uint16_t)atomicAdd(&S.cnt[ids[r]], 1u);      // any order inside the bucket will do: the bucket is sorted as a whole later
            key = (key - S.sym[p] * pw) * a + S.sym[p + k];
        }
    }
    __syncthreads();
    // per bucket: offset inside the tile, and room in the bucket (one atomic per bucket that the tile touches)
    constexpr int BPT = BS_BMAX / BS_T;
    uint32_t c4[BPT], mysum = 0;
#pragma unroll
    for (int q = 0; q < BPT; q++) { c4[q] = S.cnt[tid * BPT + q]; mysum += c4[q]; }
    uint32_t tile_total;
    uint32_t ex = block_excl_sum<uint32_t>(mysum, S.scan, &tile_total);
#pragma unroll
    for (int q = 0; q < BPT; q++) {
        uint32_t bkt = tid * BPT + q;
        X.tbase[bkt] = ex; ex += c4[q];
        X.goff[bkt] = c4[q] ? atomicAdd(&B.cursor[(uint64_t)lb * BS_BMAX + bkt], c4[q]) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SI; r++)
        if (p0 + r < cntT) { uint32_t at = X.tbase[ids[r]] + rnk[r]; X.stage[at] = rec[r]; X.sbid[at] = (uint16_t)ids[r]; }
    __syncthreads();
    uint64_t *out = B.kv + (uint64_t)lb * BLK_STRIDE;
    for (uint32_t i = tid; i < tile_total; i += BS_T) {
        uint32_t bkt = X.sbid[i];
        out[X.goff[bkt] + (i - X.tbase[bkt])] = X.stage[i];
    }
}

__global__ void __launch_bounds__(BS_BMAX) k_bs_scan(BsP B)
{
    __shared__ uint32_t sm[33];
    const uint32_t lb = blockIdx.x, tid = threadIdx.x;
    if (B.P.mode[lb] != 1) return;
    const uint32_t n = B.P.cnt_n[lb], nbk = bs_buckets(n);
    uint32_t c = tid < nbk ? B.gcount[(uint64_t)lb * BS_BMAX + tid] : 0u, tot;
    uint32_t ex = block_excl_sum<uint32_t>(c, sm, &tot);
    B.bstart[(uint64_t)lb * BS_TAB + tid] = ex;
    B.cursor[(uint64_t)lb * BS_BMAX + tid] = ex;
    if (tid == 0) B.bstart[(uint64_t)lb * BS_TAB + BS_BMAX] = tot;
    bool over = c > (uint32_t)BS_CAP;
    if (__syncthreads_or(over)) { if (tid == 0) { B.P.mode[lb] = 2; B.flags[0] = 1; } }
}

// ---- the bucket kernel --------------------------------------------------------------------------------------
// bitonic network over 32 R values per warp: element e = s * 32 + lane is v[s] of that lane
template <int R> __device__ __forceinline__ void bs_warp_sort(uint64_t (&v)[R], uint32_t lane)
{
#pragma unroll
    for (int kk = 2; kk <= 32 * R; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int js = j >> 5;
#pragma unroll
                for (int s = 0; s < R; s++) {
                    if ((s & js) == 0) {
                        const bool up = ((s * 32) & kk) == 0;         // lane < 32 <= j < kk: bit kk of e is bit kk of s * 32
                        uint64_t x = v[s], y = v[s | js];
                        if ((x > y) == up) { v[s] = y; v[s | js] = x; }
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < R; s++) {
                    const uint32_t e = (uint32_t)s * 32 + lane;
                    uint64_t y = __shfl_xor_sync(0xffffffffu, v[s], j);
                    const bool up = (e & (uint32_t)kk) == 0, lower = (lane & (uint32_t)j) == 0;
                    const bool keep_min = lower == up;
                    v[s] = keep_min ? (v[s] < y ? v[s] : y) : (v[s] > y ? v[s] : y);
                }
            }
        }
    }
}

struct BsSortSmem {
    uint64_t A[BS_CAP];                 // the bucket; in sub-bucket order after the partition
    uint64_t strip[BS_T / 32][BS_SUBCAP];   // per warp: deeper keys of the sub-bucket being finished
    uint64_t sp2[BS_NSUB];              // sub-bucket splitters (keys, ascending, padded with ~0)
    uint32_t cnt[BS_NSUB], start[BS_NSUB + 1];
    uint32_t bad;
    uint8_t seq[256];
};

struct BsBlk {
    const uint8_t *b; uint32_t n, k0, a, m;      // block bytes, size, symbols in the key, alphabet, symbols per deeper level
};

// the next m symbols of rotation `pos` from depth d on, as a mixed-radix number (a^m <= 2^63)
__device__ __forceinline__ uint64_t bs_deep_key(const BsBlk &K, const uint8_t *sq, uint32_t pos, uint32_t d)
{
    uint64_t q = (uint64_t)pos + d;
    if (q >= K.n) q %= K.n;
    uint32_t qq = (uint32_t)q;
    uint64_t key = 0;
    for (uint32_t t = 0; t < K.m; t++) { key = key * K.a + sq[K.b[qq]]; if (++qq == K.n) qq = 0; }
    return key;
}

// One warp: sort the g <= 32 R records at src by (key, start), rank the groups of equal keys on deeper symbols, write
// ptr[], last column, origPtr for output slots obase ...  Returns the rotations left tied.
template <int R>
__device__ __forceinline__ uint32_t bs_finish_sub(const uint64_t *src, uint32_t g, uint32_t obase, uint64_t *strip, const BsBlk &K, const uint8_t *sq,
                                                  uint32_t *sa, uint8_t *L, BlockInfo *blk_info, uint32_t lane)
{
    uint64_t v[R];
#pragma unroll
    for (int s = 0; s < R; s++) { uint32_t e = (uint32_t)s * 32 + lane; v[s] = e < g ? src[e] : ~0ull; }
    bs_warp_sort<R>(v, lane);
    // group heads: the key differs from the element before; elements past g count as heads
    uint32_t hb[R];
#pragma unroll
    for (int s = 0; s < R; s++) {
        uint64_t pv = __shfl_up_sync(0xffffffffu, v[s], 1);
        uint64_t pl = s > 0 ? __shfl_sync(0xffffffffu, v[s > 0 ? s - 1 : 0], 31) : 0ull;
        if (lane == 0) pv = pl;
        uint32_t e = (uint32_t)s * 32 + lane;
        bool head = e == 0 || e >= g || (v[s] >> VAL_BITS) != (pv >> VAL_BITS);
        hb[s] = __ballot_sync(0xffffffffu, head);
    }
    uint32_t gs[R], ge[R], pe[R];
    uint32_t tied = 0;                               // bit s: element s of this lane is still tied
#pragma unroll
    for (int s = 0; s < R; s++) {
        uint32_t e = (uint32_t)s * 32 + lane;
        // last head at or before e
        uint32_t m = hb[s] & (0xffffffffu >> (31 - lane));
        uint32_t a0 = 0;
        if (m) a0 = (uint32_t)s * 32 + 31 - (uint32_t)__clz(m);
        else {
#pragma unroll
            for (int t = R - 1; t >= 0; t--) if (t < s && hb[t] && a0 == 0) a0 = (uint32_t)t * 32 + 31 - (uint32_t)__clz(hb[t]) + 0x10000u;
            a0 &= 0xffffu;
        }
        // first head after e (32 R if none: cannot happen while e < g, positions past g are heads)
        uint32_t m2 = lane < 31 ? hb[s] & (0xfffffffeu << lane) : 0u;
        uint32_t a1 = 32u * R;
        if (m2) a1 = (uint32_t)s * 32 + (uint32_t)__ffs(m2) - 1;
        else {
            bool found = false;
#pragma unroll
            for (int t = 0; t < R; t++) if (t > s && hb[t] && !found) { a1 = (uint32_t)t * 32 + (uint32_t)__ffs(hb[t]) - 1; found = true; }
        }
        if (a1 > g) a1 = g;
        gs[s] = a0; ge[s] = a1; pe[s] = e;
        if (e < g && a1 - a0 > 1) tied |= 1u << s;
    }
    // deeper levels: all-pairs counting inside the group on the next m symbols
    for (uint32_t level = 0; level < (uint32_t)FLEVELS; level++) {
        if (!__any_sync(0xffffffffu, tied != 0)) break;
        const uint32_t d = K.k0 + level * K.m;
        uint64_t dk[R];
#pragma unroll
        for (int s = 0; s < R; s++) {
            dk[s] = 0;
            if ((tied >> s) & 1u) { dk[s] = bs_deep_key(K, sq, (uint32_t)v[s] & VMASK, d); strip[pe[s]] = dk[s]; }
        }
        __syncwarp();
        uint32_t ngs[R], neq[R], npe[R];
#pragma unroll
        for (int s = 0; s < R; s++) {
            ngs[s] = neq[s] = npe[s] = 0;
            if ((tied >> s) & 1u) {
                uint32_t lt = 0, eq = 0, eqb = 0;
                const uint64_t my = dk[s];
                for (uint32_t t = gs[s]; t < ge[s]; t++) {
                    uint64_t x = strip[t];
                    lt += x < my;
                    uint32_t is = x == my;
                    eq += is;
                    eqb += is & (uint32_t)(t < pe[s]);
                }
                ngs[s] = gs[s] + lt; neq[s] = eq; npe[s] = ngs[s] + eqb;
            }
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < R; s++) {
            if ((tied >> s) & 1u) {
                gs[s] = ngs[s]; ge[s] = ngs[s] + neq[s]; pe[s] = npe[s];
                if (neq[s] == 1) tied &= ~(1u << s);
            }
        }
    }
    uint32_t left = 0;
#pragma unroll
    for (int s = 0; s < R; s++) {
        uint32_t e = (uint32_t)s * 32 + lane;
        if (e < g) {
            uint32_t pos = (uint32_t)v[s] & VMASK;
            bool t = (tied >> s) & 1u;
            uint32_t slot = obase + pe[s];
            sa[slot] = pos | (t && pe[s] != gs[s] ? NONHEAD : 0u);
            L[slot] = sq[K.b[pos ? pos - 1 : K.n - 1]];              // bz/compress.c:166-167
            if (pos == 0) blk_info->orig_ptr = (int32_t)slot;
            left += t;
        }
    }
    return left;
}

__global__ void __launch_bounds__(BS_T) k_bs_sort(BsP B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BsSortSmem &S = *reinterpret_cast<BsSortSmem *>(smem_raw);
    const uint32_t lb = blockIdx.y, bkt = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (B.P.mode[lb] != 1) return;
    const uint32_t n = B.P.cnt_n[lb];
    if (bkt >= bs_buckets(n)) return;
    const uint32_t base = B.bstart[(uint64_t)lb * BS_TAB + bkt];
    const uint32_t nrec = (bkt + 1 < bs_buckets(n) ? B.bstart[(uint64_t)lb * BS_TAB + bkt + 1] : n) - base;
    if (nrec == 0) return;
    BsBlk K;
    K.b = B.P.blk + B.P.blocks[lb].blk_off; K.n = n; K.k0 = B.P.init_k[lb]; K.a = B.P.init_a[lb];
    { uint32_t m = 1; uint64_t pwm = K.a; while (pwm <= (1ull << 63) / K.a) { pwm *= K.a; m++; } K.m = m; }
    S.seq[tid] = B.P.seq[(uint64_t)lb * 256 + tid];
    if (tid < BS_NSUB) S.cnt[tid] = 0;
    if (tid == 0) S.bad = 0;
    const uint64_t *in = B.kv + (uint64_t)lb * BLK_STRIDE + base;
    uint64_t rec[BS_RPT];
#pragma unroll
    for (int r = 0; r < BS_RPT; r++) {
        uint32_t i = tid + (uint32_t)r * BS_T;
        rec[r] = i < nrec ? in[i] : ~0ull;
        if (i < nrec) S.A[i] = rec[r];
    }
    __syncthreads();
    uint32_t *sa = B.P.sa + (uint64_t)lb * BLK_STRIDE;
    uint8_t *L = B.lcol + (uint64_t)lb * BLK_STRIDE;
    uint32_t left = 0;
    if (nrec <= (uint32_t)BS_SUBCAP) {
        // a small bucket is one sub-bucket: warp 0 finishes it
        if (w == 0) {
            if (nrec <= 32) left = bs_finish_sub<1>(S.A, nrec, base, S.strip[0], K, S.seq, sa, L, B.blocks + lb, lane);
            else if (nrec <= 64) left = bs_finish_sub<2>(S.A, nrec, base, S.strip[0], K, S.seq, sa, L, B.blocks + lb, lane);
            else if (nrec <= 128) left = bs_finish_sub<4>(S.A, nrec, base, S.strip[0], K, S.seq, sa, L, B.blocks + lb, lane);
            else left = bs_finish_sub<8>(S.A, nrec, base, S.strip[0], K, S.seq, sa, L, B.blocks + lb, lane);
        }
    } else {
        // ---- sub-bucket splitters: 128 keys of the bucket itself, sorted by warp 0; every fourth is a splitter ----
        if (w == 0) {
            uint64_t sv[4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                uint32_t j = (uint32_t)s * 32 + lane;
                sv[s] = S.A[(uint32_t)((uint64_t)j * nrec / 128)] >> VAL_BITS;
            }
            bs_warp_sort<4>(sv, lane);
            // element e = s * 32 + lane; splitter i = element 4 i + 3
#pragma unroll
            for (int s = 0; s < 4; s++) {
                uint32_t e = (uint32_t)s * 32 + lane;
                if ((e & 3u) == 3u) S.sp2[e >> 2] = e == 127 ? ~0ull : sv[s];
            }
        }
        __syncthreads();
        uint32_t where[BS_RPT];
#pragma unroll
        for (int r = 0; r < BS_RPT; r++) {
            where[r] = 0;
            if (tid + (uint32_t)r * BS_T < nrec) {
                const uint64_t key = rec[r] >> VAL_BITS;
                uint32_t pos = 0;                                  // splitters below the key (31 of them)
#pragma unroll
                for (uint32_t step = BS_NSUB >> 1; step; step >>= 1) if (S.sp2[pos + step - 1] < key) pos += step;
                where[r] = pos << 16 | atomicAdd(&S.cnt[pos], 1u);
            }
        }
        __syncthreads();
        if (w == 0) {
            uint32_t c = S.cnt[lane];
            uint32_t inc = warp_incl_sum<uint32_t>(c);
            S.start[lane] = inc - c;
            if (lane == 31) S.start[32] = inc;
            if (__any_sync(0xffffffffu, c > (uint32_t)BS_SUBCAP)) { if (lane == 0) S.bad = 1; }
        }
        __syncthreads();
        if (S.bad) {
            // very many equal keys in one sub-bucket: the whole block goes to the radix form (its group finisher takes groups up to 2048)
            if (tid == 0) { B.P.mode[lb] = 2; B.flags[0] = 1; }
            return;
        }
#pragma unroll
        for (int r = 0; r < BS_RPT; r++)
            if (tid + (uint32_t)r * BS_T < nrec) S.A[S.start[where[r] >> 16] + (where[r] & 0xffffu)] = rec[r];
        __syncthreads();
        for (uint32_t sb = w; sb < (uint32_t)BS_NSUB; sb += BS_T / 32) {
            const uint32_t g = S.cnt[sb], s0 = S.start[sb];
            if (g == 0) continue;
            if (g <= 32) left += bs_finish_sub<1>(S.A + s0, g, base + s0, S.strip[w], K, S.seq, sa, L, B.blocks + lb, lane);
            else if (g <= 64) left += bs_finish_sub<2>(S.A + s0, g, base + s0, S.strip[w], K, S.seq, sa, L, B.blocks + lb, lane);
            else if (g <= 128) left += bs_finish_sub<4>(S.A + s0, g, base + s0, S.strip[w], K, S.seq, sa, L, B.blocks + lb, lane);
            else left += bs_finish_sub<8>(S.A + s0, g, base + s0, S.strip[w], K, S.seq, sa, L, B.blocks + lb, lane);
            __syncwarp();
        }
    }
    for (int d = 16; d; d >>= 1) left += __shfl_xor_sync(0xffffffffu, left, d);
    if (lane == 0 && left) { atomicAdd(&B.P.left[lb], left); atomicAdd(B.g_left, (unsigned long long)left); }
}

// blocks handed over to the radix form: forget what the bucket form had counted for them
__global__ void k_bs_reset(BsP B, uint32_t nb, uint32_t which)
{
    uint32_t lb = blockIdx.x * blockDim.x + threadIdx.x;
    if (lb >= nb || B.P.mode[lb] != which) return;
    uint32_t l = B.P.left[lb];
    if (l) { atomicAdd(B.g_left, (unsigned long long)0 - (unsigned long long)l); B.P.left[lb] = 0; }
}

static const auto k_bs_count = k_bs_tile<false>;
static const auto k_bs_scatter = k_bs_tile<true>;

int run_bucket_sort(Ctx *ctx, const BwtP &P, uint64_t b0, uint64_t nb, unsigned long long *g_left, uint32_t *d_flags)
{
    BsP B;
    B.P = P;
    // tables in the look-back status buffer of the radix form (unused here; its tags are reset below), bucket ids in rk
    uint32_t *tab = ctx->hist.as<uint32_t>();
    B.sp32 = tab; tab += nb * BS_BMAX;
    B.gcount = tab; tab += nb * BS_BMAX;
    B.bstart = tab; tab += nb * BS_TAB;
    B.cursor = tab;
    B.bid = reinterpret_cast<uint16_t *>(P.rk);
    B.kv = P.kv0;
    B.flags = d_flags;
    B.lcol = ctx->lcol.as<uint8_t>();
    B.blocks = ctx->blocks.as<BlockInfo>() + b0;
    B.g_left = g_left;
    ctx->sweep_cap = 0;                     // the status words of the radix passes were overwritten
    const size_t tile_smem = ((sizeof(BsTileSmem) + 15) & ~(size_t)15) + sizeof(BsScatSmem);
    if (!ctx->attr_bs) {
        S3G_CUDA(cudaFuncSetAttribute(k_bs_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_SMAX * 8));
        S3G_CUDA(cudaFuncSetAttribute(k_bs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
        S3G_CUDA(cudaFuncSetAttribute(k_bs_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BsSortSmem)));
        ctx->attr_bs = true;
    }
    double N = 0;
    for (uint64_t b = 0; b < nb && b0 + b < ctx->h_blocks.size(); b++) N += ctx->h_blocks[b0 + b].nblock;
    S3G_LAUNCH(ctx, k_bs_sample, (unsigned)nb, 1024, BS_SMAX * 8, B);
    S3G_BYTES(ctx, 3 * N);
    S3G_LAUNCH(ctx, k_bs_count, dim3(NT, (unsigned)nb), BS_T, sizeof(BsTileSmem), B);
    S3G_LAUNCH(ctx, k_bs_scan, (unsigned)nb, BS_BMAX, 0, B);
    S3G_BYTES(ctx, 11 * N);
    S3G_LAUNCH(ctx, k_bs_scatter, dim3(NT, (unsigned)nb), BS_T, tile_smem, B);
    S3G_BYTES(ctx, 13 * N);
    S3G_LAUNCH(ctx, k_bs_sort, dim3(BS_BMAX, (unsigned)nb), BS_T, sizeof(BsSortSmem), B);
    S3G_LAUNCH(ctx, k_bs_reset, (unsigned)((nb + 127) / 128), 128, 0, B, (uint32_t)nb, 2u);
    return check_launch("bucket sort");
}

}  // namespace s3g
