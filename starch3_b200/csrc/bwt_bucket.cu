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
constexpr int BS_CAP = 2560;                     // rotations a bucket may hold: 2.9 x the average (16 samples per bucket: once in 10^7 buckets by chance)
constexpr int BS_T = 256;                        // threads of the tile kernels and of the bucket kernel
constexpr int BS_NSUB = 32;                      // sub-buckets of a bucket
constexpr int BS_TAB = BS_BMAX + 1;
constexpr int BS_LUT = 4096;                     // the key's top 12 bits (keys are below 2^44)
constexpr int BS_ST = 512;                       // threads of the scatter kernel: tiles of 8192 records, 8 per bucket and tile

__host__ __device__ inline uint32_t bs_buckets(uint32_t n)
{
    uint32_t want = (n + BS_AVG - 1) / BS_AVG, b = 1;
    while (b < want && b < (uint32_t)BS_BMAX) b <<= 1;
    return b;
}

struct BsP {
    BwtP P;
    uint64_t *spl;             // [nb][BS_BMAX] splitters: keys, ascending, padded with ~0
    uint16_t *lut;             // [nb][BS_LUT + 2] splitters below every value of the key's top 12 bits
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
    __shared__ uint64_t sp_s[BS_BMAX];
    for (uint32_t i = tid; i < (uint32_t)BS_BMAX; i += 1024) {
        uint64_t v = i + 1 < nbk ? smp[BS_A * (i + 1) - 1] : ~0ull;
        sp_s[i] = v;
        B.spl[(uint64_t)lb * BS_BMAX + i] = v;
        B.gcount[(uint64_t)lb * BS_BMAX + i] = 0;
    }
    __syncthreads();
    // lut[t] = splitters below t << 32: the search for a key with top bits t only looks at [lut[t], lut[t + 1])
    for (uint32_t t = tid; t <= (uint32_t)BS_LUT; t += 1024) {
        const uint64_t x = (uint64_t)t << 32;
        uint32_t pos = 0;
        for (uint32_t step = nbk >> 1; step; step >>= 1) if (sp_s[pos + step - 1] < x) pos += step;
        B.lut[(uint64_t)lb * (BS_LUT + 2) + t] = (uint16_t)pos;
    }
}

// ---- the tile kernels: keys of TH * 16 consecutive positions -----------------------------------------------
template <int TH> struct BsTileSmem {
    __align__(16) uint8_t sym[TH * SI + 64];
    uint8_t seq[256], frac[256];
    uint32_t cnt[BS_BMAX];
    uint32_t scan[33];
};
struct BsCountSmem {
    uint64_t sp[BS_BMAX];
    uint16_t lut[BS_LUT + 2];
};
template <int TH> struct BsScatSmem {
    uint64_t stage[TH * SI];
    uint16_t sbid[TH * SI];
    uint32_t tbase[BS_BMAX], goff[BS_BMAX];
};

// symbols of the tile (and the k + 1 that follow it, cyclic) as ranks, into S.sym
template <int TH>
__device__ __forceinline__ void bs_load_tile(BsTileSmem<TH> &S, const uint8_t *b, uint32_t tbase0, uint32_t cntT, uint32_t n, uint32_t k, uint32_t tid)
{
    constexpr uint32_t TILE = TH * SI;
    for (uint32_t i0 = tid * 16; i0 < TILE; i0 += TH * 16) {
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
    for (uint32_t i = TILE + tid; i < cntT + k + 1; i += TH) {
        uint32_t q = tbase0 + i;
        if (q >= n) { q -= n; if (q >= n) q %= n; }
        S.sym[i] = S.seq[b[q]];
    }
}

// SCATTER = false: bucket of every position (bid) and the bucket sizes; true: the records, in bucket order
template <bool SCATTER, int TH>
__global__ void __launch_bounds__(TH, SCATTER ? 2 : 4) k_bs_tile(BsP B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t TILE = TH * SI;
    constexpr size_t S_BYTES = (sizeof(BsTileSmem<TH>) + 15) & ~(size_t)15;
    BsTileSmem<TH> &S = *reinterpret_cast<BsTileSmem<TH> *>(smem_raw);
    BsCountSmem &C = *reinterpret_cast<BsCountSmem *>(smem_raw + S_BYTES);
    BsScatSmem<TH> &X = *reinterpret_cast<BsScatSmem<TH> *>(smem_raw + S_BYTES);
    const uint32_t lb = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (B.P.mode[lb] != 1) return;
    const uint32_t n = B.P.cnt_n[lb];
    const uint32_t tbase0 = tile * TILE;
    if (tbase0 >= n) return;
    const uint32_t cntT = min(TILE, n - tbase0), nbk = bs_buckets(n);
    const uint8_t *b = B.P.blk + B.P.blocks[lb].blk_off;
    const uint32_t k = B.P.init_k[lb], a = B.P.init_a[lb], f = B.P.init_f[lb];
    if (tid < 256) {
        S.seq[tid] = B.P.seq[(uint64_t)lb * 256 + tid];
        S.frac[tid] = (uint8_t)(tid < a ? tid * f / a : 0);
    }
    for (uint32_t i = tid; i < (uint32_t)BS_BMAX; i += TH) { S.cnt[i] = 0; if (!SCATTER) C.sp[i] = B.spl[(uint64_t)lb * BS_BMAX + i]; }
    if (!SCATTER) for (uint32_t i = tid; i < (uint32_t)BS_LUT + 2; i += TH) C.lut[i] = B.lut[(uint64_t)lb * (BS_LUT + 2) + i];
    __syncthreads();
    bs_load_tile<TH>(S, b, tbase0, cntT, n, k, tid);
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
                const uint64_t x = key * f + S.frac[S.sym[p + k]];
                // number of splitters below x: they are ascending, and those that share x's top 12 bits sit in [lut[t], lut[t + 1])
                const uint32_t t = (uint32_t)(x >> 32);
                uint32_t lo = C.lut[t], hi = C.lut[t + 1];
                while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (C.sp[mid] < x) lo = mid + 1; else hi = mid; }
                ids[r] = lo;
                atomicAdd(&S.cnt[lo], 1u);
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
        for (uint32_t i = tid; i < nbk; i += TH) { uint32_t c = S.cnt[i]; if (c) atomicAdd(&B.gcount[(uint64_t)lb * BS_BMAX + i], c); }
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
    // the records go to the staging buffer in two steps (rank first, place after the bucket offsets are known); the key is
    // rebuilt for the second step rather than kept in 32 registers
    uint16_t rnk[SI];
#pragma unroll
    for (int r = 0; r < SI; r++) {
        rnk[r] = 0;
        if (p0 + r < cntT) rnk[r] = (uint16_t)atomicAdd(&S.cnt[ids[r]], 1u);      // any order inside the bucket will do: the bucket is sorted as a whole later
    }
    __syncthreads();
    // per bucket: offset inside the tile, and room in the bucket (one atomic per bucket that the tile touches)
    constexpr int BPT = BS_BMAX / TH;
    static_assert(BS_BMAX % TH == 0, "every bucket needs an owner thread");
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
    for (int r = 0; r < SI; r++) {
        uint32_t p = p0 + r;
        if (p < cntT) {
            uint32_t at = X.tbase[ids[r]] + rnk[r];
            X.stage[at] = ((key * f + S.frac[S.sym[p + k]]) << VAL_BITS) | (tbase0 + p);
            X.sbid[at] = (uint16_t)ids[r];
            key = (key - S.sym[p] * pw) * a + S.sym[p + k];
        }
    }
    __syncthreads();
    uint64_t *out = B.kv + (uint64_t)lb * BLK_STRIDE;
    for (uint32_t i = tid; i < tile_total; i += TH) {
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
// One CTA per bucket, a thread per record, everything in shared memory:
//   splitters   64 records of the bucket ranked by counting; every second one splits -> 32 sub-buckets of ~27 records
//   partition   count pass, scan, scatter pass (the sub-bucket search is repeated rather than kept in registers)
//   rank        every record counts the smaller records of its sub-bucket: its final slot by (key, start)
//   ties        groups of equal keys: deeper symbols from the block bytes, counting inside the group, level by level
//   emit        ptr[], last column, origPtr; what is still tied after FLEVELS levels goes out flagged for the doubling rounds
struct BsSortSmem {
    uint64_t A[BS_CAP];                 // the bucket as it arrives; sorted by (key, start) after the rank pass
    uint64_t Bk[BS_CAP];                // the bucket in sub-bucket order; the deeper keys afterwards
    uint16_t grp[2][BS_CAP];            // tie resolution: first final slot of the (sub-)group a sorted position is in; 0xffff = done
    uint8_t sbB[BS_CAP];                // sub-bucket of every Bk slot
    uint64_t sp2[BS_NSUB];              // sub-bucket splitters (keys, ascending, padded with ~0)
    uint32_t cnt[BS_NSUB], start[BS_NSUB + 1], cur[BS_NSUB];
    uint32_t bad;
    uint8_t seq[256];
};

struct BsBlk {
    const uint8_t *b; uint32_t n, k0, a, m;      // block bytes, size, symbols in the key, alphabet, symbols per HALF of a deeper key
};

// the next 2 m symbols (m <= 8) of rotation `pos` from depth d on: two mixed-radix numbers below 2^32 side by side.
// The bytes come in three aligned 8-byte loads issued together (dependent byte loads cost an L2 round trip each).
__device__ __forceinline__ uint64_t bs_deep_key(const BsBlk &K, const uint8_t *sq, uint32_t pos, uint32_t d)
{
    uint64_t q = (uint64_t)pos + d;
    if (q >= K.n) q %= K.n;
    uint32_t qq = (uint32_t)q;
    uint32_t hi = 0, lo = 0;
    if (qq + 16 <= K.n) {
        // block bytes start 128-byte aligned and their slot is padded (blk_slot_bytes): the 24 bytes from qq & ~7 are readable
        const uint64_t *wp = reinterpret_cast<const uint64_t *>(K.b + (qq & ~7u));
        const uint64_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const uint32_t sh = (qq & 7u) * 8;
        const uint64_t b0 = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;       // bytes qq .. qq+7
        const uint64_t b1 = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;       // bytes qq+8 .. qq+15
#pragma unroll
        for (uint32_t t = 0; t < 16; t++) {
            const uint32_t sy = sq[(uint32_t)((t < 8 ? b0 : b1) >> (8 * (t & 7))) & 255u];
            if (t < K.m) hi = hi * K.a + sy;
            else if (t < 2 * K.m) lo = lo * K.a + sy;
        }
    } else {
        for (uint32_t t = 0; t < 2 * K.m; t++) {
            const uint32_t sy = sq[K.b[qq]];
            if (t < K.m) hi = hi * K.a + sy; else lo = lo * K.a + sy;
            if (++qq == K.n) qq = 0;
        }
    }
    return (uint64_t)hi << 32 | lo;
}

// sub-bucket of a key: the splitters below it; a key that fills more than one sample interval (consecutive equal
// splitters) gets the second of those intervals for itself, so that a big group of equal keys does not share a sub-bucket
__device__ __forceinline__ uint32_t bs_sub(const uint64_t *sp2, uint64_t key)
{
    uint32_t pos = 0;
#pragma unroll
    for (uint32_t step = BS_NSUB >> 1; step; step >>= 1) if (sp2[pos + step - 1] < key) pos += step;
    if (pos + 1 < (uint32_t)BS_NSUB && sp2[pos] == key && sp2[pos + 1] == key) pos++;
    return pos;
}

__global__ void __launch_bounds__(BS_T, 4) k_bs_sort(BsP B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BsSortSmem &S = *reinterpret_cast<BsSortSmem *>(smem_raw);
    const uint32_t lb = blockIdx.y, bkt = blockIdx.x, tid = threadIdx.x;
    if (B.P.mode[lb] != 1) return;
    const uint32_t n = B.P.cnt_n[lb], nbk = bs_buckets(n);
    if (bkt >= nbk) return;
    const uint32_t base = B.bstart[(uint64_t)lb * BS_TAB + bkt];
    const uint32_t nrec = (bkt + 1 < nbk ? B.bstart[(uint64_t)lb * BS_TAB + bkt + 1] : n) - base;
    if (nrec == 0) return;
    BsBlk K;
    K.b = B.P.blk + B.P.blocks[lb].blk_off; K.n = n; K.k0 = B.P.init_k[lb]; K.a = B.P.init_a[lb];
    { uint32_t m = 1; uint64_t pwm = K.a; while (m < 8 && pwm * K.a <= (1ull << 32)) { pwm *= K.a; m++; } K.m = m; }
    S.seq[tid] = B.P.seq[(uint64_t)lb * 256 + tid];
    if (tid < BS_NSUB) { S.cnt[tid] = 0; S.sp2[tid] = ~0ull; }
    if (tid == 0) S.bad = 0;
    const uint64_t *in = B.kv + (uint64_t)lb * BLK_STRIDE + base;
    for (uint32_t i = tid; i < nrec; i += BS_T) S.A[i] = in[i];
    __syncthreads();
    // ---- splitters ----
    if (nrec > 64) {
        if (tid < 64) S.Bk[64 + tid] = S.A[(uint32_t)((uint64_t)tid * nrec / 64)] >> VAL_BITS;
        __syncthreads();
        if (tid < 64) {
            const uint64_t my = S.Bk[64 + tid];
            uint32_t r = 0;
            for (uint32_t t = 0; t < 64; t++) { uint64_t x = S.Bk[64 + t]; r += (x < my) | ((x == my) & (t < tid)); }
            S.Bk[r] = my;                                  // the sample, ascending
        }
        __syncthreads();
        if (tid < BS_NSUB - 1) S.sp2[tid] = S.Bk[2 * tid + 1];
        __syncthreads();
    }
    // ---- partition into sub-buckets ----
    for (uint32_t i = tid; i < nrec; i += BS_T) atomicAdd(&S.cnt[bs_sub(S.sp2, S.A[i] >> VAL_BITS)], 1u);
    __syncthreads();
    if (tid < 32) {
        uint32_t c = S.cnt[tid];
        uint32_t inc = warp_incl_sum<uint32_t>(c);
        S.start[tid] = inc - c; S.cur[tid] = inc - c;
        if (tid == 31) S.start[32] = inc;
        if (c > 512u) S.bad = 1;                           // very many equal keys: counting inside the sub-bucket would take too long
    }
    __syncthreads();
    if (S.bad) { if (tid == 0) { B.P.mode[lb] = 2; B.flags[0] = 1; } return; }
    for (uint32_t i = tid; i < nrec; i += BS_T) {
        const uint64_t rec = S.A[i];
        const uint32_t sb = bs_sub(S.sp2, rec >> VAL_BITS);
        const uint32_t slot = atomicAdd(&S.cur[sb], 1u);
        S.Bk[slot] = rec; S.sbB[slot] = (uint8_t)sb;
    }
    __syncthreads();
    // ---- rank inside the sub-bucket: records are distinct (the start is part of them), so the counts are the slots ----
    for (uint32_t j = tid; j < nrec; j += BS_T) {
        const uint64_t rec = S.Bk[j];
        const uint32_t sb = S.sbB[j], s0 = S.start[sb], s1 = S.start[sb + 1];
        uint32_t r = 0;
        for (uint32_t t = s0; t < s1; t++) r += S.Bk[t] < rec;
        S.A[s0 + r] = rec;
    }
    __syncthreads();
    // ---- singletons out; groups of equal keys marked ----
    uint32_t *sa = B.P.sa + (uint64_t)lb * BLK_STRIDE + base;
    uint8_t *L = B.lcol + (uint64_t)lb * BLK_STRIDE + base;
    auto emit = [&](uint32_t slot, uint64_t rec, uint32_t flag) {
        const uint32_t pos = (uint32_t)rec & VMASK;
        sa[slot] = pos | flag;
        L[slot] = S.seq[K.b[pos ? pos - 1 : K.n - 1]];                 // bz/compress.c:166-167
        if (pos == 0) B.blocks[lb].orig_ptr = (int32_t)(base + slot);
    };
    bool anytie = false;
    for (uint32_t p = tid; p < nrec; p += BS_T) {
        const uint64_t rec = S.A[p], key = rec >> VAL_BITS;
        const bool head = p == 0 || (S.A[p - 1] >> VAL_BITS) != key;
        const bool tail = p + 1 == nrec || (S.A[p + 1] >> VAL_BITS) != key;
        if (head && tail) { emit(p, rec, 0u); S.grp[0][p] = 0xffffu; }
        else {
            uint32_t gs = p;
            while (gs > 0 && (S.A[gs - 1] >> VAL_BITS) == key && p - gs <= 256u) gs--;
            if (p - gs > 256u) S.bad = 1;                  // a group the radix form's finisher is built for
            S.grp[0][p] = (uint16_t)gs;
            anytie = true;
        }
    }
    anytie = __syncthreads_or(anytie);
    if (S.bad) { if (tid == 0) { B.P.mode[lb] = 2; B.flags[0] = 1; } return; }
    // ---- deeper levels ----
    uint32_t cb = 0;
    for (uint32_t level = 0; level < (uint32_t)FLEVELS && anytie; level++) {
        const uint32_t d = K.k0 + level * 2 * K.m;
        for (uint32_t p = tid; p < nrec; p += BS_T)
            if (S.grp[cb][p] != 0xffffu) S.Bk[p] = bs_deep_key(K, S.seq, (uint32_t)S.A[p] & VMASK, d);
        __syncthreads();
        bool still = false;
        for (uint32_t p = tid; p < nrec; p += BS_T) {
            const uint32_t g = S.grp[cb][p];
            uint32_t ng = 0xffffu;
            if (g != 0xffffu) {
                const uint64_t rec = S.A[p], key = rec >> VAL_BITS, my = S.Bk[p];
                uint32_t lo = p, hi = p + 1;                // the group of equal keys around p
                while (lo > 0 && (S.A[lo - 1] >> VAL_BITS) == key) lo--;
                while (hi < nrec && (S.A[hi] >> VAL_BITS) == key) hi++;
                uint32_t lt = 0, eq = 0;
                for (uint32_t t = lo; t < hi; t++) {
                    if (S.grp[cb][t] != g) continue;        // another sub-group of the same key
                    const uint64_t x = S.Bk[t];
                    lt += x < my; eq += x == my;
                }
                if (eq == 1) emit(g + lt, rec, 0u);
                else { ng = g + lt; still = true; }
            }
            S.grp[cb ^ 1][p] = (uint16_t)ng;
        }
        cb ^= 1;
        anytie = __syncthreads_or(still);
    }
    // ---- what is still tied: out in arrival order, flagged as one unsorted group ----
    uint32_t left = 0;
    if (anytie) {
        for (uint32_t p = tid; p < nrec; p += BS_T) {
            const uint32_t g = S.grp[cb][p];
            if (g == 0xffffu) continue;
            const uint64_t rec = S.A[p], key = rec >> VAL_BITS;
            uint32_t lo = p, eqb = 0;
            while (lo > 0 && (S.A[lo - 1] >> VAL_BITS) == key) lo--;
            for (uint32_t t = lo; t < p; t++) eqb += S.grp[cb][t] == g;
            emit(g + eqb, rec, eqb ? NONHEAD : 0u);
            left++;
        }
    }
    for (int dd = 16; dd; dd >>= 1) left += __shfl_xor_sync(0xffffffffu, left, dd);
    if ((tid & 31) == 0 && left) { atomicAdd(&B.P.left[lb], left); atomicAdd(B.g_left, (unsigned long long)left); }
}

// blocks handed over to the radix form: forget what the bucket form had counted for them
__global__ void k_bs_reset(BsP B, uint32_t nb, uint32_t which)
{
    uint32_t lb = blockIdx.x * blockDim.x + threadIdx.x;
    if (lb >= nb || B.P.mode[lb] != which) return;
    uint32_t l = B.P.left[lb];
    if (l) { atomicAdd(B.g_left, (unsigned long long)0 - (unsigned long long)l); B.P.left[lb] = 0; }
}

static const auto k_bs_count = k_bs_tile<false, BS_T>;
static const auto k_bs_scatter = k_bs_tile<true, BS_ST>;

int run_bucket_sort(Ctx *ctx, const BwtP &P, uint64_t b0, uint64_t nb, unsigned long long *g_left, uint32_t *d_flags)
{
    BsP B;
    B.P = P;
    // tables in the look-back status buffer of the radix form (unused here; its tags are reset below), bucket ids in rk
    uint32_t *tab = ctx->hist.as<uint32_t>();
    B.spl = reinterpret_cast<uint64_t *>(tab); tab += 2 * nb * BS_BMAX;
    B.gcount = tab; tab += nb * BS_BMAX;
    B.bstart = tab; tab += nb * BS_TAB;
    B.cursor = tab; tab += nb * BS_BMAX;
    B.lut = reinterpret_cast<uint16_t *>(tab);
    B.bid = reinterpret_cast<uint16_t *>(P.rk);
    B.kv = P.kv0;
    B.flags = d_flags;
    B.lcol = ctx->lcol.as<uint8_t>();
    B.blocks = ctx->blocks.as<BlockInfo>() + b0;
    B.g_left = g_left;
    ctx->sweep_cap = 0;                     // the status words of the radix passes were overwritten
    const size_t count_smem = ((sizeof(BsTileSmem<BS_T>) + 15) & ~(size_t)15) + sizeof(BsCountSmem);
    const size_t tile_smem = ((sizeof(BsTileSmem<BS_ST>) + 15) & ~(size_t)15) + sizeof(BsScatSmem<BS_ST>);
    const unsigned scat_tiles = (unsigned)((BLK_STRIDE + BS_ST * SI - 1) / (BS_ST * SI));
    if (!ctx->attr_bs) {
        S3G_CUDA(cudaFuncSetAttribute(k_bs_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_SMAX * 8));
        S3G_CUDA(cudaFuncSetAttribute(k_bs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
        S3G_CUDA(cudaFuncSetAttribute(k_bs_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)count_smem));
        S3G_CUDA(cudaFuncSetAttribute(k_bs_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BsSortSmem)));
        ctx->attr_bs = true;
    }
    double N = 0;
    for (uint64_t b = 0; b < nb && b0 + b < ctx->h_blocks.size(); b++) N += ctx->h_blocks[b0 + b].nblock;
    S3G_LAUNCH(ctx, k_bs_sample, (unsigned)nb, 1024, BS_SMAX * 8, B);
    S3G_BYTES(ctx, 3 * N);
    S3G_LAUNCH(ctx, k_bs_count, dim3(NT, (unsigned)nb), BS_T, count_smem, B);
    S3G_LAUNCH(ctx, k_bs_scan, (unsigned)nb, BS_BMAX, 0, B);
    S3G_BYTES(ctx, 11 * N);
    S3G_LAUNCH(ctx, k_bs_scatter, dim3(scat_tiles, (unsigned)nb), BS_ST, tile_smem, B);
    S3G_BYTES(ctx, 13 * N);
    S3G_LAUNCH(ctx, k_bs_sort, dim3(BS_BMAX, (unsigned)nb), BS_T, sizeof(BsSortSmem), B);
    S3G_LAUNCH(ctx, k_bs_reset, (unsigned)((nb + 127) / 128), 128, 0, B, (uint32_t)nb, 2u);
    return check_launch("bucket sort");
}

}  // namespace s3g
