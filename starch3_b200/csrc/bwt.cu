// bwt.cu -- kernel (3b): the block sort of bzip2 (BZ2_blockSort,
// bz/blocksort.c:1031-1089) for a batch of blocks: radix sort on a prefix key, groups finished in
// shared memory, prefix doubling as the fallback.
//
// What must match the reference: ptr[0..n) = rotation starts in ascending
// order of the n cyclic rotations of the block (bz/blocksort.c:347-469 compares
// with wrap-around) and origPtr = the index of rotation 0 (:1083-1086).  Which
// algorithm produces the order is free (bz/blocksort.c:1058-1061).
//
// Algorithm (all blocks of a batch in the same launches; DESIGN.md section 4, "The block sort"):
//   keys     k_keys: the 32-bit key (k32 symbols) of every position, and the digit histograms of all radix
//            passes over the 44-bit initial key (mixed-radix key of the first k symbols, k = max{k : A^k <= 2^44},
//            A = symbols in use, plus a class of the next symbol in the room that is left).
//   sort     k_sweep x 5: LSD radix passes over 9-bit digits, one kernel each (the first builds the 64-bit
//            records key << 20 | rotation start from the block bytes; tile ranking, decoupled look-back).
//   finish   k_finish_rows / k_finish_mid / k_finish_big (a warp's window / a warp / a CTA per group, by size):
//            every group of equal keys is ranked on the 32-bit keys of the
//            positions k, k+k32, ... further on -- bzip2's own "bucket, then compare strings"
//            (bz/blocksort.c:751-1011) for blocks with short common prefixes.  Writes ptr[], the last column
//            and origPtr, and checks that the keys it receives ascend.
//   fallback what the finisher leaves (groups > 2048, ties deeper than its levels) goes through Manber-Myers
//            prefix-doubling rounds: given the order by the first h symbols (SA, grouped; RK[i] = SA position
//            of the first member of i's group, bit 31 = group is a singleton), walking SA in order,
//            j = SA[k]-h arrives in ascending RK[j+h]; a STABLE counting sort of those j by RK[j] sorts every
//            group by its second half (two 10-bit radix passes over the still-unsorted rotations).  Rounds
//            stop when every group is a singleton or h >= n (equal rotations: a periodic block, flagged in
//            BlockInfo.tie and resolved by k_fallback_exact).
#include "bwt.cuh"

namespace s3g {

// record p of block lb for the given source mode; returns false if the item does not take part
template <int MODE>
__device__ __forceinline__ bool get_item(const BwtP &P, uint32_t lb, uint32_t p, uint32_t n, uint32_t cnt, uint32_t h,
                                         const uint64_t *kv_in, uint64_t &rec)
{
    if (p >= cnt) return false;
    if (MODE == MODE_INIT) {
        const uint8_t *b = P.blk + P.blocks[lb].blk_off;
        const uint8_t *sq = P.seq + (uint64_t)lb * 256;
        uint32_t k = P.init_k[lb], k32 = P.init_k32[lb], a = P.init_a[lb];
        uint64_t key_ = 0;
        uint32_t q = p;
        for (uint32_t t = 0; t < k; t++) {
            key_ = key_ * a + sq[b[q]];
            q++; if (q == n) q = 0;
            // the key of the first k32 symbols stays behind in rk: the group finisher reads deeper symbols from it
            if (t + 1 == k32) P.rk[(uint64_t)lb * BLK_STRIDE + p] = (uint32_t)key_;
        }
        rec = (key_ << VAL_BITS) | p;
        return true;
    } else if (MODE == MODE_MM) {
        uint32_t s = P.sa[(uint64_t)lb * BLK_STRIDE + p];
        uint32_t j = s >= h ? s - h : s + n - h;          // h < n is guaranteed by the caller
        uint32_t r = P.rk[(uint64_t)lb * BLK_STRIDE + j];
        if (r & FINAL) return false;
        rec = ((uint64_t)r << 32) | j;
        return true;
    } else {
        rec = kv_in[(uint64_t)lb * BLK_STRIDE + p];
        return MODE == MODE_KVX ? rec != ~0ull : true;
    }
}

// depth (symbols already sorted) of block lb in doubling round `round`
__device__ __forceinline__ uint32_t depth_of(const BwtP &P, uint32_t lb, uint32_t round)
{
    return round >= 25 ? 0x7fffffffu : P.init_k[lb] << round;      // init_k <= 40
}
// phase 0 = initial sort (every block), phase 1 = doubling round (unsorted blocks whose depth is below n)
__device__ __forceinline__ bool block_live(const BwtP &P, uint32_t lb, int phase, uint32_t round, const uint32_t *act_cur)
{
    if (phase == 0) return true;
    return act_cur[lb] != 0 && depth_of(P, lb, round) < P.cnt_n[lb];
}

// ---- radix pass: per-tile digit histogram ------------------------------------
template <int MODE>
__global__ void __launch_bounds__(ST) k_hist(BwtP P, int rshift, int phase, uint32_t round, const uint32_t *cnt_arr,
                                             const uint64_t *kv_in, const uint32_t *act_cur, uint64_t *kv_save)
{
    __shared__ uint32_t sh[NBINS];
    uint32_t lb = blockIdx.y, tile = blockIdx.x;
    if (!block_live(P, lb, phase, round, act_cur)) return;
    uint32_t h = depth_of(P, lb, round);
    uint32_t n = P.cnt_n[lb], cnt = cnt_arr[lb];
    if ((uint64_t)tile * STILE >= cnt) return;
    for (int i = threadIdx.x; i < NBINS; i += ST) sh[i] = 0;
    __syncthreads();
    uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31;
    uint32_t base = tile * STILE + w * (SI * 32) + l;
    uint32_t dg[SI];
#pragma unroll
    for (int r = 0; r < SI; r++) {
        uint64_t rec = 0;
        bool ok = get_item<MODE>(P, lb, base + r * 32, n, cnt, h, kv_in, rec);
        dg[r] = ok ? ((uint32_t)(rec >> rshift) & (NBINS - 1)) : 0xffffffffu;
        // the gathered records are kept so that the scatter of this pass reads them back coalesced
        if (kv_save && base + r * 32 < cnt) kv_save[(uint64_t)lb * BLK_STRIDE + base + r * 32] = ok ? rec : ~0ull;
    }
#pragma unroll
    for (int r = 0; r < SI; r++) {
        unsigned peers = __match_any_sync(0xffffffffu, dg[r]);
        if (dg[r] != 0xffffffffu && (peers & ((1u << l) - 1)) == 0) atomicAdd(&sh[dg[r]], __popc(peers));
    }
    __syncthreads();
    uint32_t *out = P.hist + ((uint64_t)lb * NT + tile) * NBINS;
    for (int i = threadIdx.x; i < NBINS; i += ST) out[i] = sh[i];
}

// ---- radix pass: exclusive scan of hist[digit][tile] per block ---------------
__global__ void __launch_bounds__(NBINS) k_hist_scan(BwtP P, int phase, uint32_t round, const uint32_t *cnt_arr, uint32_t *total_out,
                                                     const uint32_t *act_cur)
{
    __shared__ uint32_t sm[33];
    uint32_t lb = blockIdx.x;
    if (!block_live(P, lb, phase, round, act_cur)) return;
    uint32_t cnt = cnt_arr[lb];
    uint32_t ntiles = (cnt + STILE - 1) / STILE;
    // thread d owns digit d: column d of the [tile][digit] matrix
    uint32_t *col = P.hist + (uint64_t)lb * NT * NBINS + threadIdx.x;
    uint32_t s = 0;
#pragma unroll 16
    for (uint32_t t = 0; t < ntiles; t++) s += col[(uint64_t)t * NBINS];
    uint32_t tot;
    uint32_t ex = block_excl_sum<uint32_t>(s, sm, &tot);
#pragma unroll 16
    for (uint32_t t = 0; t < ntiles; t++) { uint32_t v = col[(uint64_t)t * NBINS]; col[(uint64_t)t * NBINS] = ex; ex += v; }
    if (threadIdx.x == 0 && total_out) total_out[lb] = tot;
}

// ---- radix pass: stable scatter ----------------------------------------------
// Ranks are computed warp-synchronously (match.any on the digit), items are first
// placed in digit order inside shared memory, then written out so that neighbouring
// threads store to neighbouring addresses (runs of equal digits are contiguous in HBM).
struct ScatterSmem {
    uint64_t stage[STILE];                 // tile in digit order
    uint32_t gbase[NBINS];                 // global offset of (digit, this tile)
    uint32_t tbase[NBINS];                 // offset of the digit inside the tile
    uint16_t wcnt[(ST / 32) * NBINS];      // per-warp digit counters -> exclusive warp offsets
    uint32_t scan[33];
};

template <int MODE>
__global__ void __launch_bounds__(ST, 3) k_scatter(BwtP P, int rshift, int phase, uint32_t round, const uint32_t *cnt_arr,
                                                const uint64_t *kv_in, uint64_t *kv_out, const uint32_t *act_cur)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem &S = *reinterpret_cast<ScatterSmem *>(smem_raw);
    uint32_t lb = blockIdx.y, tile = blockIdx.x;
    if (!block_live(P, lb, phase, round, act_cur)) return;
    uint32_t h = depth_of(P, lb, round);
    uint32_t n = P.cnt_n[lb], cnt = cnt_arr[lb];
    if ((uint64_t)tile * STILE >= cnt) return;
    for (int i = threadIdx.x; i < (ST / 32) * NBINS; i += ST) S.wcnt[i] = 0;
    const uint32_t *hrow = P.hist + ((uint64_t)lb * NT + tile) * NBINS;
    for (int i = threadIdx.x; i < NBINS; i += ST) S.gbase[i] = hrow[i];
    uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31;
    uint32_t base = tile * STILE + w * (SI * 32) + l;
    uint16_t *mycnt = S.wcnt + w * NBINS;
    uint64_t kv[SI];
    uint16_t rnk[SI];
    uint32_t okmask = 0;
    // all loads first (independent, 16 in flight per thread); the ranking below is warp-synchronous
#pragma unroll
    for (int r = 0; r < SI; r++) {
        uint64_t rec = 0;
        bool ok = get_item<MODE>(P, lb, base + r * 32, n, cnt, h, kv_in, rec);
        kv[r] = rec;
        if (ok) okmask |= 1u << r;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SI; r++) {
        bool ok = (okmask >> r) & 1u;
        uint32_t d = ok ? ((uint32_t)(kv[r] >> rshift) & (NBINS - 1)) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned lt = peers & ((1u << l) - 1);
        uint16_t b = ok ? mycnt[d] : (uint16_t)0;
        __syncwarp();
        if (ok && lt == 0) mycnt[d] = (uint16_t)(b + __popc(peers));
        __syncwarp();
        rnk[r] = (uint16_t)(b + __popc(lt));
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, tile total; then exclusive scan of the totals over digits
    uint32_t tot4[NBINS / ST];
    uint32_t mysum = 0;
#pragma unroll
    for (int q = 0; q < NBINS / ST; q++) {
        int d = threadIdx.x * (NBINS / ST) + q;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < ST / 32; ww++) { uint32_t c = S.wcnt[ww * NBINS + d]; S.wcnt[ww * NBINS + d] = (uint16_t)run; run += c; }
        tot4[q] = run; mysum += run;
    }
    uint32_t tile_total;
    uint32_t ex = block_excl_sum<uint32_t>(mysum, S.scan, &tile_total);
#pragma unroll
    for (int q = 0; q < NBINS / ST; q++) { S.tbase[threadIdx.x * (NBINS / ST) + q] = ex; ex += tot4[q]; }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SI; r++) {
        if (okmask & (1u << r)) {
            uint32_t d = (uint32_t)(kv[r] >> rshift) & (NBINS - 1);
            S.stage[S.tbase[d] + mycnt[d] + rnk[r]] = kv[r];
        }
    }
    __syncthreads();
    uint64_t *out = kv_out + (uint64_t)lb * BLK_STRIDE;
    for (uint32_t i = threadIdx.x; i < tile_total; i += ST) {
        uint64_t it = S.stage[i];
        uint32_t d = (uint32_t)(it >> rshift) & (NBINS - 1);
        out[S.gbase[d] + (i - S.tbase[d])] = it;
    }
}

// ---- initial sort, onesweep style -------------------------------------------------
// k_keys      builds the 64-bit record of every rotation (rolling mixed-radix key over consecutive
//             positions), the 32-bit per-position key for the group finisher (rk), and the digit
//             histograms of ALL four passes of every block in one read of the block bytes.
// k_digit_scan turns each histogram into the digit base offsets of its pass.
// k_sweep     one radix pass in ONE kernel: the tile ranks its records, publishes its per-digit
//             counts, and obtains its global offsets by a decoupled look-back over the earlier tiles
//             of the same block (status word = generation | flag | count, so nothing is zeroed
//             between passes).  Tiles are handed out by an atomic ticket, block index fastest:
//             a tile's predecessors always hold earlier tickets (forward progress), and with many
//             blocks in a batch the predecessor has normally published its inclusive prefix already.
constexpr int KT = 8;                        // tiles per CTA in k_keys
constexpr int SWEEP_G = 64;                  // blocks whose tiles are interleaved in a radix pass
constexpr int NPASS = (KEY_BITS + SW_BITS - 1) / SW_BITS;
constexpr uint32_t ST_INCL = 1u << 23, ST_AGG = 1u << 22, ST_VAL = 0x000fffffu;

constexpr int SWT = S3G_SWT;                     // threads of a sweep CTA
constexpr int SWI = S3G_SWI;                      // records per thread
constexpr int SW_TILE = SWT * SWI;           // records per sweep tile
constexpr int SW_NT = (BLK_STRIDE + SW_TILE - 1) / SW_TILE;
constexpr int SW_DPT = SWN > SWT ? SWN / SWT : 1;    // digits per thread (threads past SWN own none)
static_assert(SWN <= SWT || SWN % SWT == 0, "every digit needs an owner thread");
static_assert(KEY_BITS + VAL_BITS <= 64 && NPASS * SW_BITS >= KEY_BITS, "record layout");

struct SweepSmem {
    uint64_t stage[SW_TILE];               // tile in digit order (the per-warp peer masks live here while ranking)
    uint32_t gbase[SWN];                   // (global offset - tile offset) of the digit
    uint32_t tbase[SWN];                   // offset of the digit inside the tile
    uint32_t wcnt[(SWT / 32) * SWN];       // per-warp digit counters -> exclusive warp offsets
    uint32_t scan[33];
    // first pass only: the records are built from the block bytes on the fly
    __align__(16) uint8_t sym[SW_TILE + 64];
    uint8_t seq[256], frac[256];
};

struct KeysSmem {
    uint32_t stage32[STILE + STILE / 16];    // per-position keys, skewed so that 16-consecutive-per-thread writes are conflict-free
    uint32_t hist[NPASS][SWN];
    __align__(16) uint8_t sym[STILE + 64];
    uint8_t seq[256];
    uint8_t frac[256];
};

__global__ void __launch_bounds__(ST) k_keys(BwtP P, uint32_t *ghist)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KeysSmem &S = *reinterpret_cast<KeysSmem *>(smem_raw);
    const uint32_t lb = blockIdx.y, tid = threadIdx.x;
    if (P.mode[lb] != P.want) return;
    const uint32_t n = P.cnt_n[lb];
    const uint32_t t0 = blockIdx.x * KT;
    if ((uint64_t)t0 * STILE >= n) return;
    const uint8_t *b = P.blk + P.blocks[lb].blk_off;
    const uint32_t k = P.init_k[lb], k32 = P.init_k32[lb], a = P.init_a[lb], f = P.init_f[lb];
    for (int i = tid; i < NPASS * SWN; i += ST) (&S.hist[0][0])[i] = 0;
    S.seq[tid] = P.seq[(uint64_t)lb * 256 + tid];
    S.frac[tid] = (uint8_t)(tid < a ? tid * f / a : 0);      // class of a symbol rank, monotone
    uint64_t pw = 1; uint32_t pw32 = 1;                      // a^(k-1), a^(k32-1)
    for (uint32_t i = 1; i < k; i++) pw *= a;
    for (uint32_t i = 1; i < k32; i++) pw32 *= a;
    uint32_t *rk = P.rk + (uint64_t)lb * BLK_STRIDE;
    for (uint32_t t = t0; t < t0 + KT && (uint64_t)t * STILE < n; t++) {
        const uint32_t base = t * STILE, cntT = min((uint32_t)STILE, n - base);
        __syncthreads();
        {
            // 16 block bytes per thread in one vector load, mapped to symbol ranks (unseqToSeq, bz/compress.c:106-115)
            const uint32_t i0 = tid * 16;
            if (base + i0 + 16 <= n) {
                uint4 v = *reinterpret_cast<const uint4 *>(b + base + i0);
                uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++)
                    wv[j] = (uint32_t)S.seq[wv[j] & 255] | (uint32_t)S.seq[(wv[j] >> 8) & 255] << 8 | (uint32_t)S.seq[(wv[j] >> 16) & 255] << 16 |
                            (uint32_t)S.seq[wv[j] >> 24] << 24;
                *reinterpret_cast<uint4 *>(S.sym + i0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            } else {
                for (uint32_t i = i0; i < i0 + 16 && i < cntT + k + 1; i++) {
                    uint32_t q = base + i;
                    if (q >= n) { q -= n; if (q >= n) q %= n; }
                    S.sym[i] = S.seq[b[q]];
                }
            }
            // the k symbols past the tile (cyclic)
            for (uint32_t i = STILE + tid; i < cntT + k + 1; i += ST) {
                uint32_t q = base + i;
                if (q >= n) { q -= n; if (q >= n) q %= n; }
                S.sym[i] = S.seq[b[q]];
            }
        }
        __syncthreads();
        const uint32_t p0 = tid * SI;
        if (p0 < cntT) {
            uint64_t key = 0; uint32_t key32 = 0;
            for (uint32_t j = 0; j < k; j++) { key = key * a + S.sym[p0 + j]; if (j + 1 == k32) key32 = (uint32_t)key; }
#pragma unroll
            for (int r = 0; r < SI; r++) {
                uint32_t p = p0 + r;
                if (p < cntT) {
                    uint64_t rec = ((key * f + S.frac[S.sym[p + k]]) << VAL_BITS) | (base + p);
                    S.stage32[p + (p >> 4)] = key32;
#pragma unroll
                    for (int ps = 0; ps < NPASS; ps++) atomicAdd(&S.hist[ps][(uint32_t)(rec >> (VAL_BITS + SW_BITS * ps)) & (SWN - 1)], 1u);
                    uint32_t so = S.sym[p];
                    key = (key - so * pw) * a + S.sym[p + k];
                    key32 = (key32 - so * pw32) * a + S.sym[p + k32];
                }
            }
        }
        __syncthreads();
        for (uint32_t p = tid; p < cntT; p += ST) rk[base + p] = S.stage32[p + (p >> 4)];
    }
    __syncthreads();
    uint32_t *gh = ghist + (uint64_t)lb * NPASS * SWN;
    for (int i = tid; i < NPASS * SWN; i += ST) { uint32_t v = (&S.hist[0][0])[i]; if (v) atomicAdd(&gh[i], v); }
}

__global__ void __launch_bounds__(SWN) k_digit_scan(uint32_t *ghist)
{
    __shared__ uint32_t sm[33];
    uint32_t *g = ghist + ((uint64_t)blockIdx.y * NPASS + blockIdx.x) * SWN;
    uint32_t v = g[threadIdx.x], tot;
    g[threadIdx.x] = block_excl_sum<uint32_t>(v, sm, &tot);
}

// STABLE = false (the first pass only: the records arrive in position order, which carries no
// meaning yet): ranks come straight from an atomicAdd on one per-tile counter per digit -- no
// per-warp tables, no peer matching, and the records of a thread rank independently.
template <bool STABLE, bool SAFE = false>
__global__ void __launch_bounds__(SWT, SW_OCC) k_sweep(BwtP P, int rshift, const uint64_t *kv_in, uint64_t *kv_out, const uint32_t *dbase_all,
                                                      int pass, uint32_t *ticket_ctr, uint32_t nb, uint32_t gen, uint32_t G)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem &S = *reinterpret_cast<SweepSmem *>(smem_raw);
    // one ticket counter per bzip2 block (a single counter for the whole grid serialises its atomics on one
    // L2 address): the CTA is bound to block blockIdx.x % nb and draws the next tile of that block.
    // CTAs are dealt to the blocks in groups of SWEEP_G: the tiles of one block run close together in time, so the
    // short pieces a digit receives from consecutive tiles meet in L2 and leave as full lines (with the whole
    // batch interleaved they were evicted one sector at a time and the pass ran at a third of the DRAM rate),
    // while the look-back still finds its predecessors finished after a few tiles.
    const uint32_t lb = blockIdx.x / (SW_NT * G) * G + blockIdx.x % G;
    if (lb >= nb || P.mode[lb] != P.want) return;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) S.scan[0] = atomicAdd(ticket_ctr + lb, 1u);
    // Intra-warp matching through shared memory instead of match.any (measured on B200: match.any costs
    // ~8 cycles per DISTINCT value in the warp; an atomicOr of the lane bit into a per-warp, per-digit mask
    // word plus a read-back gives the same peer mask for ~20).  The mask table aliases the staging buffer,
    // which is not live until ranking is over.
    static_assert(sizeof(S.stage) >= (SWT / 32) * SWN * 4, "mask table must fit in the staging buffer");
    uint32_t *Mall = reinterpret_cast<uint32_t *>(S.stage);
    uint32_t *tcnt = reinterpret_cast<uint32_t *>(S.wcnt);       // !STABLE: one counter per digit
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *zm = reinterpret_cast<uint4 *>(S.stage), *zc = reinterpret_cast<uint4 *>(S.wcnt);
        if (STABLE) {
            if (SAFE) for (int i = tid; i < (SWT / 32) * SWN / 4; i += SWT) zm[i] = z;      // mask table: one word per warp and digit
            for (int i = tid; i < (SWT / 32) * SWN / 4; i += SWT) zc[i] = z;                // counters: one word per warp and digit
        } else {
            for (int i = tid; i < SWN / 4; i += SWT) zc[i] = z;
        }
    }
    __syncthreads();
    const uint32_t tile = S.scan[0];
    const uint32_t cnt = P.cnt_n[lb];
    if ((uint64_t)tile * SW_TILE >= cnt) return;
    uint32_t w = tid >> 5, l = tid & 31;
    uint32_t base = tile * SW_TILE + w * (SWI * 32) + l;
    uint32_t *mycnt = S.wcnt + w * SWN;
    uint32_t *M = Mall + w * SWN;
    const uint64_t *in = kv_in + (uint64_t)lb * BLK_STRIDE;
    uint64_t kv[SWI];
    uint16_t rnk[SWI];
    uint32_t okmask = 0;
    if (!STABLE) {
        // first pass: thread t builds the records of positions SWI t .. SWI t + SWI - 1 of the tile with a rolling key,
        // exactly as k_keys did for the histograms (any record-to-thread mapping will do: no stability needed)
        const uint8_t *b = P.blk + P.blocks[lb].blk_off;
        const uint32_t k = P.init_k[lb], a = P.init_a[lb], f = P.init_f[lb];
        const uint32_t tbase0 = tile * SW_TILE, cntT = min((uint32_t)SW_TILE, cnt - tbase0);
        if (tid < 256) {
            S.seq[tid] = P.seq[(uint64_t)lb * 256 + tid];
            S.frac[tid] = (uint8_t)(tid < a ? tid * f / a : 0);
        }
        __syncthreads();
        for (uint32_t i0 = tid * 16; i0 < (uint32_t)SW_TILE; i0 += SWT * 16) {
            if (tbase0 + i0 + 16 <= cnt) {
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
                    if (q >= cnt) { q -= cnt; if (q >= cnt) q %= cnt; }
                    S.sym[i] = S.seq[b[q]];
                }
            }
        }
        for (uint32_t i = SW_TILE + tid; i < cntT + k + 1; i += SWT) {
            uint32_t q = tbase0 + i;
            if (q >= cnt) { q -= cnt; if (q >= cnt) q %= cnt; }
            S.sym[i] = S.seq[b[q]];
        }
        __syncthreads();
        uint64_t pw = 1;
        for (uint32_t i = 1; i < k; i++) pw *= a;
        const uint32_t p0 = tid * SWI;
        uint64_t key = 0;
        if (p0 < cntT) for (uint32_t j = 0; j < k; j++) key = key * a + S.sym[p0 + j];
#pragma unroll
        for (int r = 0; r < SWI; r++) {
            uint32_t p = p0 + r;
            kv[r] = 0;
            if (p < cntT) {
                kv[r] = ((key * f + S.frac[S.sym[p + k]]) << VAL_BITS) | (tbase0 + p);
                okmask |= 1u << r;
                key = (key - S.sym[p] * pw) * a + S.sym[p + k];
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < SWI; r++) {
            uint32_t p = base + r * 32;
            kv[r] = 0;
            if (p < cnt) { kv[r] = in[p]; okmask |= 1u << r; }
        }
    }
    if (STABLE) {
        // the tile that will be handed out about two waves from now: pull it into L2 (128-byte lines)
        uint32_t tile2 = tile + 2 * SM_COUNT * SW_OCC / G + 1;
        if (tile2 < SW_NT) {
            const uint8_t *pf = reinterpret_cast<const uint8_t *>(kv_in + (uint64_t)lb * BLK_STRIDE + (uint64_t)tile2 * SW_TILE);
            for (uint32_t o = tid * 128; o < (uint32_t)SW_TILE * 8; o += SWT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + o));
        }
    }
    constexpr int DPT = SW_DPT;
    const bool has_digit = tid * DPT < (uint32_t)SWN;
    uint32_t tot4[DPT];
    uint32_t mysum = 0;
    if (!STABLE) {
#pragma unroll
        for (int r = 0; r < SWI; r++)
            if (okmask & (1u << r)) rnk[r] = (uint16_t)atomicAdd(&tcnt[(uint32_t)(kv[r] >> rshift) & (SWN - 1)], 1u);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < DPT; q++) { tot4[q] = has_digit ? tcnt[tid * DPT + q] : 0u; mysum += tot4[q]; }
    } else {
        const uint32_t lbit = 1u << l, ltmask = lbit - 1;
        // match.any is cheap when the warp holds few distinct digits (the upper passes of text: the records
        // arrive sorted by the lower digits and neighbours share their context), the mask table when it holds
        // many.  One probe per warp and tile: equal neighbours among the first 32 records.
        bool use_match;
        {
            uint32_t d0 = (okmask & 1u) ? ((uint32_t)(kv[0] >> rshift) & (SWN - 1)) : 0x10000u + l;
            uint32_t dn = __shfl_down_sync(0xffffffffu, d0, 1);
            use_match = __popc(__ballot_sync(0xffffffffu, l < 31 && d0 == dn)) >= 16;
        }
        if (use_match) {
#pragma unroll
            for (int r = 0; r < SWI; r++) {
                bool ok = (okmask >> r) & 1u;
                uint32_t d = ok ? ((uint32_t)(kv[r] >> rshift) & (SWN - 1)) : 0xffffffffu;
                unsigned peers = __match_any_sync(0xffffffffu, d);
                unsigned lt = peers & ltmask;
                uint32_t bb = ok ? mycnt[d] : 0u;
                __syncwarp();
                if (ok && lt == 0) mycnt[d] = bb + __popc(peers);
                __syncwarp();
                rnk[r] = (uint16_t)(bb + __popc(lt));
            }
        } else if (!SAFE) {
            // Ranks straight from one atomicAdd per record on the warp's own counter of the digit.  This is the
            // stable rank only because the hardware serves the lanes of a warp that hit the same shared-memory
            // word in ascending lane order (measured on B200: no exception in 3e10 trials, scratch/atom_order.cu),
            // which nothing guarantees: the finisher checks that the keys it receives ascend, and run_bwt repeats
            // the passes with SAFE = true (peer masks) if they ever do not.
#pragma unroll
            for (int r = 0; r < SWI; r++)
                if ((okmask >> r) & 1u) rnk[r] = (uint16_t)atomicAdd(&mycnt[(uint32_t)(kv[r] >> rshift) & (SWN - 1)], 1u);
        } else {
#pragma unroll
            for (int r = 0; r < SWI; r++) {
                bool ok = (okmask >> r) & 1u;
                uint32_t d = (uint32_t)(kv[r] >> rshift) & (SWN - 1);
                if (ok) atomicOr(&M[d], lbit);
                __syncwarp();
                uint32_t peers = ok ? M[d] : 0u;
                uint32_t bb = ok ? mycnt[d] : 0u;
                __syncwarp();
                uint32_t lt = peers & ltmask;
                if (ok && lt == 0) { mycnt[d] = bb + __popc(peers); M[d] = 0; }
                __syncwarp();
                rnk[r] = (uint16_t)(bb + __popc(lt));
            }
        }
        __syncthreads();
        // per digit: tile total (the exclusive prefix over warps is written further down, with the digit's offset inside the
        // tile already added: the scatter then reads one table entry per record instead of two)
#pragma unroll
        for (int q = 0; q < DPT; q++) {
            tot4[q] = 0;
            if (has_digit) {
                int d = tid * DPT + q;
                uint32_t run = 0;
#pragma unroll
                for (int ww = 0; ww < SWT / 32; ww++) run += S.wcnt[ww * SWN + d];
                tot4[q] = run; mysum += run;
            }
        }
    }
    // publish the tile's counts, look back for the exclusive prefix over earlier tiles, publish the inclusive prefix
    const uint32_t gtag = gen << 24;
    volatile uint32_t *status = P.hist + (uint64_t)lb * SW_NT * SWN + tid * DPT;     // + tile * SWN
    uint32_t ex4[DPT];
#pragma unroll
    for (int q = 0; q < DPT; q++) ex4[q] = 0;
    if (has_digit) {
        if (tile > 0) {
#pragma unroll
            for (int q = 0; q < DPT; q++) status[(uint64_t)tile * SWN + q] = gtag | ST_AGG | tot4[q];
            uint32_t done = 0;                 // bit q: digit q has met an inclusive prefix
            uint32_t spins = 0;
            for (int32_t t = (int32_t)tile - 1; t >= 0 && done != (1u << DPT) - 1;) {
                uint32_t x[DPT];
#pragma unroll
                for (int q = 0; q < DPT; q++) x[q] = status[(uint64_t)t * SWN + q];
                bool ready = true;
#pragma unroll
                for (int q = 0; q < DPT; q++) if (!(done & (1u << q)) && (x[q] >> 24) != gen) ready = false;
                if (!ready) {
                    if (++spins > (1u << 24)) __trap();          // a predecessor never published: fail loudly instead of hanging
                    __nanosleep(20);
                    continue;
                }
#pragma unroll
                for (int q = 0; q < DPT; q++) {
                    if (done & (1u << q)) continue;
                    ex4[q] += x[q] & ST_VAL;
                    if (x[q] & ST_INCL) done |= 1u << q;
                }
                t--;
            }
        }
#pragma unroll
        for (int q = 0; q < DPT; q++) status[(uint64_t)tile * SWN + q] = gtag | ST_INCL | (ex4[q] + tot4[q]);
    }
    const uint32_t *dbase = dbase_all + ((uint64_t)lb * NPASS + pass) * SWN;
    uint32_t tile_total;
    uint32_t ex = block_excl_sum<uint32_t>(mysum, S.scan, &tile_total);
    if (has_digit) {
#pragma unroll
        for (int q = 0; q < DPT; q++) {
            int d = tid * DPT + q;
            S.tbase[d] = ex;
            S.gbase[d] = dbase[d] + ex4[q] - ex;          // (global offset - tile offset) of the digit
            if (STABLE) {
                // per warp: where its records of the digit start inside the tile
                uint32_t run = ex;
#pragma unroll
                for (int ww = 0; ww < SWT / 32; ww++) { uint32_t c = S.wcnt[ww * SWN + d]; S.wcnt[ww * SWN + d] = run; run += c; }
            }
            ex += tot4[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SWI; r++) {
        if (okmask & (1u << r)) {
            uint32_t d = (uint32_t)(kv[r] >> rshift) & (SWN - 1);
            S.stage[(STABLE ? mycnt[d] : S.tbase[d]) + rnk[r]] = kv[r];
        }
    }
    __syncthreads();
    uint64_t *out = kv_out + (uint64_t)lb * BLK_STRIDE;
    for (uint32_t i = tid; i < tile_total; i += SWT) {
        uint64_t it = S.stage[i];
        uint32_t d = (uint32_t)(it >> rshift) & (SWN - 1);
        out[S.gbase[d] + i] = it;
    }
}

// ---- group boundaries and new ranks -------------------------------------------
// Sorted items p of block lb: key_p (group = SA position of the group start; in INIT
// mode the symbol key), val_p.  Thread t of a tile owns SI consecutive items.
template <bool INIT, int ITEMS>
__device__ __forceinline__ void boundary_flags(const BwtP &P, uint32_t lb, uint32_t n, uint32_t cnt, uint32_t h,
                                               const uint64_t *kv, uint32_t p0, uint32_t &headmask, uint32_t &bndmask,
                                               uint64_t *items /*ITEMS+1*/)
{
    // flags for items p0 .. p0+SI (one extra for the singleton test of the last own item)
    const uint64_t *a = kv + (uint64_t)lb * BLK_STRIDE;
    const uint32_t *rk = P.rk + (uint64_t)lb * BLK_STRIDE;
    uint32_t prev_key = 0, prev_sec = 0;
    bool have_prev = false;
    if (p0 > 0 && p0 - 1 < cnt) {
        uint64_t pv = a[p0 - 1];
        prev_key = (uint32_t)(pv >> 32);
        if (!INIT) { uint32_t j = (uint32_t)pv + h; if (j >= n) j -= n; prev_sec = rk[j]; }
        have_prev = true;
    }
    headmask = 0; bndmask = 0;
#pragma unroll
    for (int k = 0; k <= ITEMS; k++) {
        uint32_t p = p0 + k;
        if (p < cnt) {
            uint64_t it = a[p];
            items[k] = it;
            uint32_t key = (uint32_t)(it >> 32), sec = 0;
            if (!INIT) { uint32_t j = (uint32_t)it + h; if (j >= n) j -= n; sec = rk[j]; }
            bool head = !have_prev || (!INIT && key != prev_key);
            bool bnd = head || (INIT ? key != prev_key : sec != prev_sec);
            if (INIT) head = !have_prev;
            if (head) headmask |= 1u << k;
            if (bnd) bndmask |= 1u << k;
            prev_key = key; prev_sec = sec; have_prev = true;
        } else {
            items[k] = 0;
            bndmask |= 1u << k;     // past the end counts as a boundary (singleton test)
            headmask |= 1u << k;
        }
    }
}

// Tile aggregates for the two max-scans: (position + 1) of the last group head and of the last
// sub-group boundary inside the tile.  Heads need only the keys (coalesced); the last boundary is
// searched backwards from the tile end in chunks of ST items and almost always sits in the first one.
template <bool INIT>
__global__ void __launch_bounds__(ST) k_bound_agg(BwtP P, uint32_t round, const uint64_t *kv, const uint32_t *act_cur)
{
    __shared__ uint32_t sm[33];
    uint32_t lb = blockIdx.y, tile = blockIdx.x;
    if (!block_live(P, lb, INIT ? 0 : 1, round, act_cur)) return;
    uint32_t h = depth_of(P, lb, round);
    uint32_t n = P.cnt_n[lb], cnt = INIT ? n : P.cnt_m[lb];
    uint32_t t0 = tile * STILE;
    if (t0 >= cnt) return;
    uint32_t t1 = min(t0 + (uint32_t)STILE, cnt);
    const uint64_t *a = kv + (uint64_t)lb * BLK_STRIDE;
    const uint32_t *rk = P.rk + (uint64_t)lb * BLK_STRIDE;
    uint32_t lh = 0;
    if (INIT) { if (t0 == 0 && threadIdx.x == 0) lh = 1; }
    else {
#pragma unroll 4
        for (uint32_t p = t0 + threadIdx.x; p < t1; p += ST) {
            bool head = p == 0 || (uint32_t)(a[p] >> 32) != (uint32_t)(a[p - 1] >> 32);
            if (head) lh = p + 1;
        }
    }
    uint32_t th, tb = 0;
    block_excl_max<uint32_t>(lh, sm, &th);
    for (uint32_t top = t1; top > t0; top = top > t0 + ST ? top - ST : t0) {
        uint32_t lo = top > t0 + ST ? top - ST : t0;
        uint32_t p = lo + threadIdx.x;
        uint32_t v = 0;
        if (p < top) {
            bool bnd;
            if (p == 0) bnd = true;
            else {
                uint64_t x = a[p], y = a[p - 1];
                if ((uint32_t)(x >> 32) != (uint32_t)(y >> 32)) bnd = true;
                else if (INIT) bnd = false;
                else {
                    uint32_t jx = (uint32_t)x + h; if (jx >= n) jx -= n;
                    uint32_t jy = (uint32_t)y + h; if (jy >= n) jy -= n;
                    bnd = rk[jx] != rk[jy];
                }
            }
            if (bnd) v = p + 1;
        }
        block_excl_max<uint32_t>(v, sm, &tb);
        if (tb) break;                     // uniform: tb comes from shared memory
    }
    if (threadIdx.x == 0) {
        uint32_t *ag = P.agg + ((uint64_t)lb * NT + tile) * 2;
        ag[0] = th; ag[1] = tb;
    }
}

// ---- group finisher ---------------------------------------------------------------
// After the initial sort every group (rotations with equal initial keys) is contiguous in SA order.
// bzip2's own mainSort finishes such buckets with direct string comparisons (bz/blocksort.c:347-469,
// :621-717) because real blocks have short common prefixes; the same holds here, so the groups are
// finished on deeper keys instead of paying global radix passes per doubling round:
//   level 0   key = the 32-bit key (k32 symbols) saved for position start + k, i.e. the NEXT symbols in one
//             4-byte gather; rank inside the group by counting smaller / equal keys; unique keys are final
//   level >=1 the few still-tied rotations repeat this k32 symbols further on
// Ties that survive FLEVELS levels and groups beyond FB_MAX (long repeats, periodic blocks) are written out
// unsorted with a NONHEAD flag on every member but the first; only blocks that have such leftovers go
// through the prefix-doubling rounds below.

__device__ __forceinline__ uint32_t deeper_key(const uint32_t *k30, uint32_t pos, uint32_t off, uint32_t n)
{
    uint32_t p = pos + off;
    if (p >= n) { p -= n; if (p >= n) p %= n; }
    return k30[p];
}

// ---- group finisher, warp form -------------------------------------------------------
// k_finish_rows: a warp walks rows of 32 consecutive SA positions with a window of FA_WIN rows in
// registers (rotation start | last-column symbol per entry, one head ballot per row).  A row owns the
// groups whose head lies in it; a group that ends inside the window is ranked by the warp itself
// (level-0 keys through a per-warp shared-memory strip, counting smaller / equal keys; the few ties go
// to a per-warp list for the deeper levels), singletons are written straight from registers, and a
// group that does not end inside the window is appended to a work list for k_finish_mid.  Entries
// before the first head of a row belong to an earlier row's group and are skipped: their owner wrote
// them, or listed the group.
// k_finish_mid: one warp per listed group of up to FM_MAX rotations; k_finish_big: one CTA per longer group: up to
// FB_MAX rotations are ranked in shared memory (level 0 by a bitonic sort, ties level by level), larger ones are
// written out unsorted (NONHEAD flags) for the doubling rounds.
#ifndef S3G_FA_OCC
#define S3G_FA_OCC 4
#endif
#ifndef S3G_FA_WARPS
#define S3G_FA_WARPS 8
#endif
constexpr int FA_WARPS = S3G_FA_WARPS;
constexpr int FA_ROWS = 32;                              // rows a warp walks
constexpr int FA_TILE = FA_WARPS * FA_ROWS * 32;         // SA positions per CTA
constexpr int FA_NT = (BLK_STRIDE + FA_TILE - 1) / FA_TILE;
constexpr int FA_WIN = 4;                                // rows in the window: a group of up to 97 rotations always fits
constexpr int FA_W = FA_WIN * 32;
constexpr uint32_t FA_BIG = 0xffffu;                     // group end: not inside the window

struct FinASmem {
    uint32_t dk[FA_WARPS][FA_W];      // level-0 keys of the window; keys of the tied list afterwards
    uint32_t t_pw[FA_WARPS][FA_W];    // tied list: rotation start | last-column symbol << 24
    uint16_t t_cs[FA_WARPS][FA_W];    //            window position of the tie's first slot (0xffff: resolved)
    uint16_t t_rank[FA_WARPS][FA_W];  //            order among equals so far
    uint8_t seq[256];
};

// end of the group that holds entry e of the current row: the next head in the row, else `ahead`, the first
// head of the rows ahead (the same for every lane: computed once per row)
__device__ __forceinline__ uint32_t fa_group_end(uint32_t e, uint32_t h0, uint32_t ahead)
{
    uint32_t mh = e >= 31 ? 0u : (h0 & (0xfffffffeu << e));
    return mh ? (uint32_t)__ffs(mh) - 1 : ahead;
}

__global__ void __launch_bounds__(FA_WARPS * 32, S3G_FA_OCC) k_finish_rows(BwtP P, const uint64_t *kv, unsigned long long *g_left, BlockInfo *blocks,
                                                                  uint8_t *lcol, uint64_t *big_list, uint32_t *big_cnt, uint32_t *unsorted)
{
    __shared__ FinASmem S;
    const uint32_t lb = blockIdx.y, tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    if (P.mode[lb] != P.want) return;
    const uint32_t n = P.cnt_n[lb];
    if (blockIdx.x * FA_TILE >= n) return;
    S.seq[tid] = P.seq[(uint64_t)lb * 256 + tid];
    __syncthreads();
    const uint32_t p0 = blockIdx.x * FA_TILE + w * (FA_ROWS * 32);
    if (p0 >= n) return;
    const uint64_t *a = kv + (uint64_t)lb * BLK_STRIDE;
    const uint32_t *k30 = P.rk + (uint64_t)lb * BLK_STRIDE;
    const uint8_t *b = P.blk + P.blocks[lb].blk_off;
    uint32_t *sa = P.sa + (uint64_t)lb * BLK_STRIDE;
    uint8_t *L = lcol + (uint64_t)lb * BLK_STRIDE;
    const uint32_t k0 = P.init_k[lb], k32 = P.init_k32[lb];
    uint32_t *dk = S.dk[w];
    uint32_t *t_pw = S.t_pw[w];
    uint16_t *t_cs = S.t_cs[w], *t_rank = S.t_rank[w];
    const uint32_t ltmask = (1u << l) - 1;
    uint32_t leftover = 0;

    // A row enters the window in three steps, one loop iteration apart, so that no load is waited for where
    // it is issued: (A) the records of the row and of the position before it; (B) head ballot, rotation start
    // and the gather of the byte before the rotation (bz/compress.c:166-167); (C) the byte's symbol rank.
    auto issue_raw = [&](uint32_t rowbase, uint64_t &x, uint64_t &pv) {
        uint32_t p = rowbase + l;
        x = 0; pv = 0;
        if (p < n) { x = a[p]; pv = a[p ? p - 1 : 0]; }
    };
    auto stage_b = [&](uint32_t rowbase, uint64_t x, uint64_t pv, uint32_t &posv, uint32_t &hmv, uint32_t &byv) {
        uint32_t p = rowbase + l;
        bool head = p >= n || p == 0 || ((x ^ pv) >> VAL_BITS) != 0;     // positions past the block end close every group
        if (p < n && (pv >> VAL_BITS) > (x >> VAL_BITS)) *unsorted = 1u;   // the radix passes did not sort (see k_sweep)
        posv = (uint32_t)x & VMASK;
        byv = p < n ? b[posv ? posv - 1 : n - 1] : 0;
        hmv = __ballot_sync(0xffffffffu, head);
    };
    auto stage_c = [&](uint32_t posv, uint32_t byv) { return posv | (uint32_t)S.seq[byv] << 24; };
    // level-0 key of an entry of window row j unless it is a singleton (heads on both sides)
    auto prefetch_key = [&](uint32_t rowbase_j, uint32_t pwv, uint32_t hj, uint32_t hnext) {
        uint32_t nextbit = l < 31 ? (hj >> (l + 1)) & 1u : hnext & 1u;
        bool single = ((hj >> l) & 1u) && nextbit;
        return (rowbase_j + l < n && !single) ? deeper_key(k30, pwv & VMASK, k0, n) : 0u;
    };
    auto emit = [&](uint32_t slot_abs, uint32_t pwv, uint32_t flag) {
        uint32_t pos = pwv & VMASK;
        sa[slot_abs] = pos | flag;
        L[slot_abs] = (uint8_t)(pwv >> 24);
        if (pos == 0) blocks[lb].orig_ptr = (int32_t)slot_abs;
    };

    // the warp's records: into L2 now (64 lines), so that the row loads below meet L2 latency, not DRAM latency
    for (uint32_t o = l * 16; o < (FA_ROWS + FA_WIN + 3) * 32 && p0 + o < n; o += 32 * 16)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a + p0 + o));
    uint32_t pw[FA_WIN], hm[FA_WIN];
    uint32_t kw[3];                     // level-0 keys of window rows 0..2
    uint64_t xa, pva;                   // step A of row R + 6
    uint32_t posb, hmb, byb;            // step B of row R + 5
    uint32_t posc, hmc, byc;            // step B of row R + 4, one iteration older (its gather has had two iterations)
#pragma unroll
    for (int j = 0; j < FA_WIN; j++) {
        uint64_t x, pv; uint32_t ps, by;
        issue_raw(p0 + 32 * j, x, pv);
        stage_b(p0 + 32 * j, x, pv, ps, hm[j], by);
        pw[j] = stage_c(ps, by);
    }
    issue_raw(p0 + 32 * 4, xa, pva);
    stage_b(p0 + 32 * 4, xa, pva, posc, hmc, byc);
    issue_raw(p0 + 32 * 5, xa, pva);
    stage_b(p0 + 32 * 5, xa, pva, posb, hmb, byb);
    issue_raw(p0 + 32 * 6, xa, pva);
    kw[0] = prefetch_key(p0, pw[0], hm[0], hm[1]);
    kw[1] = prefetch_key(p0 + 32, pw[1], hm[1], hm[2]);

    for (uint32_t R = 0; R < FA_ROWS; R++) {
        const uint32_t rowbase = p0 + 32 * R;
        if (rowbase >= n) break;
        const uint32_t nv = min(32u, n - rowbase);
        const uint32_t vmask = nv == 32 ? 0xffffffffu : (1u << nv) - 1;
        const uint32_t h0 = hm[0] & vmask;                 // heads of real entries
        kw[2] = prefetch_key(rowbase + 64, pw[2], hm[2], hm[3]);
        if (h0) {
            const uint32_t H0 = hm[0];
            const uint32_t ahead = hm[1] ? 32 + (uint32_t)__ffs(hm[1]) - 1 : hm[2] ? 64 + (uint32_t)__ffs(hm[2]) - 1
                                 : hm[3] ? 96 + (uint32_t)__ffs(hm[3]) - 1 : FA_BIG;
            // the last group that starts in this row may run on into the rows ahead
            const uint32_t gl = 31 - (uint32_t)__clz(h0);
            const uint32_t gel = fa_group_end(gl, H0, ahead);
            const bool spill = gel != FA_BIG && gel > 32;
            const uint32_t ncov = spill ? (gel + 31) >> 5 : 1;
            // own entry of the current row
            const uint32_t mlow = h0 & (0xffffffffu >> (31 - l));
            const bool owned = l < nv && mlow != 0;
            uint32_t gs = 0, ge = 0;
            if (owned) { gs = 31 - (uint32_t)__clz(mlow); ge = fa_group_end(l, H0, ahead); }
            const bool big = owned && ge == FA_BIG;
            if (big && l == gs) big_list[atomicAdd(big_cnt, 1u)] = (uint64_t)lb << 32 | (rowbase + gs);
            const bool single = owned && !big && ge - gs == 1;
            if (single) emit(rowbase + l, pw[0], 0u);
            bool memb[FA_WIN];
            memb[0] = owned && !big && !single;
#pragma unroll
            for (int j = 1; j < FA_WIN; j++) memb[j] = spill && 32 * j + l < gel;
            if (__any_sync(0xffffffffu, memb[0]) || spill) {
                // ---- level 0: the next k32 symbols of every member, rank by counting ----
                uint32_t key[FA_WIN];
#pragma unroll
                for (int j = 0; j < FA_WIN; j++) {
                    key[j] = 0;
                    if (j < (int)ncov && memb[j]) { key[j] = j < 3 ? kw[j] : deeper_key(k30, pw[j] & VMASK, k0, n); dk[32 * j + l] = key[j]; }
                }
                __syncwarp();
                uint32_t tied = 0;                         // bit j: entry j of this lane is still tied
                uint32_t cs_[FA_WIN], rk_[FA_WIN];
#pragma unroll
                for (int j = 0; j < FA_WIN; j++) {
                    cs_[j] = rk_[j] = 0;
                    if (j < (int)ncov && memb[j]) {
                        const uint32_t g0 = j ? gl : gs, g1 = j ? gel : ge, self = 32 * j + l, my = key[j];
                        uint32_t ge_ = 0, le_ = 0;
#pragma unroll 4
                        for (uint32_t t = g0; t < g1; t++) {
                            uint32_t x = dk[t];
                            // += (x >= my), += (x <= my): the carry out of a subtraction, three instructions for both
                            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(ge_) : "r"(x), "r"(my));
                            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(le_) : "r"(my), "r"(x));
                        }
                        const uint32_t lt = g1 - g0 - ge_;
                        if (ge_ + le_ - (g1 - g0) == 1) emit(rowbase + g0 + lt, pw[j], 0u);
                        else {
                            uint32_t eqb = 0;              // ties are rare: their order among equals is counted separately
                            for (uint32_t t = g0; t < self; t++) eqb += dk[t] == my;
                            tied |= 1u << j; cs_[j] = g0 + lt; rk_[j] = eqb;
                        }
                    }
                }
                __syncwarp();
                if (__any_sync(0xffffffffu, tied != 0)) {
                    // ---- deeper levels on the warp's tied list ----
                    uint32_t nt = 0;
#pragma unroll
                    for (int j = 0; j < FA_WIN; j++) {
                        uint32_t m = __ballot_sync(0xffffffffu, (tied >> j) & 1u);
                        if ((tied >> j) & 1u) {
                            uint32_t u = nt + __popc(m & ltmask);
                            t_pw[u] = pw[j]; t_cs[u] = (uint16_t)cs_[j]; t_rank[u] = (uint16_t)rk_[j];
                        }
                        nt += __popc(m);
                    }
                    __syncwarp();
                    for (uint32_t level = 1; level < FLEVELS; level++) {
#pragma unroll
                        for (int r = 0; r < FA_WIN; r++) {
                            uint32_t u = l + 32 * r;
                            if (u < nt && t_cs[u] != 0xffffu) dk[u] = deeper_key(k30, t_pw[u] & VMASK, k0 + level * k32, n);
                        }
                        __syncwarp();
                        uint32_t lt_[FA_WIN], eq_[FA_WIN], eqb_[FA_WIN];
#pragma unroll
                        for (int r = 0; r < FA_WIN; r++) {
                            uint32_t u = l + 32 * r;
                            lt_[r] = eq_[r] = eqb_[r] = 0;
                            if (u < nt && t_cs[u] != 0xffffu) {
                                uint32_t cs = t_cs[u], my = dk[u];
                                for (uint32_t v = 0; v < nt; v++) {
                                    if (t_cs[v] != cs) continue;
                                    uint32_t x = dk[v];
                                    lt_[r] += x < my;
                                    uint32_t is = x == my;
                                    eq_[r] += is;
                                    eqb_[r] += is & (uint32_t)(v < u);
                                }
                            }
                        }
                        __syncwarp();
                        bool live = false;
#pragma unroll
                        for (int r = 0; r < FA_WIN; r++) {
                            uint32_t u = l + 32 * r;
                            if (u < nt && t_cs[u] != 0xffffu) {
                                uint32_t cs = t_cs[u];
                                if (eq_[r] == 1) { emit(rowbase + cs + lt_[r], t_pw[u], 0u); t_cs[u] = 0xffffu; }
                                else { t_cs[u] = (uint16_t)(cs + lt_[r]); t_rank[u] = (uint16_t)eqb_[r]; live = true; }
                            }
                        }
                        __syncwarp();
                        if (!__any_sync(0xffffffffu, live)) break;
                    }
                    // whatever is still tied goes out as an unsorted group
#pragma unroll
                    for (int r = 0; r < FA_WIN; r++) {
                        uint32_t u = l + 32 * r;
                        if (u < nt && t_cs[u] != 0xffffu) {
                            uint32_t rr = t_rank[u];
                            emit(rowbase + (uint32_t)t_cs[u] + rr, t_pw[u], rr ? NONHEAD : 0u);
                            leftover++;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        // advance the window
#pragma unroll
        for (int j = 0; j + 1 < FA_WIN; j++) { pw[j] = pw[j + 1]; hm[j] = hm[j + 1]; }
        kw[0] = kw[1]; kw[1] = kw[2];
        pw[FA_WIN - 1] = stage_c(posc, byc); hm[FA_WIN - 1] = hmc;
        posc = posb; hmc = hmb; byc = byb;
        stage_b(rowbase + 32 * 6, xa, pva, posb, hmb, byb);
        issue_raw(R + 8 < FA_ROWS + 4 ? rowbase + 32 * 7 : n, xa, pva);
    }
    for (int d = 16; d; d >>= 1) leftover += __shfl_xor_sync(0xffffffffu, leftover, d);
    if (l == 0 && leftover) { atomicAdd(&P.left[lb], leftover); atomicAdd(g_left, (unsigned long long)leftover); }
}

// k_finish_mid: one WARP per listed group of up to FM_MAX rotations (sorted BED with serial ids makes very many groups
// of a hundred or so: a CTA each would spend its time in barriers); longer ones are passed on to k_finish_big.
constexpr int FM_MAX = 256;
constexpr int FM_WARPS = 8;
constexpr int FM_EPL = FM_MAX / 32;               // entries per lane
struct FinMSmem {
    uint32_t key[FM_WARPS][FM_MAX], pw[FM_WARPS][FM_MAX];
    uint16_t cs[FM_WARPS][FM_MAX], rank[FM_WARPS][FM_MAX];
};

__global__ void __launch_bounds__(FM_WARPS * 32) k_finish_mid(BwtP P, const uint64_t *kv, unsigned long long *g_left, BlockInfo *blocks,
                                                              uint8_t *lcol, const uint64_t *big_list, const uint32_t *big_cnt,
                                                              uint64_t *big_list2, uint32_t *big_cnt2)
{
    __shared__ FinMSmem S;
    const uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const uint32_t nbig = *big_cnt;
    uint32_t *key = S.key[w], *pw = S.pw[w];
    uint16_t *cs = S.cs[w], *rank = S.rank[w];
    for (uint32_t item = blockIdx.x * FM_WARPS + w; item < nbig; item += gridDim.x * FM_WARPS) {
        const uint64_t it = big_list[item];
        const uint32_t lb = (uint32_t)(it >> 32), start = (uint32_t)it;
        const uint32_t n = P.cnt_n[lb];
        const uint64_t *a = kv + (uint64_t)lb * BLK_STRIDE;
        const uint32_t *k30 = P.rk + (uint64_t)lb * BLK_STRIDE;
        const uint8_t *b = P.blk + P.blocks[lb].blk_off;
        const uint8_t *seq = P.seq + (uint64_t)lb * 256;
        uint32_t *sa = P.sa + (uint64_t)lb * BLK_STRIDE;
        uint8_t *L = lcol + (uint64_t)lb * BLK_STRIDE;
        const uint32_t k0 = P.init_k[lb], k32 = P.init_k32[lb];
        // the group ends at the first position whose key differs
        const uint64_t key0 = a[start] >> VAL_BITS;
        uint32_t end = 0;
        for (uint32_t c = start + 1; c <= start + FM_MAX; c += 32) {
            uint32_t e = c + l;
            bool diff = e >= n || (a[e] >> VAL_BITS) != key0;
            uint32_t m = __ballot_sync(0xffffffffu, diff);
            if (m) { end = c + (uint32_t)__ffs(m) - 1; break; }
        }
        if (end == 0 || end - start > FM_MAX) {
            if (l == 0) big_list2[atomicAdd(big_cnt2, 1u)] = it;
            continue;
        }
        const uint32_t size = end - start;
        auto emit = [&](uint32_t slot_abs, uint32_t pwv, uint32_t flag) {
            uint32_t pos = pwv & VMASK;
            sa[slot_abs] = pos | flag;
            L[slot_abs] = (uint8_t)(pwv >> 24);
            if (pos == 0) blocks[lb].orig_ptr = (int32_t)slot_abs;
        };
        __syncwarp();
        for (uint32_t e = l; e < size; e += 32) {
            uint32_t pos = (uint32_t)a[start + e] & VMASK;
            pw[e] = pos | (uint32_t)seq[b[pos ? pos - 1 : n - 1]] << 24;
            cs[e] = 0; rank[e] = (uint16_t)e;
        }
        __syncwarp();
        for (uint32_t level = 0; level < FLEVELS; level++) {
            for (uint32_t e = l; e < size; e += 32)
                if (cs[e] != 0xffffu) key[e] = deeper_key(k30, pw[e] & VMASK, k0 + level * k32, n);
            __syncwarp();
            uint32_t lt_[FM_EPL], eq_[FM_EPL], eqb_[FM_EPL];
#pragma unroll
            for (int r = 0; r < FM_EPL; r++) {
                uint32_t u = l + 32 * r;
                lt_[r] = eq_[r] = eqb_[r] = 0;
                if (u < size && cs[u] != 0xffffu) {
                    const uint32_t c = cs[u], my = key[u];
                    if (level == 0) {
                        // everybody is in the one group: two carry-out additions per comparison (see k_finish_rows)
                        uint32_t ge_ = 0, le_ = 0;
#pragma unroll 4
                        for (uint32_t v = 0; v < size; v++) {
                            uint32_t x = key[v];
                            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(ge_) : "r"(x), "r"(my));
                            asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(le_) : "r"(my), "r"(x));
                        }
                        lt_[r] = size - ge_; eq_[r] = ge_ + le_ - size;
                        if (eq_[r] > 1) for (uint32_t v = 0; v < u; v++) eqb_[r] += key[v] == my;
                    } else {
                        for (uint32_t v = 0; v < size; v++) {
                            if (cs[v] != c) continue;
                            uint32_t x = key[v];
                            lt_[r] += x < my;
                            uint32_t is = x == my;
                            eq_[r] += is;
                            eqb_[r] += is & (uint32_t)(v < u);
                        }
                    }
                }
            }
            __syncwarp();
            bool live = false;
#pragma unroll
            for (int r = 0; r < FM_EPL; r++) {
                uint32_t u = l + 32 * r;
                if (u < size && cs[u] != 0xffffu) {
                    uint32_t c = cs[u];
                    if (eq_[r] == 1) { emit(start + c + lt_[r], pw[u], 0u); cs[u] = 0xffffu; }
                    else { cs[u] = (uint16_t)(c + lt_[r]); rank[u] = (uint16_t)eqb_[r]; live = true; }
                }
            }
            __syncwarp();
            if (!__any_sync(0xffffffffu, live)) break;
        }
        uint32_t leftover = 0;
        for (uint32_t u = l; u < size; u += 32) {
            if (cs[u] != 0xffffu) {
                uint32_t rr = rank[u];
                emit(start + (uint32_t)cs[u] + rr, pw[u], rr ? NONHEAD : 0u);
                leftover++;
            }
        }
        for (int d = 16; d; d >>= 1) leftover += __shfl_xor_sync(0xffffffffu, leftover, d);
        if (l == 0 && leftover) { atomicAdd(&P.left[lb], leftover); atomicAdd(g_left, (unsigned long long)leftover); }
        __syncwarp();
    }
}

constexpr int FB_TH = 256;
constexpr int FB_MAX = 2048;                     // largest group ranked in shared memory
constexpr int FB_EPT = FB_MAX / FB_TH;

constexpr int FB_SORT_MIN = 256;                 // groups above this size get their level-0 order from a bitonic sort
struct FinBSmem {
    uint64_t srt[FB_MAX];                        // (level-0 key << 32 | member index), sorted
    uint32_t pw[FB_MAX], key[FB_MAX];
    uint16_t cs[FB_MAX], rank[FB_MAX];
    uint32_t end, live;
    uint8_t seq[256];
};

__global__ void __launch_bounds__(FB_TH) k_finish_big(BwtP P, const uint64_t *kv, unsigned long long *g_left, BlockInfo *blocks, uint8_t *lcol,
                                                      const uint64_t *big_list, const uint32_t *big_cnt)
{
    __shared__ FinBSmem S;
    const uint32_t tid = threadIdx.x;
    const uint32_t nbig = *big_cnt;
    for (uint32_t item = blockIdx.x; item < nbig; item += gridDim.x) {
        const uint64_t it = big_list[item];
        const uint32_t lb = (uint32_t)(it >> 32), start = (uint32_t)it;
        const uint32_t n = P.cnt_n[lb];
        const uint64_t *a = kv + (uint64_t)lb * BLK_STRIDE;
        const uint32_t *k30 = P.rk + (uint64_t)lb * BLK_STRIDE;
        const uint8_t *b = P.blk + P.blocks[lb].blk_off;
        uint32_t *sa = P.sa + (uint64_t)lb * BLK_STRIDE;
        uint8_t *L = lcol + (uint64_t)lb * BLK_STRIDE;
        const uint32_t k0 = P.init_k[lb], k32 = P.init_k32[lb];
        __syncthreads();                                  // the previous item is done with the shared arrays
        S.seq[tid] = P.seq[(uint64_t)lb * 256 + tid];
        if (tid == 0) S.end = n;
        __syncthreads();
        // the group ends at the first position whose key differs
        const uint64_t key0 = a[start] >> VAL_BITS;
        for (uint32_t c = start + 1; c < n; c += FB_TH) {
            uint32_t e = c + tid;
            bool diff = e < n && (a[e] >> VAL_BITS) != key0;
            if (diff) atomicMin(&S.end, e);
            if (__syncthreads_or(diff)) break;
        }
        __syncthreads();
        const uint32_t end = S.end, size = end - start;
        auto emit = [&](uint32_t slot_abs, uint32_t pwv, uint32_t flag) {
            uint32_t pos = pwv & VMASK;
            sa[slot_abs] = pos | flag;
            L[slot_abs] = (uint8_t)(pwv >> 24);
            if (pos == 0) blocks[lb].orig_ptr = (int32_t)slot_abs;
        };
        if (size > FB_MAX) {
            // left to the doubling rounds as one unsorted group
            for (uint32_t e = start + tid; e < end; e += FB_TH) {
                uint32_t pos = (uint32_t)a[e] & VMASK;
                emit(e, pos | (uint32_t)S.seq[b[pos ? pos - 1 : n - 1]] << 24, e > start ? NONHEAD : 0u);
            }
            if (tid == 0) { atomicAdd(&P.left[lb], size); atomicAdd(g_left, (unsigned long long)size); }
            continue;
        }
        for (uint32_t e = tid; e < size; e += FB_TH) {
            uint32_t pos = (uint32_t)a[start + e] & VMASK;
            S.pw[e] = pos | (uint32_t)S.seq[b[pos ? pos - 1 : n - 1]] << 24;
            S.cs[e] = 0; S.rank[e] = (uint16_t)e;
        }
        if (tid == 0) S.live = 1;
        __syncthreads();
        uint32_t level0 = 0;
        bool need_levels = true;
        if (size > FB_SORT_MIN) {
            // level 0 by sorting: counting costs size^2 comparisons, and sorted BED with serial ids makes groups of
            // hundreds of rotations whose next symbols all differ
            uint32_t p2 = 512;
            while (p2 < size) p2 <<= 1;
            for (uint32_t e = tid; e < p2; e += FB_TH)
                S.srt[e] = e < size ? (uint64_t)deeper_key(k30, S.pw[e] & VMASK, k0, n) << 32 | e : ~0ull;
            __syncthreads();
            for (uint32_t kk = 2; kk <= p2; kk <<= 1) {
                for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
                    for (uint32_t t = tid; t < p2 / 2; t += FB_TH) {
                        uint32_t i = 2 * t - (t & (j - 1)), ix = i + j;          // the pair (i, i + j), i without bit j
                        uint64_t x = S.srt[i], y = S.srt[ix];
                        bool up = (i & kk) == 0;
                        if ((x > y) == up) { S.srt[i] = y; S.srt[ix] = x; }
                    }
                    __syncthreads();
                }
            }
            // unique keys are final; runs of equal keys go on to the deeper levels as ties
            bool mytie = false;
            for (uint32_t p_ = tid; p_ < size; p_ += FB_TH) {
                uint64_t x = S.srt[p_];
                uint32_t key = (uint32_t)(x >> 32), e = (uint32_t)x;
                bool tie_lo = p_ > 0 && (uint32_t)(S.srt[p_ - 1] >> 32) == key;
                bool tie_hi = p_ + 1 < size && (uint32_t)(S.srt[p_ + 1] >> 32) == key;
                if (!tie_lo && !tie_hi) { emit(start + p_, S.pw[e], 0u); S.cs[e] = 0xffffu; }
                else {
                    uint32_t c = p_;
                    while (c > 0 && (uint32_t)(S.srt[c - 1] >> 32) == key) c--;
                    S.cs[e] = (uint16_t)c; S.rank[e] = (uint16_t)(p_ - c);
                    mytie = true;
                }
            }
            need_levels = __syncthreads_or(mytie);
            level0 = 1;
        }
        for (uint32_t level = level0; need_levels && level < FLEVELS; level++) {
            for (uint32_t e = tid; e < size; e += FB_TH)
                if (S.cs[e] != 0xffffu) S.key[e] = deeper_key(k30, S.pw[e] & VMASK, k0 + level * k32, n);
            __syncthreads();
            uint32_t lt_[FB_EPT], eq_[FB_EPT], eqb_[FB_EPT];
#pragma unroll
            for (int r = 0; r < FB_EPT; r++) {
                uint32_t u = tid + r * FB_TH;
                lt_[r] = eq_[r] = eqb_[r] = 0;
                if (u < size && S.cs[u] != 0xffffu) {
                    uint32_t cs = S.cs[u], my = S.key[u];
                    for (uint32_t v = 0; v < size; v++) {
                        if (S.cs[v] != cs) continue;
                        uint32_t x = S.key[v];
                        lt_[r] += x < my;
                        uint32_t is = x == my;
                        eq_[r] += is;
                        eqb_[r] += is & (uint32_t)(v < u);
                    }
                }
            }
            __syncthreads();
            bool live = false;
#pragma unroll
            for (int r = 0; r < FB_EPT; r++) {
                uint32_t u = tid + r * FB_TH;
                if (u < size && S.cs[u] != 0xffffu) {
                    uint32_t cs = S.cs[u];
                    if (eq_[r] == 1) { emit(start + cs + lt_[r], S.pw[u], 0u); S.cs[u] = 0xffffu; }
                    else { S.cs[u] = (uint16_t)(cs + lt_[r]); S.rank[u] = (uint16_t)eqb_[r]; live = true; }
                }
            }
            if (!__syncthreads_or(live)) break;
        }
        uint32_t leftover = 0;
        for (uint32_t u = tid; u < size; u += FB_TH) {
            if (S.cs[u] != 0xffffu) {
                uint32_t rr = S.rank[u];
                emit(start + (uint32_t)S.cs[u] + rr, S.pw[u], rr ? NONHEAD : 0u);
                leftover++;
            }
        }
        if (leftover) { atomicAdd(&P.left[lb], leftover); atomicAdd(g_left, (unsigned long long)leftover); }
    }
}

// Blocks with leftovers: ranks (SA position of the group head, FINAL on singletons) from the NONHEAD
// flags, flags cleared, so that the doubling rounds can take over at depth k.  Two tile-parallel
// launches: the last group head of every tile, then the ranks with the carry from the earlier tiles.
constexpr int RBT = 1024;                        // threads; 4 positions each = one sort tile

__global__ void __launch_bounds__(RBT) k_rebuild_agg(BwtP P)
{
    __shared__ uint32_t sm[33];
    const uint32_t lb = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (tile == 0 && tid == 0) P.act[lb] = P.left[lb];
    if (!P.left[lb]) return;
    const uint32_t n = P.cnt_n[lb];
    if ((uint64_t)tile * STILE >= n) return;
    const uint32_t *sa = P.sa + (uint64_t)lb * BLK_STRIDE;
    uint32_t lh = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t p = tile * STILE + k * RBT + tid;
        if (p < n && !(sa[p] & NONHEAD)) lh = p + 1;
    }
    uint32_t tot;
    block_excl_max<uint32_t>(lh, sm, &tot);
    if (tid == 0) P.agg[((uint64_t)lb * NT + tile) * 2] = tot;
}

__global__ void __launch_bounds__(RBT) k_rebuild_apply(BwtP P, uint32_t *sa_clean)
{
    __shared__ uint32_t sm[33];
    __shared__ uint32_t s_carry;
    const uint32_t lb = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (!P.left[lb]) return;
    const uint32_t n = P.cnt_n[lb];
    if ((uint64_t)tile * STILE >= n) return;
    {
        uint32_t c = 0;
        const uint32_t *ag = P.agg + (uint64_t)lb * NT * 2;
        for (uint32_t t = tid; t < tile; t += RBT) c = max(c, ag[t * 2]);
        uint32_t tot;
        block_excl_max<uint32_t>(c, sm, &tot);
        if (tid == 0) s_carry = tot;
        __syncthreads();
    }
    const uint32_t *sa = P.sa + (uint64_t)lb * BLK_STRIDE;
    uint32_t *rk = P.rk + (uint64_t)lb * BLK_STRIDE;
    uint32_t *out = sa_clean + (uint64_t)lb * BLK_STRIDE;
    const uint32_t p0 = tile * STILE + tid * 4;
    uint32_t v[5];
    uint32_t lh = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        uint32_t p = p0 + k;
        v[k] = p < n ? sa[p] : 0u;                        // past the end counts as a head
        if (k < 4 && p < n && !(v[k] & NONHEAD)) lh = p + 1;
    }
    uint32_t tot;
    uint32_t hp = max(block_excl_max<uint32_t>(lh, sm, &tot), s_carry);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t p = p0 + k;
        if (p < n) {
            bool head = !(v[k] & NONHEAD);
            if (head) hp = p + 1;
            bool single = head && !(v[k + 1] & NONHEAD);
            uint32_t val = v[k] & VMASK;
            rk[val] = (hp - 1) | (single ? FINAL : 0u);
            out[p] = val;                                 // flags cleared into a second array: neighbouring tiles still read them
        }
    }
}

__global__ void __launch_bounds__(RBT) k_copy_clean(BwtP P, const uint32_t *sa_clean)
{
    const uint32_t lb = blockIdx.y, tile = blockIdx.x;
    if (!P.left[lb]) return;
    const uint32_t n = P.cnt_n[lb];
    uint32_t *sa = P.sa + (uint64_t)lb * BLK_STRIDE;
    const uint32_t *in = sa_clean + (uint64_t)lb * BLK_STRIDE;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t p = tile * STILE + k * RBT + threadIdx.x;
        if (p < n) sa[p] = in[p];
    }
}

constexpr int BT = 512;                 // boundary kernels: 512 threads x 8 items per tile (same 4096-item tiles)
constexpr int BI = STILE / BT;

template <bool INIT>
__global__ void __launch_bounds__(BT, 2) k_bound_apply(BwtP P, uint32_t round, const uint64_t *kv, uint32_t *newrank,
                                                    const uint32_t *act_cur, uint32_t *act_next, unsigned long long *g_act_next)
{
    __shared__ uint32_t sm[33];
    __shared__ uint32_t carry[2];
    uint32_t lb = blockIdx.y, tile = blockIdx.x;
    if (!block_live(P, lb, INIT ? 0 : 1, round, act_cur)) return;
    uint32_t h = depth_of(P, lb, round);
    uint32_t n = P.cnt_n[lb], cnt = INIT ? n : P.cnt_m[lb];
    if ((uint64_t)tile * STILE >= cnt) return;
    // carry-in: max over the aggregates of the earlier tiles
    {
        uint32_t ch = 0, cb = 0;
        const uint32_t *ag = P.agg + (uint64_t)lb * NT * 2;
        for (uint32_t t = threadIdx.x; t < tile; t += BT) { ch = max(ch, ag[t * 2]); cb = max(cb, ag[t * 2 + 1]); }
        uint32_t th, tb;
        block_excl_max<uint32_t>(ch, sm, &th);
        block_excl_max<uint32_t>(cb, sm, &tb);
        if (threadIdx.x == 0) { carry[0] = th; carry[1] = tb; }
        __syncthreads();
    }
    uint32_t p0 = tile * STILE + threadIdx.x * BI;
    uint32_t hm, bm; uint64_t items[BI + 1];
    boundary_flags<INIT, BI>(P, lb, n, cnt, h, kv, p0, hm, bm, items);
    uint32_t own = (1u << BI) - 1;
    uint32_t valid = p0 < cnt ? (cnt - p0 >= BI ? own : (1u << (cnt - p0)) - 1) : 0;
    uint32_t hmo = hm & valid, bmo = bm & valid;
    uint32_t lh = hmo ? p0 + (31 - __clz(hmo)) + 1 : 0;
    uint32_t lbn = bmo ? p0 + (31 - __clz(bmo)) + 1 : 0;
    uint32_t th, tb;
    uint32_t eh = block_excl_max<uint32_t>(lh, sm, &th);
    uint32_t eb = block_excl_max<uint32_t>(lbn, sm, &tb);
    uint32_t hp = max(eh, carry[0]), bp = max(eb, carry[1]);   // position + 1
    uint32_t *sa = P.sa + (uint64_t)lb * BLK_STRIDE;
    uint32_t *rk = P.rk + (uint64_t)lb * BLK_STRIDE;
    uint32_t *nr = newrank + (uint64_t)lb * BLK_STRIDE;
    uint32_t still = 0;
#pragma unroll
    for (int k = 0; k < BI; k++) {
        if (valid & (1u << k)) {
            uint32_t p = p0 + k;
            if (hm & (1u << k)) hp = p + 1;
            if (bm & (1u << k)) bp = p + 1;
            uint32_t key = (uint32_t)(items[k] >> 32), val = (uint32_t)items[k];
            bool single = (bm & (1u << k)) && (bm & (1u << (k + 1)));
            if (!single) still++;
            if (INIT) {
                sa[p] = val;
                rk[val] = (bp - 1) | (single ? FINAL : 0u);
            } else {
                uint32_t q = key + (p - (hp - 1));
                sa[q] = val;
                nr[p] = (key + ((bp - 1) - (hp - 1))) | (single ? FINAL : 0u);
            }
        }
    }
    // count what is still unsorted
    uint32_t tot;
    block_excl_sum<uint32_t>(still, sm, &tot);
    if (threadIdx.x == 0 && tot) { atomicAdd(&act_next[lb], tot); atomicAdd(g_act_next, (unsigned long long)tot); }
}

__global__ void __launch_bounds__(ST) k_rank_update(BwtP P, uint32_t round, const uint64_t *kv, const uint32_t *newrank,
                                                    const uint32_t *act_cur)
{
    uint32_t lb = blockIdx.y, tile = blockIdx.x;
    if (!block_live(P, lb, 1, round, act_cur)) return;
    uint32_t cnt = P.cnt_m[lb];
    if ((uint64_t)tile * STILE >= cnt) return;
    const uint64_t *a = kv + (uint64_t)lb * BLK_STRIDE;
    const uint32_t *nr = newrank + (uint64_t)lb * BLK_STRIDE;
    uint32_t *rk = P.rk + (uint64_t)lb * BLK_STRIDE;
    for (int r = 0; r < SI; r++) {
        uint32_t p = tile * STILE + r * ST + threadIdx.x;
        if (p < cnt) rk[(uint32_t)a[p]] = nr[p];
    }
}

// bucket_min: blocks of at least this many bytes (and two symbols) are sorted by the bucket form; 0 = none
__global__ void k_bwt_setup(BwtP P, uint32_t nb, uint32_t bucket_min)
{
    uint32_t lb = blockIdx.x * blockDim.x + threadIdx.x;
    if (lb >= nb) return;
    uint32_t n = P.blocks[lb].nblock, a = P.blocks[lb].n_in_use;
    P.mode[lb] = bucket_min && n >= bucket_min && a >= 2 ? 1u : 0u;
    if (a < 1) a = 1;
    uint32_t k = 1, k32 = 1;
    if (a >= 2) {
        uint64_t pw = a;
        while (pw * a <= (1ull << KEY_BITS)) { pw *= a; k++; if (pw <= 0xffffffffull) k32 = k; }
    }
    // the room left above a^k is given to a coarse, order-preserving class of the NEXT symbol: key = key_k * f + floor(s_k * f / a)
    uint64_t pwk = 1;
    for (uint32_t i = 0; i < k; i++) pwk *= a;
    uint64_t f = (1ull << KEY_BITS) / pwk;
    if (f > a) f = a;
    if (f < 1) f = 1;
    P.cnt_n[lb] = n; P.init_k[lb] = k; P.init_k32[lb] = k32; P.init_a[lb] = a; P.init_f[lb] = (uint32_t)f;
    P.act[lb] = 0; P.act[nb + lb] = 0; P.left[lb] = 0;
}

// origPtr, tie flag and the BWT last column (bz/compress.c:166-167 reads block[ptr[i]-1])
__global__ void __launch_bounds__(ST) k_bwt_finish(BwtP P, BlockInfo *blocks, uint8_t *lcol)
{
    uint32_t lb = blockIdx.y, tile = blockIdx.x;
    if (!P.left[lb]) return;            // finished by k_finish_rows / k_finish_big
    uint32_t n = P.cnt_n[lb];
    if ((uint64_t)tile * STILE >= n) return;
    const uint32_t *sa = P.sa + (uint64_t)lb * BLK_STRIDE;
    const uint32_t *rk = P.rk + (uint64_t)lb * BLK_STRIDE;
    const uint8_t *b = P.blk + P.blocks[lb].blk_off;
    const uint8_t *sq = P.seq + (uint64_t)lb * 256;
    uint8_t *L = lcol + (uint64_t)lb * BLK_STRIDE;
    bool tie = false;
    for (int r = 0; r < SI; r++) {
        uint32_t k = tile * STILE + r * ST + threadIdx.x;
        if (k < n) {
            uint32_t s = sa[k];
            if (s == 0) blocks[lb].orig_ptr = (int32_t)k;
            if (!(rk[s] & FINAL)) tie = true;
            L[k] = sq[b[s ? s - 1 : n - 1]];
        }
    }
    if (tie) blocks[lb].tie = 1;
}


// ---- periodic blocks: exact origPtr ---------------------------------------------
// When a block is a power of a shorter string, equal rotations tie and the position of
// rotation 0 inside its tie group is whatever libbz2's fallbackSort leaves behind
// (bz/blocksort.c:212-329 with its helper sorts :32-61 and :93-180; SURVEY.md section 7 shows
// libbz2 always ends in fallbackSort for such blocks).  k_fallback_exact replays that algorithm
// on the intact block.  Rare: nblockMAX is prime, so only stream-tail blocks can be periodic.
struct FbState {
    uint32_t *fmap, *ec, *bh;
};
#define FB_SET(z) (st.bh[(z) >> 5] |= (1u << ((z) & 31)))
#define FB_CLR(z) (st.bh[(z) >> 5] &= ~(1u << ((z) & 31)))
#define FB_GET(z) (st.bh[(z) >> 5] & (1u << ((z) & 31)))

__device__ void fb_small_sort(const FbState &st, int lo, int hi)
{
    uint32_t *fmap = st.fmap; const uint32_t *ec = st.ec;
    if (lo == hi) return;
    if (hi - lo > 3) {
        for (int i = hi - 4; i >= lo; i--) {
            uint32_t t = fmap[i], e = ec[t]; int j;
            for (j = i + 4; j <= hi && e > ec[fmap[j]]; j += 4) fmap[j - 4] = fmap[j];
            fmap[j - 4] = t;
        }
    }
    for (int i = hi - 1; i >= lo; i--) {
        uint32_t t = fmap[i], e = ec[t]; int j;
        for (j = i + 1; j <= hi && e > ec[fmap[j]]; j++) fmap[j - 1] = fmap[j];
        fmap[j - 1] = t;
    }
}

__device__ void fb_qsort3(const FbState &st, int lo0, int hi0)
{
    uint32_t *fmap = st.fmap; const uint32_t *ec = st.ec;
    int slo[100], shi[100], sp = 0;
    uint32_t r = 0;
    slo[sp] = lo0; shi[sp] = hi0; sp++;
    while (sp > 0) {
        sp--; int lo = slo[sp], hi = shi[sp];
        if (hi - lo < 10) { fb_small_sort(st, lo, hi); continue; }
        r = (r * 7621 + 1) % 32768;
        uint32_t med;
        uint32_t r3 = r % 3;
        if (r3 == 0) med = ec[fmap[lo]]; else if (r3 == 1) med = ec[fmap[(lo + hi) >> 1]]; else med = ec[fmap[hi]];
        int unLo = lo, ltLo = lo, unHi = hi, gtHi = hi;
        for (;;) {
            while (unLo <= unHi) {
                int32_t d = (int32_t)ec[fmap[unLo]] - (int32_t)med;
                if (d == 0) { uint32_t t = fmap[unLo]; fmap[unLo] = fmap[ltLo]; fmap[ltLo] = t; ltLo++; unLo++; continue; }
                if (d > 0) break;
                unLo++;
            }
            while (unLo <= unHi) {
                int32_t d = (int32_t)ec[fmap[unHi]] - (int32_t)med;
                if (d == 0) { uint32_t t = fmap[unHi]; fmap[unHi] = fmap[gtHi]; fmap[gtHi] = t; gtHi--; unHi--; continue; }
                if (d < 0) break;
                unHi--;
            }
            if (unLo > unHi) break;
            { uint32_t t = fmap[unLo]; fmap[unLo] = fmap[unHi]; fmap[unHi] = t; }
            unLo++; unHi--;
        }
        if (gtHi < ltLo) continue;
        int a = ltLo - lo, b = unLo - ltLo, n = a < b ? a : b;
        for (int p1 = lo, p2 = unLo - n; n > 0; n--, p1++, p2++) { uint32_t t = fmap[p1]; fmap[p1] = fmap[p2]; fmap[p2] = t; }
        a = hi - gtHi; b = gtHi - unHi; int m = a < b ? a : b;
        for (int p1 = unLo, p2 = hi - m + 1; m > 0; m--, p1++, p2++) { uint32_t t = fmap[p1]; fmap[p1] = fmap[p2]; fmap[p2] = t; }
        n = lo + unLo - ltLo - 1;
        m = hi - (gtHi - unHi) + 1;
        if (n - lo > hi - m) { slo[sp] = lo; shi[sp] = n; sp++; slo[sp] = m; shi[sp] = hi; sp++; }
        else                 { slo[sp] = m; shi[sp] = hi; sp++; slo[sp] = lo; shi[sp] = n; sp++; }
    }
}

// a block (mode == which) that is sorted once more: forget the leftovers counted for it
__global__ void k_reset_left(BwtP P, uint32_t nb, uint32_t which, unsigned long long *g_left)
{
    uint32_t lb = blockIdx.x * blockDim.x + threadIdx.x;
    if (lb >= nb || P.mode[lb] != which) return;
    uint32_t l = P.left[lb];
    if (l) { atomicAdd(g_left, (unsigned long long)0 - (unsigned long long)l); P.left[lb] = 0; }
}

// largest set bit <= i / smallest set bit >= i of the bucket-head bit vector (bit 0 and bit n are always set)
__device__ __forceinline__ int fb_head(const uint32_t *bh, int i)
{
    int w = i >> 5;
    uint32_t m = bh[w] & (0xffffffffu >> (31 - (i & 31)));
    while (!m) m = bh[--w];
    return (w << 5) + 31 - __clz((int)m);
}
__device__ __forceinline__ int fb_next(const uint32_t *bh, int i)
{
    int w = i >> 5;
    uint32_t m = bh[w] & (0xffffffffu << (i & 31));
    while (!m) m = bh[++w];
    return (w << 5) + __ffs((int)m) - 1;
}

// One CTA per periodic block.  fallbackSort is a sequence of rounds; inside a round every bucket (maximal range of
// equal rank so far) is sorted by its own fallbackQSort3 call, which starts from r = 0 (bz/blocksort.c:104), reads
// only eclass[] -- fixed during the round -- and permutes only its own range of fmap: the calls of a round are
// independent, so they are replayed one thread per bucket.  A bucket whose keys are all equal is left untouched by
// fallbackQSort3 / fallbackSimpleSort (bz/blocksort.c:32-61, :93-180: only strict comparisons move anything), so only
// the non-uniform buckets are replayed at all; everything else in a round (eclass update, the scan for the buckets,
// the new bucket heads) is position-parallel.
constexpr int FBT = 1024;
__global__ void __launch_bounds__(FBT) k_fallback_exact(BwtP P, BlockInfo *blocks)
{
    const uint32_t lb = blockIdx.x;
    if (!blocks[lb].tie) return;
    const int n = (int)P.cnt_n[lb];
    const uint8_t *blk = P.blk + P.blocks[lb].blk_off;
    FbState st;
    st.fmap = P.sa + (uint64_t)lb * BLK_STRIDE;
    st.ec = P.rk + (uint64_t)lb * BLK_STRIDE;
    st.bh = reinterpret_cast<uint32_t *>(P.kv0 + (uint64_t)lb * BLK_STRIDE);
    uint32_t *scr = reinterpret_cast<uint32_t *>(P.kv1 + (uint64_t)lb * BLK_STRIDE);     // 2 * BLK_STRIDE words
    uint32_t *cnt = scr;                                   // [256][FBT] occurrences of a byte value in a thread's segment
    uint32_t *dlist = scr + 256 * FBT;                     // heads of the non-uniform buckets of the round (at most n / 2)
    uint32_t *dmark = dlist + BLK_STRIDE / 2 + 64;         // [n] round in which a bucket was listed
    __shared__ uint32_t s_tot[256], s_start[256];
    __shared__ uint32_t s_nd, s_notdone, s_refined;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int seg = (n + FBT - 1) / FBT;
    const int i0 = t * seg < n ? t * seg : n, i1 = i0 + seg < n ? i0 + seg : n;
    // ---- initial order: by first byte, positions DESCENDING inside a byte value (bz/blocksort.c:236-241) ----
    for (int c = 0; c < 256; c++) cnt[c * FBT + t] = 0;
    for (int i = i0; i < i1; i++) cnt[blk[i] * FBT + t]++;
    for (int i = i0; i < i1; i++) dmark[i] = 0;
    __syncthreads();
    for (int c = warp; c < 256; c += FBT / 32) {
        uint32_t run = 0;
        for (int k = 0; k < FBT; k += 32) {
            uint32_t v = cnt[c * FBT + k + lane];
            uint32_t inc = warp_incl_sum<uint32_t>(v);
            cnt[c * FBT + k + lane] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_tot[c] = run;
    }
    __syncthreads();
    if (t == 0) { uint32_t acc = 0; for (int c = 0; c < 256; c++) { s_start[c] = acc; acc += s_tot[c]; } }
    const int nbh = n / 32 + 8;
    for (int w = t; w < nbh; w += FBT) st.bh[w] = 0;
    __syncthreads();
    for (int i = i0; i < i1; i++) {
        uint32_t c = blk[i];
        uint32_t r = cnt[c * FBT + t]++;
        st.fmap[s_start[c] + s_tot[c] - 1 - r] = (uint32_t)i;
    }
    if (t < 256) atomicOr(&st.bh[s_start[t] >> 5], 1u << (s_start[t] & 31));
    if (t < 32) atomicOr(&st.bh[(n + 2 * t) >> 5], 1u << ((n + 2 * t) & 31));
    __syncthreads();
    uint32_t round = 0;
    for (long long H = 1;;) {
        round++;
        if (t == 0) { s_nd = 0; s_notdone = 0; s_refined = 0; }
        if (i0 < i1) {
            int j = fb_head(st.bh, i0);
            for (int i = i0; i < i1; i++) {
                if (FB_GET(i)) j = i;
                int k = (int)st.fmap[i] - (int)H; if (k < 0) k += n;
                st.ec[k] = (uint32_t)j;
            }
        }
        __syncthreads();
        if (i0 < i1) {
            int j = fb_head(st.bh, i0);
            uint32_t prevkey = i0 > 0 ? st.ec[st.fmap[i0 - 1]] : 0;
            bool nonhead = false;
            for (int i = i0; i < i1; i++) {
                uint32_t key = st.ec[st.fmap[i]];
                if (FB_GET(i)) j = i;
                else {
                    nonhead = true;
                    if (key != prevkey && atomicExch(&dmark[j], round) != round) dlist[atomicAdd(&s_nd, 1u)] = (uint32_t)j;
                }
                prevkey = key;
            }
            if (nonhead) s_notdone = 1;
        }
        __syncthreads();
        const uint32_t nd = s_nd;
        for (uint32_t b = t; b < nd; b += FBT) {
            int l = (int)dlist[b];
            int r = fb_next(st.bh, l + 1) - 1;
            fb_qsort3(st, l, r);
        }
        __syncthreads();
        if (i0 < i1) {
            uint32_t prevkey = i0 > 0 ? st.ec[st.fmap[i0 - 1]] : 0;
            bool refined = false;
            for (int i = i0; i < i1; i++) {
                uint32_t key = st.ec[st.fmap[i]];
                if (!FB_GET(i) && key != prevkey) { atomicOr(&st.bh[i >> 5], 1u << (i & 31)); refined = true; }
                prevkey = key;
            }
            if (refined) s_refined = 1;
        }
        __syncthreads();
        H *= 2;
        // A round that splits no bucket is a fixed point of the doubling (every bucket's keys are equal), so the remaining
        // rounds of the reference change nothing.  Periodic blocks get here after few rounds.
        const bool stop = H > n || !s_notdone || !s_refined;
        __syncthreads();
        if (stop) break;
    }
    for (int i = i0; i < i1; i++) if (st.fmap[i] == 0) blocks[lb].orig_ptr = i;
}
#undef FB_SET
#undef FB_CLR
#undef FB_GET

// the three forms of a radix pass under the names the per-kernel report uses
static const auto k_sweep_first = k_sweep<false, false>;      // pass 0: records built from the block bytes, no order to keep
static const auto k_sweep_ordered = k_sweep<true, false>;     // stable ranks from ordered shared-memory atomics (checked by the finisher)
static const auto k_sweep_masks = k_sweep<true, true>;        // stable ranks from peer masks (the repeat if that check fails)

int run_bwt(Ctx *ctx, uint64_t b0, uint64_t nb)
{
    if (nb == 0) return S3G_OK;
    size_t slots = (size_t)nb * BLK_STRIDE;
    S3G_TRY(ctx->sa.ensure(slots * 4));
    S3G_TRY(ctx->rk.ensure(slots * 4));
    S3G_TRY(ctx->kv0.ensure(slots * 8));
    S3G_TRY(ctx->kv1.ensure(slots * 8));
    S3G_TRY(ctx->hist.ensure((size_t)nb * NBINS * NT * 4));
    S3G_TRY(ctx->bwt_misc.ensure((size_t)nb * (10 * 4 + NT * 2 * 4) + 64));
    S3G_TRY(ctx->lcol.ensure(slots));
    BwtP P;
    P.blk = ctx->blk_bytes.as<uint8_t>();
    P.seq = ctx->seq_map.as<uint8_t>() + b0 * 256;
    P.blocks = ctx->blocks.as<BlockInfo>() + b0;
    P.sa = ctx->sa.as<uint32_t>(); P.rk = ctx->rk.as<uint32_t>();
    P.kv0 = ctx->kv0.as<uint64_t>(); P.kv1 = ctx->kv1.as<uint64_t>();
    P.hist = ctx->hist.as<uint32_t>();
    uint32_t *misc = ctx->bwt_misc.as<uint32_t>();
    P.g_act = reinterpret_cast<unsigned long long *>(misc); misc += 4;
    uint32_t *big_cnt = misc; misc += 4;                       // [0] listed groups, [1] keys-out-of-order flag, [2] groups passed on to k_finish_big
    P.cnt_n = misc; misc += nb;
    P.cnt_m = misc; misc += nb;
    P.act = misc; misc += 2 * nb;
    P.init_k = misc; misc += nb;
    P.init_a = misc; misc += nb;
    P.left = misc; misc += nb;
    P.init_k32 = misc; misc += nb;
    P.init_f = misc; misc += nb;
    P.mode = misc; misc += nb;
    P.want = 0;
    P.agg = misc;
    dim3 grid(NT, (unsigned)nb);
    double N = 0;                       // rotations in this batch
    for (uint64_t b = 0; b < nb && b0 + b < ctx->h_blocks.size(); b++) N += ctx->h_blocks[b0 + b].nblock;
    const double HS = (double)nb * NT * NBINS * 4 * 2;
    // per context, not per process: the attribute belongs to the device the context is bound to
    if (!ctx->attr_bwt) {
        S3G_CUDA(cudaFuncSetAttribute(k_scatter<MODE_KVX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
        S3G_CUDA(cudaFuncSetAttribute(k_scatter<MODE_KV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
        S3G_CUDA(cudaFuncSetAttribute(k_keys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KeysSmem)));
        S3G_CUDA(cudaFuncSetAttribute(k_sweep_ordered, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SweepSmem)));
        S3G_CUDA(cudaFuncSetAttribute(k_sweep_masks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SweepSmem)));
        S3G_CUDA(cudaFuncSetAttribute(k_sweep_first, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SweepSmem)));
        ctx->attr_bwt = true;
    }
    const uint32_t *no_act = nullptr;
    const uint64_t *no_kv = nullptr;
    uint32_t *no_out = nullptr;
    uint64_t *no_save = nullptr;
    (void)no_act;
    S3G_TRY(ctx->bwt_ghist.ensure((size_t)nb * NPASS * (SWN + 1) * 4));
    uint32_t *ghist = ctx->bwt_ghist.as<uint32_t>();
    uint32_t *tickets = ghist + (size_t)nb * NPASS * SWN;        // [NPASS][nb]
    unsigned long long *h_act = reinterpret_cast<unsigned long long *>(ctx->h_scalars + 32);
    const uint32_t *h_act32 = reinterpret_cast<const uint32_t *>(h_act);
    uint32_t G = SWEEP_G;
    if (const char *e = getenv("S3G_SWEEP_G")) { int v = atoi(e); if (v > 0) G = (uint32_t)v; }
    const unsigned sweep_grid = (unsigned)((nb + G - 1) / G) * G * SW_NT;
    // Which form sorts a block.  The radix form below is the default.  S3G_SORT=bucket gives the blocks of at least
    // BUCKET_MIN bytes to the bucket form (bwt_bucket.cu: 27 HBM bytes per rotation instead of 111, but bound by
    // instruction issue and slower on B200 today -- DESIGN.md section 7); the radix form then sorts the small blocks and
    // whatever the bucket form hands back.  S3G_SORT=safe skips the radix form's first attempt (ordered atomics) and
    // S3G_SORT=broken makes that attempt rank without any order (tests of the ascending-key check and of the second attempt).
    const char *sort_env = getenv("S3G_SORT");
    const bool force_safe = sort_env && !strcmp(sort_env, "safe"), broken = sort_env && !strcmp(sort_env, "broken");
    const bool radix_only = !(sort_env && !strcmp(sort_env, "bucket"));
    constexpr uint32_t BUCKET_MIN = 8192;
    uint64_t n_bucket = 0;
    if (!radix_only)
        for (uint64_t b = 0; b < nb && b0 + b < ctx->h_blocks.size(); b++)
            n_bucket += ctx->h_blocks[b0 + b].nblock >= BUCKET_MIN && ctx->h_blocks[b0 + b].n_in_use >= 2;
    S3G_CUDA(cudaMemsetAsync(P.g_act, 0, 32, ctx->stream));          // leftover totals, list counters, flags
    S3G_LAUNCH(ctx, k_bwt_setup, (unsigned)((nb + 127) / 128), 128, 0, P, (uint32_t)nb, radix_only ? 0u : BUCKET_MIN);

    // ---- radix form on the blocks whose mode is `want`: order by the first k symbols (radix passes on the initial key),
    // then the group finisher; ends with the totals on the host ----
    auto radix_form = [&](uint32_t want) -> int {
        P.want = want;
        // attempt 0 ranks with ordered atomics (k_sweep, SAFE = false); if the finisher finds a key out of order
        // the passes are repeated with peer masks
        for (int attempt = force_safe ? 1 : 0; attempt < 2; attempt++) {
            const bool safe = attempt == 1;
            S3G_CUDA(cudaMemsetAsync(big_cnt, 0, 12, ctx->stream));      // listed groups, out-of-order flag, groups passed on
            S3G_CUDA(cudaMemsetAsync(ghist, 0, (size_t)nb * NPASS * (SWN + 1) * 4, ctx->stream));
            // look-back status words carry a generation tag; the table is cleared only when it is new, was used
            // by the doubling rounds or the bucket form, or the tag wraps
            if (ctx->sweep_cap != ctx->hist.cap || ctx->sweep_gen + NPASS > 255) {
                S3G_CUDA(cudaMemsetAsync(ctx->hist.p, 0, ctx->hist.cap, ctx->stream));
                ctx->sweep_cap = ctx->hist.cap; ctx->sweep_gen = 0;
            }
            S3G_BYTES(ctx, 5 * N);
            S3G_LAUNCH(ctx, k_keys, dim3((NT + KT - 1) / KT, (unsigned)nb), ST, sizeof(KeysSmem), P, ghist);
            S3G_LAUNCH(ctx, k_digit_scan, dim3(NPASS, (unsigned)nb), SWN, 0, ghist);
            uint64_t *src = P.kv0, *dst = P.kv1;
            for (int pass = 0; pass < NPASS; pass++) {
                S3G_BYTES(ctx, 16 * N);
                const int rshift = VAL_BITS + SW_BITS * pass;
                uint32_t *tk = tickets + (size_t)pass * nb;
                if (pass == 0 || (broken && !safe))
                    S3G_LAUNCH(ctx, k_sweep_first, sweep_grid, SWT, sizeof(SweepSmem), P, rshift, src, dst, ghist, pass, tk, (uint32_t)nb, ++ctx->sweep_gen, G);
                else if (!safe)
                    S3G_LAUNCH(ctx, k_sweep_ordered, sweep_grid, SWT, sizeof(SweepSmem), P, rshift, src, dst, ghist, pass, tk, (uint32_t)nb, ++ctx->sweep_gen, G);
                else
                    S3G_LAUNCH(ctx, k_sweep_masks, sweep_grid, SWT, sizeof(SweepSmem), P, rshift, src, dst, ghist, pass, tk, (uint32_t)nb, ++ctx->sweep_gen, G);
                std::swap(src, dst);
            }
            // every group that ends inside a warp's window is finished there, larger ones by one CTA each; SA, last column, origPtr
            S3G_BYTES(ctx, 18 * N);
            S3G_LAUNCH(ctx, k_finish_rows, dim3(FA_NT, (unsigned)nb), FA_WARPS * 32, 0, P, src, P.g_act, ctx->blocks.as<BlockInfo>() + b0,
                       ctx->lcol.as<uint8_t>(), dst, big_cnt, big_cnt + 1);
            uint64_t *list2 = dst + slots / 2;                         // second half of the free sort buffer
            S3G_LAUNCH(ctx, k_finish_mid, SM_COUNT * 4, FM_WARPS * 32, 0, P, src, P.g_act, ctx->blocks.as<BlockInfo>() + b0, ctx->lcol.as<uint8_t>(),
                       dst, big_cnt, list2, big_cnt + 2);
            S3G_LAUNCH(ctx, k_finish_big, SM_COUNT * 4, FB_TH, 0, P, src, P.g_act, ctx->blocks.as<BlockInfo>() + b0, ctx->lcol.as<uint8_t>(), list2,
                       big_cnt + 2);
            S3G_TRY(check_launch("bwt init"));
            S3G_CUDA(cudaMemcpyAsync(h_act, P.g_act, 32, cudaMemcpyDeviceToHost, ctx->stream));
            S3G_CUDA(cudaStreamSynchronize(ctx->stream));
            const uint32_t unsorted = h_act32[5];
            if (getenv("S3G_DEBUG"))
                fprintf(stderr, "[s3g] bwt radix form (mode %u) attempt %d: %llu of %.0f rotations left to the doubling rounds%s\n", want, attempt, *h_act, N,
                        unsorted ? "; keys OUT OF ORDER" : "");
            if (!unsorted) return S3G_OK;
            if (safe) { set_error("bwt: the radix passes did not sort the initial keys"); return S3G_E_CUDA; }
            ctx->sort_retries++;
            S3G_LAUNCH(ctx, k_reset_left, (unsigned)((nb + 127) / 128), 128, 0, P, (uint32_t)nb, want, P.g_act);   // the repeat counts them again
        }
        return S3G_OK;
    };

    bool have_totals = false;
    if (n_bucket) {
        S3G_TRY(run_bucket_sort(ctx, P, b0, nb, P.g_act, big_cnt + 3));
        ctx->bucket_blocks += n_bucket;
    }
    if (n_bucket < nb) { S3G_TRY(radix_form(0)); have_totals = n_bucket == 0; }
    if (!have_totals) {
        S3G_CUDA(cudaMemcpyAsync(h_act, P.g_act, 32, cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        S3G_TRY(check_launch("bwt bucket form"));
        if (getenv("S3G_DEBUG"))
            fprintf(stderr, "[s3g] bwt bucket form: %llu blocks, %llu of %.0f rotations left to the doubling rounds%s\n", (unsigned long long)n_bucket, *h_act, N,
                    h_act32[7] ? "; some blocks handed to the radix form" : "");
        if (h_act32[7]) {
            std::vector<uint32_t> hm(nb);
            S3G_CUDA(cudaMemcpy(hm.data(), P.mode, nb * 4, cudaMemcpyDeviceToHost));
            uint64_t back = 0;
            for (uint32_t m : hm) back += m == 2;
            ctx->bucket_handed_back += back;
            if (getenv("S3G_DEBUG")) fprintf(stderr, "[s3g] bwt bucket form handed %llu of %llu blocks back\n", (unsigned long long)back, (unsigned long long)nb);
            S3G_TRY(radix_form(2));
        }
    }
    if (*h_act == 0) return S3G_OK;
    // (sa with flags) -> ranks; the flag-free order lands in kv1's storage and is copied back
    S3G_LAUNCH(ctx, k_rebuild_agg, grid, RBT, 0, P);
    S3G_LAUNCH(ctx, k_rebuild_apply, grid, RBT, 0, P, reinterpret_cast<uint32_t *>(P.kv1));
    S3G_LAUNCH(ctx, k_copy_clean, grid, RBT, 0, P, reinterpret_cast<const uint32_t *>(P.kv1));
    ctx->sweep_cap = 0;                  // the rounds below reuse the status table as per-tile histograms
    // ---- doubling rounds: block b sorts by depth init_k[b] << round ----
    for (uint32_t round = 0; round < 32; round++) {
        uint32_t *act_cur = P.act + (size_t)(round & 1) * nb;
        uint32_t *act_next = P.act + (size_t)((round + 1) & 1) * nb;
        unsigned long long *g_cur = P.g_act + (round & 1), *g_next = P.g_act + ((round + 1) & 1);
        S3G_CUDA(cudaMemcpyAsync(h_act, g_cur, 8, cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        if (getenv("S3G_DEBUG")) fprintf(stderr, "[s3g] bwt round %u: %llu of %llu rotations unsorted\n", round, *h_act, (unsigned long long)nb * BLK_STRIDE);
        if (*h_act == 0) break;
        S3G_CUDA(cudaMemsetAsync(act_next, 0, nb * 4, ctx->stream));
        S3G_CUDA(cudaMemsetAsync(g_next, 0, 8, ctx->stream));
        uint32_t *newrank = reinterpret_cast<uint32_t *>(P.kv0);
        const double M = (double)*h_act;          // unsorted rotations entering this round
        S3G_BYTES(ctx, 16 * N);
        S3G_LAUNCH(ctx, k_hist<MODE_MM>, grid, ST, 0, P, 32, 1, round, P.cnt_n, no_kv, act_cur, P.kv1);
        S3G_BYTES(ctx, HS);
    S3G_LAUNCH(ctx, k_hist_scan, (unsigned)nb, NBINS, 0, P, 1, round, P.cnt_n, P.cnt_m, act_cur);
        S3G_BYTES(ctx, 8 * N + 8 * M);
        S3G_LAUNCH(ctx, k_scatter<MODE_KVX>, grid, ST, sizeof(ScatterSmem), P, 32, 1, round, P.cnt_n, P.kv1, P.kv0, act_cur);
        S3G_BYTES(ctx, 8 * M);
        S3G_LAUNCH(ctx, k_hist<MODE_KV>, grid, ST, 0, P, 42, 1, round, P.cnt_m, P.kv0, act_cur, no_save);
        S3G_BYTES(ctx, HS);
    S3G_LAUNCH(ctx, k_hist_scan, (unsigned)nb, NBINS, 0, P, 1, round, P.cnt_m, no_out, act_cur);
        S3G_BYTES(ctx, 16 * M);
        S3G_LAUNCH(ctx, k_scatter<MODE_KV>, grid, ST, sizeof(ScatterSmem), P, 42, 1, round, P.cnt_m, P.kv0, P.kv1, act_cur);
        S3G_BYTES(ctx, 8 * M);
        S3G_LAUNCH(ctx, k_bound_agg<false>, grid, ST, 0, P, round, P.kv1, act_cur);
        S3G_BYTES(ctx, 20 * M);
        S3G_LAUNCH(ctx, k_bound_apply<false>, grid, BT, 0, P, round, P.kv1, newrank, act_cur, act_next, g_next);
        S3G_BYTES(ctx, 16 * M);
        S3G_LAUNCH(ctx, k_rank_update, grid, ST, 0, P, round, P.kv1, newrank, act_cur);
        S3G_TRY(check_launch("bwt round"));
    }
    S3G_BYTES(ctx, 10 * N);
    S3G_LAUNCH(ctx, k_bwt_finish, grid, ST, 0, P, ctx->blocks.as<BlockInfo>() + b0, ctx->lcol.as<uint8_t>());
    S3G_LAUNCH(ctx, k_fallback_exact, (unsigned)nb, FBT, 0, P, ctx->blocks.as<BlockInfo>() + b0);
    return check_launch("bwt finish");
}

}  // namespace s3g
