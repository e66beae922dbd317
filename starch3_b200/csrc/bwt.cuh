// bwt.cuh -- declarations shared by the two forms of the block sort: bwt.cu (radix passes + group finisher + prefix
// doubling) and bwt_bucket.cu (sample-sort buckets finished in shared memory).
#pragma once
#include "common.cuh"

namespace s3g {

constexpr int ST = 256;                 // threads per sort tile
constexpr int SI = 16;                  // items per thread
constexpr int STILE = ST * SI;          // 4096 items per tile
constexpr int NBINS = 1024;             // 10-bit digits
constexpr int NT = (BLK_STRIDE + STILE - 1) / STILE;   // tiles per block slot (220)
constexpr uint32_t FINAL = 0x80000000u;

struct BwtP {
    const uint8_t *blk;        // packed block bytes of the call (block lb of the batch at blocks[lb].blk_off)
    const uint8_t *seq;        // unseqToSeq maps, 256 per block
    const BlockInfo *blocks;   // batch block 0
    uint32_t *sa, *rk;         // [nb][BLK_STRIDE]
    uint64_t *kv0, *kv1;       // [nb][BLK_STRIDE] (key << 32 | val)
    uint32_t *hist;            // [nb][NT][NBINS] (tile-major: every access below is coalesced over digits)
    uint32_t *cnt_n;           // [nb] block sizes
    uint32_t *cnt_m;           // [nb] active items after pass 1
    uint32_t *act;             // [2][nb] unsorted rotations per block (ping-pong by round)
    uint32_t *agg;             // [nb][NT][2] tile aggregates of the boundary scans
    unsigned long long *g_act; // [2] batch totals
    uint32_t *init_k;          // [nb] symbols in the initial key
    uint32_t *init_k32;        // [nb] symbols in the 32-bit per-position key the group finisher compares (<= init_k)
    uint32_t *init_f;          // [nb] classes of the symbol after the k-th that still fit below 2^40 (floor(2^40 / a^k) >= 1)
    uint32_t *init_a;          // [nb] alphabet size
    uint32_t *left;            // [nb] rotations the group finisher left unsorted (blocks that need doubling rounds)
    uint32_t *mode;            // [nb] which form sorts the block: 0 radix passes (bwt.cu), 1 buckets (bwt_bucket.cu), 2 radix passes after the bucket form gave it up
    uint32_t want;             // the radix-form kernels act on the blocks whose mode equals this
};

enum { MODE_INIT = 0, MODE_MM = 1, MODE_KV = 2, MODE_KVX = 3 };   // KVX: key/value pairs saved by the histogram pass, ~0 = not taking part

// Records are 64 bits.  Initial sort: (44-bit symbol key << 20) | rotation start.  Doubling rounds:
// (rank << 32) | rotation start.  A radix pass takes its digit at bit `rshift` of the record.
#ifndef S3G_KEY_BITS
#define S3G_KEY_BITS 44
#endif
constexpr int KEY_BITS = S3G_KEY_BITS;   // initial key (at most 64 - VAL_BITS)
#ifndef S3G_SW_BITS
#define S3G_SW_BITS 9
#endif
#ifndef S3G_SW_OCC
#define S3G_SW_OCC 3
#endif
#ifndef S3G_SWT
#define S3G_SWT 256
#endif
#ifndef S3G_SWI
#define S3G_SWI 16
#endif
constexpr int SW_BITS = S3G_SW_BITS;               // digit width of the onesweep passes
constexpr int SWN = 1 << SW_BITS;        // their bins
constexpr int SW_OCC = S3G_SW_OCC;                // resident CTAs per SM the sweep is compiled for
constexpr int VAL_BITS = 20;             // rotation starts are < 2^20 (BLK_STRIDE)


constexpr int FLEVELS = 12;                      // levels of deeper symbols before a tie is left to the doubling rounds
constexpr uint32_t NONHEAD = 0x80000000u;        // SA flag: same (unsorted) group as the previous position
constexpr uint32_t VMASK = 0x000fffffu;          // rotation start (< 2^20)

// bucket form (bwt_bucket.cu): sorts the blocks whose P.mode is 1; a block it cannot take (a bucket beyond its
// shared-memory capacity: very many equal keys) gets mode 2 and is sorted by the radix form afterwards
int run_bucket_sort(Ctx *ctx, const BwtP &P, uint64_t b0, uint64_t nb, unsigned long long *g_left, uint32_t *d_flags);

}  // namespace s3g
