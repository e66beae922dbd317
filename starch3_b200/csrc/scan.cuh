// scan.cuh -- device-wide exclusive scan in three launches (tile reduce, aggregate
// scan, tile apply).  The functor F supplies
//     typedef T;  static T identity();  static T op(T a, T b);   (associative, order-preserving)
//     T load(uint64_t i) const;                                   value of item i
//     void store(uint64_t i, T excl, T val) const;                called once per item in the apply pass
// Items are assigned blocked: thread t of a tile owns SCAN_ITEMS consecutive items.
#pragma once
#include "common.cuh"

namespace s3g {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

#ifdef __CUDACC__
// shuffle of any trivially copyable T whose size is a multiple of 4 bytes
template <class T> __device__ __forceinline__ T shfl_up_any(T v, int d)
{
    static_assert(sizeof(T) % 4 == 0, "shuffled word by word");
    constexpr int W = sizeof(T) / 4;
    union U { T t; uint32_t w[W]; __device__ U() {} } a, b;
    a.t = v;
#pragma unroll
    for (int i = 0; i < W; i++) b.w[i] = __shfl_up_sync(0xffffffffu, a.w[i], d);
    return b.t;
}

// Exclusive scan of one partial per thread, any T / associative op (operands stay in thread order):
// shuffles inside the warp, the eight warp totals through shared memory.
template <class F, int NTH = SCAN_THREADS> __device__ __forceinline__ typename F::T block_scan_partials(typename F::T part, typename F::T *sm,
                                                                                                    typename F::T *block_total)
{
    typedef typename F::T T;
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = part;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_any(inc, d);
        if (l >= d) inc = F::op(o, inc);
    }
    if (l == 31) sm[w] = inc;
    __syncthreads();
    T pre = F::identity(), tot = F::identity();
#pragma unroll
    for (int ww = 0; ww < NTH / 32; ww++) {
        T x = sm[ww];
        if (ww < w) pre = F::op(pre, x);
        tot = F::op(tot, x);
    }
    T up = shfl_up_any(inc, 1);
    T excl = l == 0 ? pre : F::op(pre, up);
    *block_total = tot;
    __syncthreads();
    return excl;
}

template <class F> __global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(F f, uint64_t n, typename F::T *agg)
{
    typedef typename F::T T;
    __shared__ T sm[SCAN_THREADS];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    T acc = F::identity();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) acc = F::op(acc, f.load(base + k));
    T tot;
    block_scan_partials<F>(acc, sm, &tot);
    if (threadIdx.x == 0) agg[blockIdx.x] = tot;
}

// single CTA: exclusive scan of the tile aggregates in place; total -> *total
constexpr int AGG_THREADS = 1024;
template <class F> __global__ void __launch_bounds__(AGG_THREADS) k_scan_agg(typename F::T *agg, uint64_t ntiles,
                                                                             typename F::T *total)
{
    typedef typename F::T T;
    __shared__ T sm[AGG_THREADS / 32];
    T carry = F::identity();
    for (uint64_t base = 0; base < ntiles; base += (uint64_t)AGG_THREADS * SCAN_ITEMS) {
        uint64_t i0 = base + (uint64_t)threadIdx.x * SCAN_ITEMS;
        T v[SCAN_ITEMS];
        T acc = F::identity();
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = i0 + k < ntiles ? agg[i0 + k] : F::identity();
            acc = F::op(acc, v[k]);
        }
        T tot;
        T excl = block_scan_partials<F, AGG_THREADS>(acc, sm, &tot);
        T run = F::op(carry, excl);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (i0 + k < ntiles) agg[i0 + k] = run;
            run = F::op(run, v[k]);
        }
        carry = F::op(carry, tot);
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

template <class F> __global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(F f, uint64_t n, const typename F::T *agg)
{
    typedef typename F::T T;
    __shared__ T sm[SCAN_THREADS];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = F::identity();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? f.load(base + k) : F::identity();
        acc = F::op(acc, v[k]);
    }
    T tot;
    T excl = block_scan_partials<F>(acc, sm, &tot);
    T run = F::op(agg[blockIdx.x], excl);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) f.store(base + k, run, v[k]);
        run = F::op(run, v[k]);
    }
}

// agg needs ceil(n / SCAN_TILE) + 1 slots of F::T; total (device) may be null
template <class F> int device_scan(Ctx *ctx, F f, uint64_t n, typename F::T *agg, typename F::T *d_total)
{
    uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ntiles == 0) ntiles = 1;
    if (ntiles > 0x7fffffffull) { set_error("scan too large"); return S3G_E_LIMIT; }
    S3G_LAUNCH(ctx, k_scan_reduce<F>, (unsigned)ntiles, SCAN_THREADS, 0, f, n, agg);
    S3G_LAUNCH(ctx, k_scan_agg<F>, 1, AGG_THREADS, 0, agg, ntiles, d_total);
    S3G_LAUNCH(ctx, k_scan_apply<F>, (unsigned)ntiles, SCAN_THREADS, 0, f, n, agg);
    return check_launch("device_scan");
}
#endif

}  // namespace s3g
