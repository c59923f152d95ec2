// radix.cu -- stable radix partition passes over (key, payload) pairs, and a generic
// exclusive scan.  Two users:
//
//   * index build  (replaces quicksort/partition, /root/reference/src/index.c:25-46): four
//     8-bit LSD passes on the sign-flipped key give (values ascending, positions) with ties
//     in ascending row order -- the canonical stable order; identical to the reference's
//     unstable quicksort whenever keys are unique;
//   * hash join    (hash_join.cu): one or two passes on the top bits of a multiplicative
//     hash split both inputs into partitions small enough for a shared-memory table while
//     keeping the original row order inside every partition (stability is what makes the
//     join's output order reproducible).
//
// One pass = histogram (per-CTA, per-bucket counts) -> row scan of the [bucket][cta] matrix
// -> bucket bases -> stable scatter.  Inside the scatter a step covers 256 consecutive rows:
// lanes that share a bucket are ranked with match_any in lane order, warps are ordered
// through a small per-warp count table, so equal digits keep their input order.
#include "adb_common.cuh"

namespace adb {

constexpr int RX_THREADS = 256;
constexpr int RX_WARPS = RX_THREADS / kWarp;
constexpr int RX_BUCKETS = 256;

__device__ __forceinline__ uint32_t rx_digit(uint32_t key, const RadixPass &p) {
    const uint32_t f = p.hash == 1 ? key * 0x9E3779B1u : p.hash == 2 ? key * 0x85EBCA6Bu : key ^ 0x80000000u;
    return (f >> p.shift) & ((1u << p.bits) - 1u);
}

__global__ void __launch_bounds__(RX_THREADS)
rx_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t rows_per_cta, RadixPass p,
               uint32_t *__restrict__ hist /* [bucket][gridDim.x] */) {
    __shared__ uint32_t s_h[RX_BUCKETS];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t begin = blockIdx.x * rows_per_cta;
    const uint32_t end = min(n, begin + rows_per_cta);
    for (uint32_t i = begin + threadIdx.x; i < end; i += RX_THREADS)
        atomicAdd(&s_h[rx_digit(keys[i], p)], 1u);
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s_h[threadIdx.x];
}

// Exclusive scan of every row of a [rows][cols] uint32 matrix in place; row totals out.
__global__ void __launch_bounds__(1024)
rx_row_scan_kernel(uint32_t *__restrict__ mat, uint32_t cols, uint32_t *__restrict__ totals) {
    __shared__ uint32_t s_warp[32];
    uint32_t *row = mat + (size_t)blockIdx.x * cols;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < cols; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t x = i < cols ? row[i] : 0u;
        const uint32_t incl = warp_incl_scan(x, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) s_warp[lane] = warp_incl_scan(s_warp[lane], lane);
        __syncthreads();
        const uint32_t wexcl = warp ? s_warp[warp - 1] : 0u;
        if (i < cols) row[i] = carry + wexcl + incl - x;
        carry += s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// base[b] = sum of totals[0..b)  (<= 256 buckets, one CTA of 256 threads)
__global__ void rx_bucket_base_kernel(const uint32_t *__restrict__ totals, uint32_t *__restrict__ base) {
    __shared__ uint32_t s_warp[RX_WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = totals[threadIdx.x];
    const uint32_t incl = warp_incl_scan(x, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wexcl = 0;
    for (uint32_t w = 0; w < warp; ++w) wexcl += s_warp[w];
    base[threadIdx.x] = wexcl + incl - x;
}

// Stable scatter, one 4096-row tile at a time.  Warp w of the CTA owns the contiguous rows
// tile + w*512 .. +512 (row = ... + i*32 + lane), so ranking its keys round by round with
// match_any gives every key its rank among the warp's equal digits in row order.  Per-digit
// prefixes over the eight warps and over the 256 digits turn that into a slot in the
// tile-sorted order; keys and payloads are parked there in shared memory and written out
// slot by slot, so each digit's run leaves as one contiguous (coalesced) piece instead of
// thirty-two 4-byte scatters per warp.
constexpr int RX_KPT = 16;
constexpr int RX_TILE = RX_THREADS * RX_KPT;             // 4096 rows

// (r01k tried finding the peers through a per-warp shared-memory table -- atomicOr of the
// lane bit, sync, read back, leader clears -- instead of one ballot per digit bit: 20.4 ms
// against 19.6 ms for the 500 M-key sort, so the ballots stayed.)
// REMOTE: bucket d is a destination rank and is written into that rank's receive buffer over
// NVLink (peer_base[d] + key_off / pay_off) instead of one local output array; `base` then
// holds the offset of this rank's piece inside every destination buffer.  The run-contiguous
// write-out is what makes the remote stores full 128-byte transactions.
template <bool REMOTE>
__global__ void __launch_bounds__(RX_THREADS)
rx_scatter_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ pay, uint32_t n,
                  uint32_t rows_per_cta, RadixPass p, const uint32_t *__restrict__ hist,
                  const uint32_t *__restrict__ base, uint32_t *__restrict__ keys_out,
                  uint32_t *__restrict__ pay_out, uint32_t *const *__restrict__ peer_base,
                  unsigned long long key_off, unsigned long long pay_off,
                  const uint32_t *__restrict__ abort_flag) {
    __shared__ uint32_t s_key[RX_TILE];
    __shared__ uint32_t s_pay[RX_TILE];
    __shared__ uint32_t s_wcnt[RX_WARPS][RX_BUCKETS];      // per-warp digit counts -> prefix over warps
    __shared__ uint32_t s_off[RX_BUCKETS];                 // next free global slot of every digit
    __shared__ uint32_t s_tbase[RX_BUCKETS];               // first tile slot of every digit
    __shared__ uint32_t s_gofs[RX_BUCKETS];                // global address = s_gofs[d] + tile slot
    __shared__ uint32_t s_ws[RX_WARPS];
    __shared__ uint32_t *s_peer[REMOTE ? kMaxPeers : 1];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (REMOTE) {
        if (*abort_flag) return;                            // a receive region would overflow
        if (threadIdx.x < (1u << p.bits)) s_peer[threadIdx.x] = peer_base[threadIdx.x];
    }
    s_off[threadIdx.x] = base[threadIdx.x] + hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
    const uint32_t begin = blockIdx.x * rows_per_cta;
    const uint32_t end = min(n, begin + rows_per_cta);
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t tile = begin; tile < end; tile += RX_TILE) {
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) s_wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        uint32_t key[RX_KPT];
        uint32_t dr[RX_KPT];                               // digit | rank-in-warp << 8, ~0 = past the end
        const uint32_t wrow = tile + warp * (kWarp * RX_KPT) + lane;
        uint32_t *cnt = s_wcnt[warp];
        // all sixteen loads are issued before the first rank round needs a key
#pragma unroll
        for (int i = 0; i < RX_KPT; ++i) {
            const uint32_t row = wrow + i * kWarp;
            key[i] = row < end ? ld_stream(reinterpret_cast<const int32_t *>(keys) + row) : 0u;
        }
#pragma unroll
        for (int i = 0; i < RX_KPT; ++i) {
            const uint32_t row = wrow + i * kWarp;
            const bool live = row < end;
            dr[i] = 0xFFFFFFFFu;
            const uint32_t active = __ballot_sync(kFull, live);
            uint32_t d = 0, peers = 0, before = 0;
            if (live) d = rx_digit(key[i], p);
            // lanes holding the same digit, one ballot per digit bit: MATCH.ANY runs on the
            // XU pipe and saturated it at 16 rounds per tile (ncu r01f: xu 177 % of peak)
            peers = active;
            for (int b = 0; b < p.bits; ++b) {
                const bool bit = (d >> b) & 1u;
                const uint32_t vote = __ballot_sync(kFull, bit);
                peers &= bit ? vote : ~vote;
            }
            if (live) before = cnt[d];
            __syncwarp();
            if (live) {
                const uint32_t r = __popc(peers & lt);
                if (r == 0) cnt[d] = before + __popc(peers);
                dr[i] = d | ((before + r) << 8);
            }
            __syncwarp();
        }
        __syncthreads();
        // digit `threadIdx.x`: exclusive prefix over the warps, then over the digits
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) {
            const uint32_t c = s_wcnt[w][threadIdx.x];
            s_wcnt[w][threadIdx.x] = tot;
            tot += c;
        }
        const uint32_t incl = warp_incl_scan(tot, lane);
        if (lane == 31) s_ws[warp] = incl;
        __syncthreads();
        uint32_t wexcl = 0;
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) wexcl += (uint32_t)w < warp ? s_ws[w] : 0u;
        const uint32_t tbase = wexcl + incl - tot;
        s_tbase[threadIdx.x] = tbase;
        s_gofs[threadIdx.x] = s_off[threadIdx.x] - tbase;   // modular: slot >= tbase for this digit
        s_off[threadIdx.x] += tot;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RX_KPT; ++i) {
            if (dr[i] != 0xFFFFFFFFu) {
                const uint32_t d = dr[i] & 0xFFu, r = dr[i] >> 8;
                const uint32_t slot = s_tbase[d] + s_wcnt[warp][d] + r;
                const uint32_t row = wrow + i * kWarp;
                s_key[slot] = key[i];
                s_pay[slot] = pay ? pay[row] : row;
            }
        }
        __syncthreads();
        const uint32_t count = min((uint32_t)RX_TILE, end - tile);
        for (uint32_t slot = threadIdx.x; slot < count; slot += RX_THREADS) {
            const uint32_t k = s_key[slot];
            const uint32_t d = rx_digit(k, p);
            const uint32_t dst = s_gofs[d] + slot;
            if (REMOTE) {
                uint32_t *pb = s_peer[d];
                pb[key_off + dst] = k;
                pb[pay_off + dst] = s_pay[slot];
            } else {
                keys_out[dst] = k;
                pay_out[dst] = s_pay[slot];
            }
        }
        __syncthreads();
    }
}

// One CTA per 4096-row tile, launched in row order: the CTAs resident at any moment work on
// neighbouring tiles, so each of the 256 output streams is appended to by many CTAs at
// neighbouring addresses (DRAM-page and L2 friendly) instead of every CTA opening its own 256
// far-apart streams.  The price is a [256][tiles] histogram (1 KB per 32 KB of input).
RadixGeom radix_geom(uint32_t n, int sm_count) {
    (void)sm_count;
    RadixGeom g{};
    g.rows_per_cta = RX_TILE;
    g.ctas = (n + RX_TILE - 1) / RX_TILE;
    if (g.ctas == 0) g.ctas = 1;
    return g;
}

// scratch: hist = 256 * radix_geom(n).ctas uint32, totals = 256, base = 256
int launch_radix_pass(const uint32_t *keys_in, const uint32_t *pay_in, uint32_t *keys_out,
                      uint32_t *pay_out, uint32_t n, RadixPass p, uint32_t *hist, uint32_t *totals,
                      uint32_t *base, int sm_count, cudaStream_t s) {
    if (n == 0) return 0;
    const RadixGeom g = radix_geom(n, sm_count);
    rx_hist_kernel<<<g.ctas, RX_THREADS, 0, s>>>(keys_in, n, g.rows_per_cta, p, hist);
    rx_row_scan_kernel<<<RX_BUCKETS, 1024, 0, s>>>(hist, g.ctas, totals);
    rx_bucket_base_kernel<<<1, RX_BUCKETS, 0, s>>>(totals, base);
    rx_scatter_kernel<false><<<g.ctas, RX_THREADS, 0, s>>>(keys_in, pay_in, n, g.rows_per_cta, p, hist, base,
                                                           keys_out, pay_out, nullptr, 0, 0, nullptr);
    return 4;
}

int launch_radix_hist(const uint32_t *keys_in, uint32_t n, RadixPass p, uint32_t *hist, uint32_t *totals,
                      int sm_count, cudaStream_t s) {
    if (n == 0) {
        cudaMemsetAsync(totals, 0, sizeof(uint32_t) * RX_BUCKETS, s);
        return 0;
    }
    const RadixGeom g = radix_geom(n, sm_count);
    rx_hist_kernel<<<g.ctas, RX_THREADS, 0, s>>>(keys_in, n, g.rows_per_cta, p, hist);
    rx_row_scan_kernel<<<RX_BUCKETS, 1024, 0, s>>>(hist, g.ctas, totals);
    return 2;
}

int launch_radix_scatter_remote(const uint32_t *keys_in, const uint32_t *pay_in, uint32_t n, RadixPass p,
                                const uint32_t *hist, const uint32_t *base, uint32_t *const *peer_base,
                                unsigned long long key_off, unsigned long long pay_off,
                                const uint32_t *abort_flag, int sm_count, cudaStream_t s) {
    if (n == 0) return 0;
    const RadixGeom g = radix_geom(n, sm_count);
    rx_scatter_kernel<true><<<g.ctas, RX_THREADS, 0, s>>>(keys_in, pay_in, n, g.rows_per_cta, p, hist, base,
                                                          nullptr, nullptr, peer_base, key_off, pay_off, abort_flag);
    return 1;
}

// ---- generic exclusive scan: out[i] = sum(in[0..i)), *total = sum(in[0..n)) -------------------
constexpr int SC_THREADS = 1024;

__global__ void __launch_bounds__(SC_THREADS)
sc_chunk_sum_kernel(const uint32_t *__restrict__ in, uint32_t in_stride, uint32_t n,
                    uint32_t rows_per_cta, unsigned long long *__restrict__ sums) {
    __shared__ unsigned long long s_w[32];
    const uint32_t begin = blockIdx.x * rows_per_cta, end = min(n, begin + rows_per_cta);
    unsigned long long acc = 0;
    for (uint32_t i = begin + threadIdx.x; i < end; i += SC_THREADS) acc += in[(size_t)i * in_stride];
    acc = (unsigned long long)warp_sum_i64((int64_t)acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 32; ++w) t += s_w[w];
        sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(SC_THREADS)
sc_chunk_scan_kernel(const uint32_t *__restrict__ in, uint32_t in_stride, uint32_t n,
                     uint32_t rows_per_cta, const unsigned long long *__restrict__ sums,
                     uint32_t *__restrict__ out, int64_t *__restrict__ total) {
    __shared__ unsigned long long s_w[32];
    __shared__ uint32_t s_warp[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long acc = 0;
    for (uint32_t c = threadIdx.x; c < blockIdx.x; c += SC_THREADS) acc += sums[c];
    acc = (unsigned long long)warp_sum_i64((int64_t)acc);
    if (lane == 0) s_w[warp] = acc;
    __syncthreads();
    unsigned long long base64 = 0;
    for (int w = 0; w < 32; ++w) base64 += s_w[w];
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total = (int64_t)(base64 + sums[blockIdx.x]);
    uint32_t carry = (uint32_t)base64;                      // offsets fit 32 bits (total < 2^31 checked by caller)
    const uint32_t begin = blockIdx.x * rows_per_cta, end = min(n, begin + rows_per_cta);
    for (uint32_t b = begin; b < end; b += SC_THREADS) {
        const uint32_t i = b + threadIdx.x;
        const uint32_t x = i < end ? in[(size_t)i * in_stride] : 0u;
        const uint32_t incl = warp_incl_scan(x, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) s_warp[lane] = warp_incl_scan(s_warp[lane], lane);
        __syncthreads();
        const uint32_t wexcl = warp ? s_warp[warp - 1] : 0u;
        if (i < end) out[i] = carry + wexcl + incl - x;
        carry += s_warp[31];
        __syncthreads();
    }
}

// sums scratch: >= scan_ctas(n) entries of 8 bytes
uint32_t scan_ctas(uint32_t n, int sm_count) {
    uint32_t ctas = (n + 8191) / 8192;
    const uint32_t cap = (uint32_t)sm_count * 2u;
    if (ctas > cap) ctas = cap;
    return ctas ? ctas : 1;
}

int launch_exclusive_scan(const uint32_t *in, uint32_t in_stride, uint32_t *out, uint32_t n,
                          unsigned long long *sums, int64_t *total, int sm_count, cudaStream_t s) {
    if (n == 0) {
        cudaMemsetAsync(total, 0, sizeof(int64_t), s);
        return 0;
    }
    const uint32_t ctas = scan_ctas(n, sm_count);
    uint32_t rows = (n + ctas - 1) / ctas;
    rows = (rows + SC_THREADS - 1) / SC_THREADS * SC_THREADS;
    const uint32_t grid = (n + rows - 1) / rows;
    sc_chunk_sum_kernel<<<grid, SC_THREADS, 0, s>>>(in, in_stride, n, rows, sums);
    sc_chunk_scan_kernel<<<grid, SC_THREADS, 0, s>>>(in, in_stride, n, rows, sums, out, total);
    return 2;
}

// Load this file's kernels now (CUDA loads them lazily, on first launch): a first launch that
// has to load code while another context's kernel spin-waits for this one can stall behind it.
void preload_radix() {
    preload_one(reinterpret_cast<const void *>(&rx_hist_kernel));
    preload_one(reinterpret_cast<const void *>(&rx_row_scan_kernel));
    preload_one(reinterpret_cast<const void *>(&rx_bucket_base_kernel));
    { auto *fp = &rx_scatter_kernel<true>; preload_one(reinterpret_cast<const void *>(fp)); }
    { auto *fp = &rx_scatter_kernel<false>; preload_one(reinterpret_cast<const void *>(fp)); }
    preload_one(reinterpret_cast<const void *>(&sc_chunk_scan_kernel));
    preload_one(reinterpret_cast<const void *>(&sc_chunk_sum_kernel));
}

}  // namespace adb
