// select_scan.cu -- range select as a single-pass, order-preserving stream compaction.
//
// Replaces select_column_scan (/root/reference/src/query.c:92-137) and select_result
// (query.c:38-86).  The reference appends `position[index++] = i` in one sequential
// loop, so the output must be the ascending list of qualifying rows.  An atomicAdd
// cursor would be unordered; a count/scan/write design reads the column twice.  This
// kernel reads every column byte once and writes every hit once (4N + 4H bytes, the
// algorithmic minimum of SURVEY.md section 8d):
//
//   * one CTA per 4096-row tile, 256 threads x 4 x 16-byte streaming loads, laid out so
//     each warp-wide load is one contiguous 512-byte span (fully coalesced);
//   * the 16 predicate bits of a thread are ranked with ONE packed warp scan (the four
//     per-vector hit counts ride in the four bytes of a word) and a block scan over
//     the eight warp totals;
//   * tile offsets come from a decoupled look-back over 64-bit {epoch, flag, value}
//     status words, so tiles never wait for more than their nearest finished
//     predecessor and nothing is re-read; the epoch tag makes the status array
//     reusable across launches without a memset;
//   * sparse tiles store hits straight from registers; dense tiles stage the compacted
//     tile in shared memory and write it back as fully coalesced rows.
#include "adb_common.cuh"

namespace adb {

constexpr int SEL_THREADS = 256;
constexpr int SEL_WARPS = SEL_THREADS / kWarp;
constexpr int SEL_VEC = 4;                               // int4 loads per thread
constexpr int SEL_ITEMS = SEL_VEC * 4;                   // 16 rows per thread
constexpr int SEL_WARP_ITEMS = kWarp * SEL_ITEMS;        // 512 rows per warp
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;        // 4096 rows per CTA
constexpr uint32_t SEL_DENSE = SEL_TILE / 8;             // >= this many hits: staged write-out

constexpr unsigned long long kFlagAgg = 1ull << 32;      // tile aggregate available
constexpr unsigned long long kFlagPfx = 2ull << 32;      // inclusive prefix available

__device__ __forceinline__ unsigned long long pack_status(uint32_t epoch, unsigned long long flag,
                                                          uint32_t v) {
    return ((unsigned long long)epoch << 34) | flag | v;
}

// Warp 0 walks the predecessors 32 at a time, nearest first, until it meets a tile whose
// inclusive prefix is already known.  Returns this tile's exclusive prefix.
__device__ __forceinline__ uint32_t lookback(const unsigned long long *status, uint32_t epoch,
                                             int tile, uint32_t lane) {
    uint32_t excl = 0;
    int look = tile - 1;
    while (true) {
        const int idx = look - (int)lane;
        unsigned long long w;
        if (idx >= 0) {
            do {
                w = ld_relaxed_u64(status + idx);
            } while ((uint32_t)(w >> 34) != epoch || ((w >> 32) & 3ull) == 0);
        } else {
            w = pack_status(epoch, kFlagPfx, 0);         // before the first tile: prefix 0
        }
        const uint32_t has_pfx = __ballot_sync(kFull, (w & kFlagPfx) != 0);
        const uint32_t v = (uint32_t)w;
        if (has_pfx) {
            const uint32_t first = __ffs(has_pfx) - 1;   // nearest predecessor with a prefix
            return excl + warp_sum(lane <= first ? v : 0u);
        }
        excl += warp_sum(v);
        look -= kWarp;
    }
}

template <bool PAIRS>
__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const SelectArgs a) {
    __shared__ int32_t s_stage[SEL_TILE];
    __shared__ uint32_t s_wtot[SEL_WARPS];
    __shared__ uint32_t s_excl;

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;
    uint32_t n = a.n;
    if (a.d_n) {
        const long long dn = *a.d_n;
        n = dn < (long long)a.n ? (uint32_t)(dn < 0 ? 0 : dn) : a.n;
    }
    const uint32_t tile_start = tile * (uint32_t)SEL_TILE;
    const uint32_t wbase = tile_start + warp * SEL_WARP_ITEMS;
    const int32_t *__restrict__ val = a.val;
    const Range rg = a.range;

    // ---- load + predicate: bit (4*j + k) of `mask` is row wbase + 128*j + 4*lane + k ----
    uint32_t mask = 0;
    if (tile_start + SEL_TILE <= n && (reinterpret_cast<uintptr_t>(val) & 15u) == 0) {
        int4 v[SEL_VEC];
#pragma unroll
        for (int j = 0; j < SEL_VEC; ++j)
            v[j] = ld_stream(reinterpret_cast<const int4 *>(val + wbase + j * 128 + lane * 4));
#pragma unroll
        for (int j = 0; j < SEL_VEC; ++j) {
            mask |= (in_range(v[j].x, rg) ? 1u : 0u) << (4 * j);
            mask |= (in_range(v[j].y, rg) ? 2u : 0u) << (4 * j);
            mask |= (in_range(v[j].z, rg) ? 4u : 0u) << (4 * j);
            mask |= (in_range(v[j].w, rg) ? 8u : 0u) << (4 * j);
        }
    } else if (tile_start < n) {                         // ragged last tile / unaligned column
#pragma unroll
        for (int j = 0; j < SEL_VEC; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t idx = wbase + j * 128 + lane * 4 + k;
                if (idx < n && in_range(ld_stream(val + idx), rg)) mask |= 1u << (4 * j + k);
            }
    }

    // ---- rank: four per-vector counts packed into one word, one warp scan --------------
    const uint32_t c = __popc(mask & 0xFu) | (__popc(mask & 0xF0u) << 8) |
                       (__popc(mask & 0xF00u) << 16) | (__popc(mask & 0xF000u) << 24);
    const uint32_t incl = warp_incl_scan(c, lane);       // bytes stay <= 128: no carries
    const uint32_t tot = __shfl_sync(kFull, incl, 31);
    const uint32_t ex = incl - c;                        // per-vector exclusive rank in warp
    const uint32_t t0 = tot & 0xFF, t1 = (tot >> 8) & 0xFF, t2 = (tot >> 16) & 0xFF;
    const uint32_t wtot = t0 + t1 + t2 + (tot >> 24);
    // rank (inside the warp) of this thread's first hit in vector j
    const uint32_t r0 = (ex & 0xFF);
    const uint32_t r1 = t0 + ((ex >> 8) & 0xFF);
    const uint32_t r2 = t0 + t1 + ((ex >> 16) & 0xFF);
    const uint32_t r3 = t0 + t1 + t2 + (ex >> 24);

    if (lane == 0) s_wtot[warp] = wtot;
    __syncthreads();
    uint32_t wexcl = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < SEL_WARPS; ++w) {
        const uint32_t x = s_wtot[w];
        tile_total += x;
        if ((uint32_t)w < warp) wexcl += x;
    }

    // ---- tile offset: decoupled look-back (warp 0) --------------------------------------
    if (warp == 0) {
        unsigned long long *st = a.status + tile;
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(st, pack_status(a.epoch, kFlagPfx, tile_total));
        } else {
            if (lane == 0) st_relaxed_u64(st, pack_status(a.epoch, kFlagAgg, tile_total));
            excl = lookback(a.status, a.epoch, (int)tile, lane);
            if (lane == 0) st_relaxed_u64(st, pack_status(a.epoch, kFlagPfx, excl + tile_total));
        }
        if (lane == 0) {
            s_excl = excl;
            if (tile == gridDim.x - 1) *a.d_count = (int64_t)(excl + tile_total);
        }
    }
    __syncthreads();
    if (tile_total == 0) return;
    const uint32_t tile_excl = s_excl;

    // ---- write-out ----------------------------------------------------------------------
    if (tile_total >= SEL_DENSE) {
        // dense: compact the tile in shared memory, then stream it out coalesced
        const uint32_t rr[SEL_VEC] = {r0, r1, r2, r3};
#pragma unroll
        for (int j = 0; j < SEL_VEC; ++j) {
            uint32_t r = wexcl + rr[j];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (mask & (1u << (4 * j + k))) {
                    const uint32_t idx = wbase + j * 128 + lane * 4 + k;
                    s_stage[r++] = PAIRS ? a.pos_in[idx] : (int32_t)idx + a.base_pos;
                }
        }
        __syncthreads();
        int32_t *__restrict__ out = a.out + tile_excl;
        for (uint32_t i = threadIdx.x; i < tile_total; i += SEL_THREADS) out[i] = s_stage[i];
    } else {
        // sparse: a handful of hits per warp, store them straight from registers
        const unsigned long long rpack = (unsigned long long)r0 | ((unsigned long long)r1 << 16) |
                                         ((unsigned long long)r2 << 32) |
                                         ((unsigned long long)r3 << 48);
        int32_t *__restrict__ out = a.out + tile_excl + wexcl;
        uint32_t m = mask;
        while (m) {
            const uint32_t b = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t j = b >> 2;
            const uint32_t below = __popc(mask & ((1u << b) - 1u) & (0xFu << (4 * j)));
            const uint32_t r = (uint32_t)(rpack >> (16 * j)) & 0xFFFFu;
            const uint32_t idx = wbase + j * 128 + lane * 4 + (b & 3);
            out[r + below] = PAIRS ? a.pos_in[idx] : (int32_t)idx + a.base_pos;
        }
    }
}

uint32_t select_tile_count(uint32_t n) { return (n + SEL_TILE - 1) / SEL_TILE; }

int launch_select(const SelectArgs &a, cudaStream_t s) {
    const uint32_t tiles = select_tile_count(a.n);
    if (tiles == 0) {
        cudaMemsetAsync(a.d_count, 0, sizeof(int64_t), s);
        return 0;
    }
    if (a.pos_in)
        select_kernel<true><<<tiles, SEL_THREADS, 0, s>>>(a);
    else
        select_kernel<false><<<tiles, SEL_THREADS, 0, s>>>(a);
    return 1;
}

}  // namespace adb
