// hash_join.cu -- equi-join of two (value, position) pair lists with a reproducible order.
//
// Replaces hash_join + multimap (/root/reference/src/query.c:652-696, src/multimap.c) and,
// with the sides swapped, nested_loop_join (query.c:585-650).  The reference inserts the
// build side (column_one, the larger input, parse.c:798-813) row by row into an open
// addressing multimap and then walks the probe side in row order, appending for every
// probe row all stored positions of its key in insertion order.  The output is therefore
// probe-major, build-insertion order inside a key -- and must be reproduced exactly.
//
// GPU formulation (shared-memory-partitioned hash build, probe in row order):
//   1. build side: stable LSD radix sort on h = key * 0x9E3779B1 (radix.cu), payload = the
//      build position.  The odd multiplier is a bijection on 32 bits, so equal h <=> equal
//      key: one sort both groups equal keys (insertion order kept inside a group: the sort
//      is stable) and orders the groups by partition id = top bits of h;
//   2. partition boundaries by binary search; one CTA per partition builds an open
//      addressing table {key tag -> group start / single build position, group size} in
//      shared memory and streams it to the partition's slot range in global memory;
//   3. probe rows, in their original order, look their key up in their partition's table:
//      {group start or build position, match count} per probe row, stored coalesced;
//   4. exclusive scan of the match counts in probe-row order = output offsets (radix.cu);
//   5. expand: probe row j copies its group's positions to out1[off[j] ..] and its own
//      position to out2[off[j] ..]; long groups (skewed keys) are spread over a warp.
// Partitions too large for the shared-memory table (heavy skew) build directly in their
// global slots.
#include "adb_common.cuh"

namespace adb {

constexpr int HJ_THREADS = 256;
constexpr uint32_t HJ_SLOTS = 4096;                    // shared-memory table slots
constexpr uint32_t kHashMul = 0x9E3779B1u;

__device__ __forceinline__ uint32_t hj_pid(uint32_t key, uint32_t part_bits) {
    return part_bits ? (key * kHashMul) >> (32 - part_bits) : 0u;
}

// ---- partition boundaries -----------------------------------------------------------------
// The build side arrives sorted on the whole hash, so partition p starts at the first row
// whose id is >= p: one binary search per partition instead of one global atomic per row.
__global__ void hj_bounds_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t part_bits,
                                 uint32_t num_parts, uint32_t *__restrict__ off) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > num_parts) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (hj_pid(keys[mid], part_bits) < p) lo = mid + 1; else hi = mid;
    }
    off[p] = lo;                                  // off[num_parts] == n
}

// ---- table geometry ------------------------------------------------------------------------------
// Partition p gets a power-of-two slot range holding its rows at <= 80 % load (0 slots when it
// is empty); toff = exclusive scan of the capacities, toff[num_parts] = total.  One CTA.
__global__ void __launch_bounds__(1024)
hj_geometry_kernel(const uint32_t *__restrict__ off1, uint32_t num_parts,
                   unsigned long long *__restrict__ toff) {
    __shared__ unsigned long long s_w[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long carry = 0;
    for (uint32_t base = 0; base < num_parts; base += 1024) {
        const uint32_t p = base + threadIdx.x;
        unsigned long long cap = 0;
        if (p < num_parts) {
            const unsigned long long sz = off1[p + 1] - off1[p];
            if (sz) {
                cap = 16;
                while (cap < sz + sz / 4 + 1) cap <<= 1;
            }
        }
        unsigned long long incl = cap;
#pragma unroll
        for (int d = 1; d < kWarp; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(kFull, incl, d);
            if (lane >= (uint32_t)d) incl += y;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        unsigned long long wexcl = 0, tot = 0;
        for (uint32_t w = 0; w < 32; ++w) {
            if (w < warp) wexcl += s_w[w];
            tot += s_w[w];
        }
        if (p < num_parts) toff[p] = carry + wexcl + incl - cap;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) toff[num_parts] = carry;
}

// ---- per-partition tables ---------------------------------------------------------------------
// All keys of a partition share the top `part_bits` bits of h = key * kHashMul, and the
// multiplier is a bijection, so the remaining low bits of h identify the key: a slot stores
// that 32-bit tag (+1, 0 = free), the group's size, and -- for a single-row group, the common
// case -- the build position itself, else the group's start in the sorted build arrays.
//
// r01k history.  (1) One CTA per partition built a shared-memory table and probed it with the
// partition's probe rows (probe side radix-partitioned too), scattering {start, count} to slot
// j: ncu showed 7.9 GB of DRAM traffic at 1.8 TB/s for 100 M probes -- every 8-byte store is a
// 32-byte read-modify-write at a random address (3.1 GB of fill reads, 3.2 GB of writes), and
// shrinking the table from 64 KB to 32 KB per CTA (35 % -> 80 % occupancy) changed nothing:
// the random-access rate of HBM was the limit.  (2) Now the tables are written to global
// memory (16-byte slots, one coalesced burst per partition) and the probe side is walked in
// its ORIGINAL order: one random 16-byte read per probe row, every store coalesced, no probe
// partitioning passes, no scatter.
struct HjSlot {                       // 16 bytes, read with one ld.global.v4
    uint32_t tag;                     // hj_tag(key), 0 = free
    uint32_t val;                     // cnt == 1: build position; else group start
    uint32_t cnt;                     // group size
    uint32_t pad;
};
static_assert(sizeof(HjSlot) == 16, "slot layout");

__device__ __forceinline__ uint32_t hj_tag(uint32_t key, uint32_t part_bits) {
    return (((key * kHashMul) << part_bits) >> part_bits) + 1u;       // part_bits >= 1: never wraps to 0
}
// well-mixed 32 bits of the tag; the caller masks with capacity - 1 (a power of two)
__device__ __forceinline__ uint32_t hj_mix(uint32_t tag) {
    return (uint32_t)(((unsigned long long)tag * 0x9E3779B97F4A7C15ull) >> 29);
}

// One CTA per partition.  Partitions whose slot range is at most HJ_SLOTS (<= 3276 build rows)
// build in shared memory in ONE phase -- every build row claims / finds its key's slot and
// adds 1 to the size, the group leader (first row of the run of equal keys) adds its offset
// in the same atomic -- and stream the finished table out.  Larger partitions (heavy skew)
// run the same steps directly on their global slots.
__global__ void __launch_bounds__(HJ_THREADS)
hj_table_build_kernel(const uint32_t *__restrict__ bkeys /* sorted by hash */,
                      const int32_t *__restrict__ bpos /* build positions in the same order */,
                      const uint32_t *__restrict__ off1, const unsigned long long *__restrict__ toff,
                      uint32_t part_bits, uint4 *__restrict__ table) {
    __shared__ uint32_t s_tag[HJ_SLOTS];
    __shared__ uint32_t s_gc[HJ_SLOTS];                 // {group start within the partition : 16, size : 16}
    const uint32_t p = blockIdx.x;
    const uint32_t b0 = off1[p], b1 = off1[p + 1];
    if (b1 == b0) return;                               // capacity 0: nothing to write
    const unsigned long long t0 = toff[p], cap64 = toff[p + 1] - t0;
    if (cap64 <= HJ_SLOTS) {
        const uint32_t cap = (uint32_t)cap64, mask = cap - 1;
        for (uint32_t s = threadIdx.x; s < cap; s += HJ_THREADS) { s_tag[s] = 0u; s_gc[s] = 0u; }
        __syncthreads();
        for (uint32_t i = b0 + threadIdx.x; i < b1; i += HJ_THREADS) {
            const uint32_t k = bkeys[i];
            const bool leader = i == b0 || bkeys[i - 1] != k;
            const uint32_t tag = hj_tag(k, part_bits);
            uint32_t s = hj_mix(tag) & mask;
            while (true) {
                const uint32_t cur = atomicCAS(&s_tag[s], 0u, tag);
                if (cur == 0u || cur == tag) break;
                s = (s + 1) & mask;
            }
            atomicAdd(&s_gc[s], 1u + (leader ? (i - b0) << 16 : 0u));
        }
        __syncthreads();
        uint4 *__restrict__ out = table + t0;
        for (uint32_t s = threadIdx.x; s < cap; s += HJ_THREADS) {
            const uint32_t w = s_gc[s], c = w & 0xFFFFu, gs = b0 + (w >> 16);
            out[s] = make_uint4(s_tag[s], c == 1 ? (uint32_t)bpos[gs] : gs, c, 0u);
        }
        return;
    }
    uint32_t *T = reinterpret_cast<uint32_t *>(table + t0);          // 4 words per slot
    const unsigned long long mask = cap64 - 1;
    for (unsigned long long s = threadIdx.x; s < cap64; s += HJ_THREADS)
        reinterpret_cast<uint4 *>(T)[s] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (uint32_t i = b0 + threadIdx.x; i < b1; i += HJ_THREADS) {
        const uint32_t k = bkeys[i];
        const bool leader = i == b0 || bkeys[i - 1] != k;
        const uint32_t tag = hj_tag(k, part_bits);
        unsigned long long s = hj_mix(tag) & mask;
        while (true) {
            const uint32_t cur = atomicCAS(&T[4 * s], 0u, tag);
            if (cur == 0u || cur == tag) break;
            s = (s + 1) & mask;
        }
        atomicAdd(&T[4 * s + 2], 1u);
        if (leader) T[4 * s + 1] = i;                    // only the leader writes this word
    }
    __syncthreads();
    for (unsigned long long s = threadIdx.x; s < cap64; s += HJ_THREADS)
        if (T[4 * s + 2] == 1u) T[4 * s + 1] = (uint32_t)bpos[T[4 * s + 1]];
}

// Probe side in its original row order: one random 16-byte slot read per row (linear probing
// almost always ends inside the same 64-byte half line, which is what a miss fetches: ld_gather), results stored coalesced.
__global__ void __launch_bounds__(HJ_THREADS)
hj_probe_kernel(const uint32_t *__restrict__ pkeys, uint32_t n_probe,
                const unsigned long long *__restrict__ toff, uint32_t part_bits,
                const uint4 *__restrict__ table, uint2 *__restrict__ gc_by_j) {
    const uint32_t stride = gridDim.x * HJ_THREADS;
    for (uint32_t j = blockIdx.x * HJ_THREADS + threadIdx.x; j < n_probe; j += stride) {
        const uint32_t k = (uint32_t)ld_stream(reinterpret_cast<const int32_t *>(pkeys) + j);
        const uint32_t p = hj_pid(k, part_bits);
        const unsigned long long t0 = toff[p], cap = toff[p + 1] - t0;
        uint2 r = make_uint2(0u, 0u);
        if (cap) {
            const uint32_t tag = hj_tag(k, part_bits);
            const unsigned long long mask = cap - 1;
            unsigned long long s = hj_mix(tag) & mask;
            while (true) {
                const uint4 sl = ld_gather(table + t0 + s);
                if (sl.x == tag) { r = make_uint2(sl.y, sl.z); break; }
                if (sl.x == 0u) break;
                s = (s + 1) & mask;
            }
        }
        gc_by_j[j] = r;
    }
}

// ---- expand ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HJ_THREADS)
hj_expand_kernel(const uint2 *__restrict__ gc_by_j,
                 const uint32_t *__restrict__ off_by_j, uint32_t n_probe,
                 const int32_t *__restrict__ build_pos_sorted, const int32_t *__restrict__ probe_pos,
                 int32_t *__restrict__ out_build, int32_t *__restrict__ out_probe) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (HJ_THREADS / kWarp);
    const uint32_t warp_id = blockIdx.x * (HJ_THREADS / kWarp) + (threadIdx.x >> 5);
    for (uint32_t j0 = warp_id * kWarp; j0 < n_probe; j0 += warps * kWarp) {
        const uint32_t j = j0 + lane;
        uint32_t cnt = 0, gs = 0, off = 0;
        int32_t pp = 0;
        if (j < n_probe) {
            const uint2 gc = gc_by_j[j];
            cnt = gc.y;
            if (cnt) { gs = gc.x; off = off_by_j[j]; pp = probe_pos[j]; }
        }
        if (cnt == 1) {                                  // gs already is the build position
            out_build[off] = (int32_t)gs;
            out_probe[off] = pp;
        } else if (cnt && cnt <= 8) {
            for (uint32_t r = 0; r < cnt; ++r) {
                out_build[off + r] = ld_gather(build_pos_sorted + gs + r);
                out_probe[off + r] = pp;
            }
        }
        uint32_t longs = __ballot_sync(kFull, cnt > 8);
        while (longs) {
            const int src = __ffs(longs) - 1;
            longs &= longs - 1;
            const uint32_t c = __shfl_sync(kFull, cnt, src), g = __shfl_sync(kFull, gs, src);
            const uint32_t o = __shfl_sync(kFull, off, src);
            const int32_t q = __shfl_sync(kFull, pp, src);
            for (uint32_t r = lane; r < c; r += kWarp) {
                out_build[o + r] = build_pos_sorted[g + r];
                out_probe[o + r] = q;
            }
        }
    }
}

// ---- the same two kernels when the tables live on several contexts ------------------------------
// The owner table travels as a kernel parameter and is copied to shared memory once per CTA.
struct HjOwnersShared {
    const unsigned long long *toff[kMaxPeers];
    const uint4 *table[kMaxPeers];
    const int32_t *bpos[kMaxPeers];
    uint32_t part_bits[kMaxPeers];
};
__device__ __forceinline__ void hj_load_owners(const JoinOwners &o, HjOwnersShared &s) {
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < kMaxPeers; ++k) {
            s.toff[k] = o.toff[k];
            s.table[k] = o.table[k];
            s.bpos[k] = o.bpos[k];
            s.part_bits[k] = o.part_bits[k];
        }
    __syncthreads();
}
__device__ __forceinline__ uint32_t hj_owner(uint32_t key, uint32_t route_bits) {
    return route_bits ? (key * 0x85EBCA6Bu) >> (32 - route_bits) : 0u;
}

__global__ void __launch_bounds__(HJ_THREADS)
hj_probe_sharded_kernel(const uint32_t *__restrict__ pkeys, uint32_t n_probe, const JoinOwners owners,
                        uint2 *__restrict__ gc_by_j) {
    __shared__ HjOwnersShared so;
    hj_load_owners(owners, so);
    const uint32_t stride = gridDim.x * HJ_THREADS;
    for (uint32_t j = blockIdx.x * HJ_THREADS + threadIdx.x; j < n_probe; j += stride) {
        const uint32_t k = (uint32_t)ld_stream(reinterpret_cast<const int32_t *>(pkeys) + j);
        const uint32_t w = hj_owner(k, owners.route_bits);
        const uint32_t part_bits = so.part_bits[w];
        const unsigned long long *__restrict__ toff = so.toff[w];
        const uint32_t p = hj_pid(k, part_bits);
        const unsigned long long t0 = toff[p], cap = toff[p + 1] - t0;      // (peer loads when w is remote)
        uint2 r = make_uint2(0u, 0u);
        if (cap) {
            const uint4 *__restrict__ table = so.table[w];
            const uint32_t tag = hj_tag(k, part_bits);
            const unsigned long long mask = cap - 1;
            unsigned long long s = hj_mix(tag) & mask;
            while (true) {
                const uint4 sl = ld_gather(table + t0 + s);
                if (sl.x == tag) { r = make_uint2(sl.y, sl.z); break; }
                if (sl.x == 0u) break;
                s = (s + 1) & mask;
            }
        }
        gc_by_j[j] = r;
    }
}

__global__ void __launch_bounds__(HJ_THREADS)
hj_expand_sharded_kernel(const uint2 *__restrict__ gc_by_j, const uint32_t *__restrict__ off_by_j,
                         uint32_t n_probe, const uint32_t *__restrict__ pkeys, const JoinOwners owners,
                         const int32_t *__restrict__ probe_pos, int32_t *__restrict__ out_build,
                         int32_t *__restrict__ out_probe) {
    __shared__ HjOwnersShared so;
    hj_load_owners(owners, so);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (HJ_THREADS / kWarp);
    const uint32_t warp_id = blockIdx.x * (HJ_THREADS / kWarp) + (threadIdx.x >> 5);
    for (uint32_t j0 = warp_id * kWarp; j0 < n_probe; j0 += warps * kWarp) {
        const uint32_t j = j0 + lane;
        uint32_t cnt = 0, gs = 0, off = 0, w = 0;
        int32_t pp = 0;
        if (j < n_probe) {
            const uint2 gc = gc_by_j[j];
            cnt = gc.y;
            if (cnt) { gs = gc.x; off = off_by_j[j]; pp = probe_pos[j]; }
            if (cnt > 1) w = hj_owner(pkeys[j], owners.route_bits);      // whose sorted build list holds the group
        }
        if (cnt == 1) {                                  // gs already is the build position
            out_build[off] = (int32_t)gs;
            out_probe[off] = pp;
        } else if (cnt && cnt <= 8) {
            const int32_t *__restrict__ bp = so.bpos[w];
            for (uint32_t r = 0; r < cnt; ++r) {
                out_build[off + r] = bp[gs + r];
                out_probe[off + r] = pp;
            }
        }
        uint32_t longs = __ballot_sync(kFull, cnt > 8);
        while (longs) {
            const int src = __ffs(longs) - 1;
            longs &= longs - 1;
            const uint32_t c = __shfl_sync(kFull, cnt, src), g = __shfl_sync(kFull, gs, src);
            const uint32_t o = __shfl_sync(kFull, off, src), ww = __shfl_sync(kFull, w, src);
            const int32_t q = __shfl_sync(kFull, pp, src);
            const int32_t *__restrict__ bp = so.bpos[ww];
            for (uint32_t r = lane; r < c; r += kWarp) {
                out_build[o + r] = bp[g + r];
                out_probe[o + r] = q;
            }
        }
    }
}

// ---- launchers --------------------------------------------------------------------------------------
int launch_hj_bounds(const uint32_t *keys, uint32_t n, uint32_t part_bits, uint32_t num_parts,
                     uint32_t *off, cudaStream_t s) {
    hj_bounds_kernel<<<(num_parts + 1 + 255) / 256, 256, 0, s>>>(keys, n, part_bits, num_parts, off);
    return 1;
}

int launch_hj_geometry(const uint32_t *off1, uint32_t num_parts, unsigned long long *toff, cudaStream_t s) {
    hj_geometry_kernel<<<1, 1024, 0, s>>>(off1, num_parts, toff);
    return 1;
}

int launch_hj_table_build(const uint32_t *bkeys, const int32_t *bpos, const uint32_t *off1,
                          const unsigned long long *toff, uint32_t num_parts, uint32_t part_bits,
                          uint4 *table, cudaStream_t s) {
    hj_table_build_kernel<<<num_parts, HJ_THREADS, 0, s>>>(bkeys, bpos, off1, toff, part_bits, table);
    return 1;
}

int launch_hj_probe(const uint32_t *pkeys, uint32_t n_probe, const unsigned long long *toff,
                    uint32_t part_bits, const uint4 *table, uint2 *gc_by_j, int sm_count, cudaStream_t s) {
    if (n_probe == 0) return 0;
    uint32_t blocks = (n_probe + HJ_THREADS - 1) / HJ_THREADS;
    if (blocks > (uint32_t)sm_count * 8) blocks = sm_count * 8;
    hj_probe_kernel<<<blocks, HJ_THREADS, 0, s>>>(pkeys, n_probe, toff, part_bits, table, gc_by_j);
    return 1;
}

int launch_hj_expand(const uint2 *gc_by_j, const uint32_t *off_by_j,
                     uint32_t n_probe, const int32_t *build_pos_sorted, const int32_t *probe_pos,
                     int32_t *out_build, int32_t *out_probe, int sm_count, cudaStream_t s) {
    if (n_probe == 0) return 0;
    uint32_t blocks = (n_probe + HJ_THREADS - 1) / HJ_THREADS;
    if (blocks > (uint32_t)sm_count * 8) blocks = sm_count * 8;
    hj_expand_kernel<<<blocks, HJ_THREADS, 0, s>>>(gc_by_j, off_by_j, n_probe,
                                                   build_pos_sorted, probe_pos, out_build, out_probe);
    return 1;
}

int launch_hj_probe_sharded(const uint32_t *pkeys, uint32_t n_probe, const JoinOwners &owners, uint2 *gc_by_j,
                            int sm_count, cudaStream_t s) {
    if (n_probe == 0) return 0;
    uint32_t blocks = (n_probe + HJ_THREADS - 1) / HJ_THREADS;
    if (blocks > (uint32_t)sm_count * 8) blocks = sm_count * 8;
    hj_probe_sharded_kernel<<<blocks, HJ_THREADS, 0, s>>>(pkeys, n_probe, owners, gc_by_j);
    return 1;
}

int launch_hj_expand_sharded(const uint2 *gc_by_j, const uint32_t *off_by_j, uint32_t n_probe,
                             const uint32_t *pkeys, const JoinOwners &owners, const int32_t *probe_pos,
                             int32_t *out_build, int32_t *out_probe, int sm_count, cudaStream_t s) {
    if (n_probe == 0) return 0;
    uint32_t blocks = (n_probe + HJ_THREADS - 1) / HJ_THREADS;
    if (blocks > (uint32_t)sm_count * 8) blocks = sm_count * 8;
    hj_expand_sharded_kernel<<<blocks, HJ_THREADS, 0, s>>>(gc_by_j, off_by_j, n_probe, pkeys, owners, probe_pos,
                                                           out_build, out_probe);
    return 1;
}

// Load this file's kernels now (CUDA loads them lazily, on first launch): a first launch that
// has to load code while another context's kernel spin-waits for this one can stall behind it.
void preload_hash_join() {
    preload_one(reinterpret_cast<const void *>(&hj_probe_sharded_kernel));
    preload_one(reinterpret_cast<const void *>(&hj_expand_sharded_kernel));
    preload_one(reinterpret_cast<const void *>(&hj_expand_kernel));
    preload_one(reinterpret_cast<const void *>(&hj_geometry_kernel));
    preload_one(reinterpret_cast<const void *>(&hj_probe_kernel));
    preload_one(reinterpret_cast<const void *>(&hj_table_build_kernel));
    preload_one(reinterpret_cast<const void *>(&hj_bounds_kernel));
}

}  // namespace adb
