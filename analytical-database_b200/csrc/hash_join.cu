// hash_join.cu -- equi-join of two (value, position) pair lists with a reproducible order.
//
// Replaces hash_join + multimap (/root/reference/src/query.c:652-696, src/multimap.c) and,
// with the sides swapped, nested_loop_join (query.c:585-650).  The reference inserts the
// build side (column_one, the larger input, parse.c:798-813) row by row into an open
// addressing multimap and then walks the probe side in row order, appending for every
// probe row all stored positions of its key in insertion order.  The output is therefore
// probe-major, build-insertion order inside a key -- and must be reproduced exactly.
//
// GPU formulation (shared-memory-partitioned hash build and probe):
//   1. build side: stable LSD radix sort on h = key * 0x9E3779B1 (radix.cu).  The odd
//      multiplier is a bijection on 32 bits, so equal h <=> equal key: one sort both groups
//      equal keys (insertion order kept inside a group: the sort is stable) and orders
//      the groups by partition id = top bits of h;
//   2. probe side: stable radix partition on the same top bits, payload = probe row j;
//   3. partition boundaries by binary search (both sides are ordered by partition id); one
//      CTA per partition: group leaders insert (key -> group start) into an open
//      addressing table in shared memory (64-bit CAS, no sentinel key needed), group
//      tails store the group end, then every probe row of the partition looks its key up
//      and scatters {group start, match count} -- or, for a single match, {build position,
//      1} -- to slot j;
//   4. exclusive scan of the match counts in probe-row order = output offsets (radix.cu);
//   5. expand: probe row j copies its group's positions to out1[off[j] ..] and its own
//      position to out2[off[j] ..]; long groups (skewed keys) are spread over a warp.
// Partitions too large for the shared-memory table (heavy skew) run the same code on a
// table in global memory.
#include "adb_common.cuh"

namespace adb {

constexpr int HJ_THREADS = 256;
constexpr uint32_t HJ_SLOTS = 4096;                    // shared-memory table slots
constexpr uint32_t HJ_SMEM_TUPLES = 3072;              // largest build partition it holds (75 %)
constexpr uint32_t kHashMul = 0x9E3779B1u;

__device__ __forceinline__ uint32_t hj_pid(uint32_t key, uint32_t part_bits) {
    return part_bits ? (key * kHashMul) >> (32 - part_bits) : 0u;
}

// ---- partition boundaries -----------------------------------------------------------------
// Both inputs arrive ordered by partition id (the build side is sorted on the whole hash, the
// probe side stably partitioned on its top bits), so partition p starts at the first row
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

// ---- per-partition build + probe ----------------------------------------------------------------
struct HjTable {
    unsigned long long *key;      // (1 << 32) | key when occupied, 0 when free
    uint32_t *gs;                 // group start (index into the sorted build arrays)
    uint32_t *ge;                 // group end
    uint32_t slots;               // power of two for shared memory, arbitrary for global
};

__device__ __forceinline__ uint32_t hj_home(uint32_t key, uint32_t slots) {
    // the low hash bits are independent of the (top-bit) partition id
    const uint32_t h = (key * kHashMul) ^ ((key * kHashMul) >> 15);
    return (h * 0x85EBCA6Bu >> 7) % slots;
}

__device__ __forceinline__ uint32_t hj_insert(const HjTable &t, uint32_t key) {
    const unsigned long long want = (1ull << 32) | key;
    uint32_t s = hj_home(key, t.slots);
    while (true) {
        const unsigned long long cur = atomicCAS(&t.key[s], 0ull, want);
        if (cur == 0ull || cur == want) return s;
        s = s + 1 == t.slots ? 0 : s + 1;
    }
}

__device__ __forceinline__ int hj_find(const HjTable &t, uint32_t key) {
    const unsigned long long want = (1ull << 32) | key;
    uint32_t s = hj_home(key, t.slots);
    while (true) {
        const unsigned long long cur = t.key[s];
        if (cur == want) return (int)s;
        if (cur == 0ull) return -1;
        s = s + 1 == t.slots ? 0 : s + 1;
    }
}

// ---- per-partition build + probe, shared-memory table ------------------------------------------
// All keys of a partition share the top `part_bits` bits of h = key * kHashMul, and the
// multiplier is a bijection, so the remaining low bits of h identify the key: the table
// stores that 32-bit tag (+1, 0 = free) instead of the 64-bit {occupied, key} word, and one
// packed word {group start within the partition : 16, group size : 16} per slot.  8 bytes
// per slot instead of 16: 32 KB per CTA, 7 resident CTAs per SM instead of 3 (ncu r01k: the
// 64 KB version ran at 35 % occupancy, long-scoreboard stall 23 per issue, 1.8 TB/s).
// Build is ONE phase: every build row claims / finds its key's slot and adds 1 to the size;
// the group leader (first row of the run of equal keys) adds its offset in the same atomic.
__device__ __forceinline__ uint32_t hj_tag(uint32_t key, uint32_t part_bits) {
    return (((key * kHashMul) << part_bits) >> part_bits) + 1u;       // part_bits >= 1: never wraps to 0
}
__device__ __forceinline__ uint32_t hj_tag_home(uint32_t tag) { return (tag * 0x85EBCA6Bu) >> 20; }   // 12 bits

__global__ void __launch_bounds__(HJ_THREADS)
hj_partition_kernel(const uint32_t *__restrict__ bkeys /* sorted by hash */,
                    const int32_t *__restrict__ bpos /* build positions in the same order */,
                    const uint32_t *__restrict__ off1, const uint32_t *__restrict__ pkeys,
                    const uint32_t *__restrict__ prows /* nullptr: identity */,
                    const uint32_t *__restrict__ off2,
                    const unsigned long long *__restrict__ big_off /* per partition, ~0 = smem */,
                    uint32_t part_bits, uint2 *__restrict__ gc_by_j) {
    __shared__ uint32_t s_tag[HJ_SLOTS];
    __shared__ uint32_t s_gc[HJ_SLOTS];
    static_assert(HJ_SLOTS == 4096, "hj_tag_home yields 12 bits");
    const uint32_t p = blockIdx.x;
    const uint32_t b0 = off1[p], b1 = off1[p + 1], q0 = off2[p], q1 = off2[p + 1];
    if (q0 == q1) return;                                   // nobody probes this partition
    if (b1 == b0) {
        for (uint32_t u = q0 + threadIdx.x; u < q1; u += HJ_THREADS)
            gc_by_j[prows ? prows[u] : u] = make_uint2(0u, 0u);
        return;
    }
    if (big_off && big_off[p] != ~0ull) return;             // hj_partition_big_kernel's job
    for (uint32_t s = threadIdx.x; s < HJ_SLOTS; s += HJ_THREADS) { s_tag[s] = 0u; s_gc[s] = 0u; }
    __syncthreads();
    for (uint32_t i = b0 + threadIdx.x; i < b1; i += HJ_THREADS) {
        const uint32_t k = bkeys[i];
        const bool leader = i == b0 || bkeys[i - 1] != k;
        const uint32_t tag = hj_tag(k, part_bits);
        uint32_t s = hj_tag_home(tag);
        while (true) {
            const uint32_t cur = atomicCAS(&s_tag[s], 0u, tag);
            if (cur == 0u || cur == tag) break;
            s = (s + 1) & (HJ_SLOTS - 1);
        }
        atomicAdd(&s_gc[s], 1u + (leader ? (i - b0) << 16 : 0u));
    }
    __syncthreads();
    for (uint32_t u = q0 + threadIdx.x; u < q1; u += HJ_THREADS) {
        const uint32_t j = prows ? prows[u] : u;
        const uint32_t tag = hj_tag(pkeys[u], part_bits);
        uint32_t s = hj_tag_home(tag);
        uint2 r = make_uint2(0u, 0u);
        while (true) {
            const uint32_t cur = s_tag[s];
            if (cur == tag) {
                // {group start, match count} leaves as ONE 8-byte scattered store per probe
                // row.  A single match (the common case) is resolved here, where the
                // partition's build positions are a contiguous, cache-resident run: the
                // expansion then has no random gather left for it.
                const uint32_t w = s_gc[s], c = w & 0xFFFFu, gs = b0 + (w >> 16);
                r = make_uint2(c == 1 ? (uint32_t)bpos[gs] : gs, c);
                break;
            }
            if (cur == 0u) break;
            s = (s + 1) & (HJ_SLOTS - 1);
        }
        gc_by_j[j] = r;
    }
}

// Partitions too large for the shared-memory table (heavy skew): the same build / probe on a
// 64-bit-keyed table in global memory, two build phases (leaders, then tails).
__global__ void __launch_bounds__(HJ_THREADS)
hj_partition_big_kernel(const uint32_t *__restrict__ bkeys /* sorted by hash */,
                    const int32_t *__restrict__ bpos /* build positions in the same order */,
                    const uint32_t *__restrict__ off1, const uint32_t *__restrict__ pkeys,
                    const uint32_t *__restrict__ prows /* nullptr: identity */,
                    const uint32_t *__restrict__ off2,
                    const unsigned long long *__restrict__ big_off /* per partition, ~0 = smem */,
                    unsigned char *__restrict__ big_mem, uint2 *__restrict__ gc_by_j) {
    const uint32_t p = blockIdx.x;
    if (big_off[p] == ~0ull) return;                        // fits shared memory: done there
    const uint32_t b0 = off1[p], b1 = off1[p + 1], q0 = off2[p], q1 = off2[p + 1];
    if (q0 == q1) return;                                   // nobody probes this partition
    const uint32_t nb = b1 - b0;
    HjTable t;
    t.slots = 2 * nb;
    unsigned char *base = big_mem + big_off[p];
    t.key = reinterpret_cast<unsigned long long *>(base);
    t.gs = reinterpret_cast<uint32_t *>(base + 8ull * t.slots);
    t.ge = t.gs + t.slots;
    for (uint32_t s = threadIdx.x; s < t.slots; s += HJ_THREADS) t.key[s] = 0ull;
    __syncthreads();
    // group leaders claim a slot and record where their group starts
    for (uint32_t i = b0 + threadIdx.x; i < b1; i += HJ_THREADS) {
        const uint32_t k = bkeys[i];
        if (i == b0 || bkeys[i - 1] != k) t.gs[hj_insert(t, k)] = i;
    }
    __syncthreads();
    // group tails record where it ends
    for (uint32_t i = b0 + threadIdx.x; i < b1; i += HJ_THREADS) {
        const uint32_t k = bkeys[i];
        if (i + 1 == b1 || bkeys[i + 1] != k) t.ge[hj_find(t, k)] = i + 1;
    }
    __syncthreads();
    for (uint32_t u = q0 + threadIdx.x; u < q1; u += HJ_THREADS) {
        const uint32_t j = prows ? prows[u] : u;
        const int s = hj_find(t, pkeys[u]);
        // {group start, match count} leaves as ONE 8-byte scattered store per probe row.  A
        // single match (the common case) is resolved here, where the partition's build
        // positions are a contiguous, cache-resident run: the expansion then has no random
        // gather left for it.
        uint2 r = make_uint2(0u, 0u);
        if (s >= 0) {
            const uint32_t gs = t.gs[s], c = t.ge[s] - gs;
            r = make_uint2(c == 1 ? (uint32_t)bpos[gs] : gs, c);
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
                out_build[off + r] = build_pos_sorted[gs + r];
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

// ---- launchers --------------------------------------------------------------------------------------
int launch_hj_bounds(const uint32_t *keys, uint32_t n, uint32_t part_bits, uint32_t num_parts,
                     uint32_t *off, cudaStream_t s) {
    hj_bounds_kernel<<<(num_parts + 1 + 255) / 256, 256, 0, s>>>(keys, n, part_bits, num_parts, off);
    return 1;
}

uint32_t hj_smem_tuples() { return HJ_SMEM_TUPLES; }

int launch_hj_partition(const uint32_t *bkeys, const int32_t *bpos, const uint32_t *off1, const uint32_t *pkeys,
                        const uint32_t *prows, const uint32_t *off2, uint32_t num_parts, uint32_t part_bits,
                        const unsigned long long *big_off, unsigned char *big_mem,
                        uint2 *gc_by_j, cudaStream_t s) {
    hj_partition_kernel<<<num_parts, HJ_THREADS, 0, s>>>(bkeys, bpos, off1, pkeys, prows, off2, big_off,
                                                         part_bits, gc_by_j);
    if (!big_off) return 1;
    hj_partition_big_kernel<<<num_parts, HJ_THREADS, 0, s>>>(bkeys, bpos, off1, pkeys, prows, off2, big_off,
                                                             big_mem, gc_by_j);
    return 2;
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

}  // namespace adb
