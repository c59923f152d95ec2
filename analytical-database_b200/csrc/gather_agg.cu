// gather_agg.cu -- fetch (gather), sum/min/max, element-wise add/sub, synthetic fill.
//
// Replaces the loops of fetch_column (/root/reference/src/query.c:229-231), sum / average /
// min / max (query.c:311-313,333-341,397-402,422-427) and add / sub (query.c:361-363,
// 379-381).  All are HBM-bound streaming kernels: persistent grids (a multiple of the SM
// count), 16-byte accesses, several independent loads in flight per thread, lengths
// optionally read from a device int64 so a select's hit count never visits the host.
#include "adb_common.cuh"

namespace adb {

constexpr int STREAM_THREADS = 256;

__device__ __forceinline__ int64_t resolve_n(int64_t n_max, const int64_t *d_n) {
    if (!d_n) return n_max;
    const int64_t dn = *d_n;
    return dn < 0 ? 0 : (dn < n_max ? dn : n_max);
}

// ---- fetch: out[i] = col[pos[i] - base] ------------------------------------------------
// Position lists from a select are ascending, so neighbouring hits share 32-byte sectors;
// each thread keeps eight independent gathers in flight.
__global__ void __launch_bounds__(STREAM_THREADS)
fetch_kernel(const int32_t *__restrict__ col, const int32_t *__restrict__ pos, int64_t n_max,
             const int64_t *__restrict__ d_n, int32_t base, int32_t *__restrict__ out) {
    const int64_t n = resolve_n(n_max, d_n);
    const int64_t nvec = n >> 2;                         // pos/out come from adb_alloc: 16B aligned
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int4 *pos4 = reinterpret_cast<const int4 *>(pos);
    int4 *out4 = reinterpret_cast<int4 *>(out);
    const bool aligned = ((reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (aligned) {
        for (; i + stride < nvec; i += 2 * stride) {
            const int4 p0 = ld_stream(pos4 + i), p1 = ld_stream(pos4 + i + stride);
            int4 v0, v1;
            v0.x = ld_gather(col + (p0.x - base)); v0.y = ld_gather(col + (p0.y - base));
            v0.z = ld_gather(col + (p0.z - base)); v0.w = ld_gather(col + (p0.w - base));
            v1.x = ld_gather(col + (p1.x - base)); v1.y = ld_gather(col + (p1.y - base));
            v1.z = ld_gather(col + (p1.z - base)); v1.w = ld_gather(col + (p1.w - base));
            out4[i] = v0;
            out4[i + stride] = v1;
        }
        for (; i < nvec; i += stride) {
            const int4 p0 = ld_stream(pos4 + i);
            int4 v0;
            v0.x = ld_gather(col + (p0.x - base)); v0.y = ld_gather(col + (p0.y - base));
            v0.z = ld_gather(col + (p0.z - base)); v0.w = ld_gather(col + (p0.w - base));
            out4[i] = v0;
        }
        for (int64_t t = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
            out[t] = ld_gather(col + (pos[t] - base));
    } else {
        for (; i < n; i += stride) out[i] = ld_gather(col + (pos[i] - base));
    }
}

// ---- fetch over a sharded column: out[i] = shard[pos[i] / shard_rows][pos[i] % shard_rows] --------
// The shard table travels as a kernel parameter and is copied to shared memory once per CTA (a
// dynamically indexed parameter array would be spilled to every thread's local memory).  Remote
// shards are read straight over NVLink (peer loads bypass the local L2).
__global__ void __launch_bounds__(STREAM_THREADS)
fetch_sharded_kernel(const ShardTable t, int n_shards, uint32_t shard_rows, const int32_t *__restrict__ pos,
                     int64_t n_max, const int64_t *__restrict__ d_n, int32_t *__restrict__ out) {
    __shared__ const int32_t *s_ptr[kMaxPeers];
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < kMaxPeers; ++k) s_ptr[k] = t.ptr[k];
    __syncthreads();
    const int64_t n = resolve_n(n_max, d_n);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto one = [&](int32_t p) {
        const uint32_t k = (uint32_t)p / shard_rows;
        return ld_gather(s_ptr[k < (uint32_t)n_shards ? k : 0] + ((uint32_t)p - k * shard_rows));
    };
    for (; i + 3 * stride < n; i += 4 * stride) {          // four independent gathers in flight
        const int32_t p0 = ld_stream(pos + i), p1 = ld_stream(pos + i + stride);
        const int32_t p2 = ld_stream(pos + i + 2 * stride), p3 = ld_stream(pos + i + 3 * stride);
        const int32_t v0 = one(p0), v1 = one(p1), v2 = one(p2), v3 = one(p3);
        out[i] = v0; out[i + stride] = v1; out[i + 2 * stride] = v2; out[i + 3 * stride] = v3;
    }
    for (; i < n; i += stride) out[i] = one(ld_stream(pos + i));
}

// ---- aggregate: {sum (int64), min, max, count} in one pass --------------------------------
__global__ void __launch_bounds__(STREAM_THREADS)
aggregate_kernel(const int32_t *__restrict__ v, int64_t n_max, const int64_t *__restrict__ d_n,
                 adb_agg *__restrict__ out, adb_agg *scratch, unsigned int *ticket) {
    const int64_t n = resolve_n(n_max, d_n);
    AggAcc acc{0, INT32_MAX, INT32_MIN};

    // peel to 16-byte alignment, then vector body, then tail
    const uintptr_t addr = reinterpret_cast<uintptr_t>(v);
    int64_t head = ((16 - (addr & 15u)) & 15u) >> 2;
    if (head > n) head = n;
    const int64_t nvec = (n - head) >> 2;
    const int4 *v4 = reinterpret_cast<const int4 *>(v + head);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = tid;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const int4 a = ld_stream(v4 + i), b = ld_stream(v4 + i + stride);
        const int4 c = ld_stream(v4 + i + 2 * stride), d = ld_stream(v4 + i + 3 * stride);
        acc.add4(a); acc.add4(b); acc.add4(c); acc.add4(d);
    }
    for (; i < nvec; i += stride) acc.add4(ld_stream(v4 + i));
    if (tid < head) acc.add(v[tid]);
    for (int64_t t = head + (nvec << 2) + tid; t < n; t += stride) acc.add(v[t]);
    agg_grid_fold<STREAM_THREADS>(acc, blockIdx.x == 0 && threadIdx.x == 0 ? n : 0, out, scratch, ticket);
}

__global__ void agg_combine_kernel(const adb_agg *__restrict__ parts, int32_t k,
                                   adb_agg *__restrict__ out) {
    AggAcc g{0, INT32_MAX, INT32_MIN};
    int64_t cnt = 0;
    for (int i = threadIdx.x; i < k; i += kWarp) {
        const adb_agg p = parts[i];
        g.sum += p.sum;
        cnt += p.count;
        g.mn = min(g.mn, p.min);
        g.mx = max(g.mx, p.max);
    }
    g.sum = warp_sum_i64(g.sum);
    cnt = warp_sum_i64(cnt);
    g.mn = warp_min_i32(g.mn);
    g.mx = warp_max_i32(g.mx);
    if (threadIdx.x == 0) *out = adb_agg{g.sum, cnt, g.mn, g.mx};
}

// Split a partial into allreduce-ready operands: {sum, count} for ncclSum and
// {max, ~min} for ncclMax (~x = -x-1 reverses the order without overflowing).
__global__ void agg_export_kernel(const adb_agg *__restrict__ a, int64_t *__restrict__ sum_count,
                                  int32_t *__restrict__ max_notmin) {
    sum_count[0] = a->sum;
    sum_count[1] = a->count;
    max_notmin[0] = a->max;
    max_notmin[1] = ~a->min;
}
__global__ void agg_import_kernel(const int64_t *__restrict__ sum_count,
                                  const int32_t *__restrict__ max_notmin, adb_agg *__restrict__ a) {
    *a = adb_agg{sum_count[0], sum_count[1], ~max_notmin[1], max_notmin[0]};
}
int launch_agg_export(const adb_agg *a, int64_t *sum_count, int32_t *max_notmin, cudaStream_t s) {
    agg_export_kernel<<<1, 1, 0, s>>>(a, sum_count, max_notmin);
    return 1;
}
int launch_agg_import(const int64_t *sum_count, const int32_t *max_notmin, adb_agg *a, cudaStream_t s) {
    agg_import_kernel<<<1, 1, 0, s>>>(sum_count, max_notmin, a);
    return 1;
}

// ---- element-wise add / sub (int32, two's-complement wrap) ------------------------------
template <bool SUB>
__global__ void __launch_bounds__(STREAM_THREADS)
ewise_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int64_t n_max,
             const int64_t *__restrict__ d_n, int32_t *__restrict__ out) {
    const int64_t n = resolve_n(n_max, d_n);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                           reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    auto op = [](int32_t x, int32_t y) {
        return SUB ? (int32_t)((uint32_t)x - (uint32_t)y) : (int32_t)((uint32_t)x + (uint32_t)y);
    };
    if (aligned) {
        const int64_t nvec = n >> 2;
        const int4 *a4 = reinterpret_cast<const int4 *>(a), *b4 = reinterpret_cast<const int4 *>(b);
        int4 *o4 = reinterpret_cast<int4 *>(out);
        int64_t i = tid;
        for (; i + stride < nvec; i += 2 * stride) {
            const int4 x0 = ld_stream(a4 + i), y0 = ld_stream(b4 + i);
            const int4 x1 = ld_stream(a4 + i + stride), y1 = ld_stream(b4 + i + stride);
            st_stream(o4 + i, make_int4(op(x0.x, y0.x), op(x0.y, y0.y), op(x0.z, y0.z), op(x0.w, y0.w)));
            st_stream(o4 + i + stride, make_int4(op(x1.x, y1.x), op(x1.y, y1.y), op(x1.z, y1.z), op(x1.w, y1.w)));
        }
        for (; i < nvec; i += stride) {
            const int4 x0 = ld_stream(a4 + i), y0 = ld_stream(b4 + i);
            st_stream(o4 + i, make_int4(op(x0.x, y0.x), op(x0.y, y0.y), op(x0.z, y0.z), op(x0.w, y0.w)));
        }
        for (int64_t t = (nvec << 2) + tid; t < n; t += stride) out[t] = op(a[t], b[t]);
    } else {
        for (int64_t t = tid; t < n; t += stride) out[t] = op(a[t], b[t]);
    }
}

// ---- synthetic columns ------------------------------------------------------------------
__global__ void __launch_bounds__(STREAM_THREADS)
synth_uniform_kernel(int32_t *__restrict__ out, int64_t n, uint64_t seed, uint64_t first_row,
                     int32_t lo, uint32_t span) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t z = mix64(seed, first_row + (uint64_t)i);
        out[i] = (int32_t)((uint32_t)lo + (uint32_t)(((z >> 32) * (uint64_t)span) >> 32));
    }
}

// ---- size_t positions -> int32 (index upload) ----------------------------------------------------
__global__ void __launch_bounds__(STREAM_THREADS)
narrow_u64_kernel(const unsigned long long *__restrict__ src, int64_t n, int32_t *__restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = (int32_t)src[i];
}

// int32 -> size_t (what ColumnIndex.positions holds on the host, cs165_api.h:65-68); with
// iota the source is the row number itself (a clustered index keeps identity positions,
// index.c:89-101,119-135)
__global__ void __launch_bounds__(STREAM_THREADS)
widen_i32_kernel(const int32_t *__restrict__ src, int64_t n, unsigned long long *__restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = src ? (unsigned long long)(uint32_t)src[i] : (unsigned long long)i;
}

__global__ void __launch_bounds__(STREAM_THREADS)
iota_kernel(int32_t *__restrict__ out, int64_t n, int32_t first) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = first + (int32_t)i;
}

// ---- the 100-bin histogram of build_histogram (src/index.c:63-84): bin = (v - min) / bin_size ----
constexpr int HIST_BINS = 128;                           // the reference keeps BIN_NUM = 100
__global__ void __launch_bounds__(STREAM_THREADS)
histogram_kernel(const int32_t *__restrict__ v, int64_t n, int32_t vmin, int32_t bin_size,
                 unsigned long long *__restrict__ counts) {
    __shared__ unsigned int s_c[HIST_BINS];
    if (threadIdx.x < HIST_BINS) s_c[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // the reference computes this in int: (data - min) wraps like its subtraction does
        const int32_t bin = (int32_t)((uint32_t)ld_stream(v + i) - (uint32_t)vmin) / bin_size;
        if (bin >= 0 && bin < HIST_BINS) atomicAdd(&s_c[bin], 1u);
    }
    __syncthreads();
    if (threadIdx.x < HIST_BINS && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
}

// out[i] = ((first_row + i) * mul + add) mod modulus: with gcd(mul, modulus) = 1 a permutation of
// 0 .. modulus-1 (the unique-key column of BASELINE config 3)
__global__ void __launch_bounds__(STREAM_THREADS)
synth_affine_kernel(int32_t *__restrict__ out, int64_t n, uint64_t first_row, uint64_t mul, uint64_t add,
                    uint64_t modulus) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (int32_t)(((first_row + (uint64_t)i) * mul + add) % modulus);
}

// ---- updates and deletes (milestone 5: relational_update / relational_delete,
// project_tests/data_generation_scripts/milestone5.py:123-262; the reference's parser has no
// branch for them, parse.c:876-960) ---------------------------------------------------------------
// update: col[pos[i] - base] = value
__global__ void __launch_bounds__(STREAM_THREADS)
scatter_value_kernel(int32_t *__restrict__ col, int64_t n_rows, const int32_t *__restrict__ pos, int64_t n,
                     int32_t base, int32_t value) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = (int64_t)ld_stream(pos + i) - base;       // rows of other shards are not this call's
        if (r >= 0 && r < n_rows) col[r] = value;
    }
}
// delete, step 1: dead[pos[i] - base] = 1 (dead was zeroed); rows outside [0, n_rows) are ignored
__global__ void __launch_bounds__(STREAM_THREADS)
mark_rows_kernel(uint32_t *__restrict__ dead, int64_t n_rows, const int32_t *__restrict__ pos, int64_t n,
                 int32_t base) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = (int64_t)ld_stream(pos + i) - base;
        if (r >= 0 && r < n_rows) dead[r] = 1u;
    }
}
// delete, step 3: surviving row i moves to i - (dead rows before i); order is kept
__global__ void __launch_bounds__(STREAM_THREADS)
compact_rows_kernel(const int32_t *__restrict__ col, const uint32_t *__restrict__ dead,
                    const uint32_t *__restrict__ dead_before, int64_t n_rows, int32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += stride)
        if (!dead[i]) out[i - dead_before[i]] = ld_stream(col + i);
}

// ---- launchers ------------------------------------------------------------------------------
static int stream_grid(int64_t work_items, int sm_count, int per_sm) {
    const int64_t want = (work_items + STREAM_THREADS - 1) / STREAM_THREADS;
    const int64_t cap = (int64_t)sm_count * per_sm;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

int launch_fetch(const int32_t *col, const int32_t *pos, int64_t n_max, const int64_t *d_n,
                  int32_t base_pos, int32_t *out, int sm_count, cudaStream_t s) {
    if (n_max <= 0) return 0;
    fetch_kernel<<<stream_grid(n_max / 4 + 1, sm_count, 8), STREAM_THREADS, 0, s>>>(
        col, pos, n_max, d_n, base_pos, out);
    return 1;
}

int launch_fetch_sharded(const ShardTable &t, int n_shards, uint32_t shard_rows, const int32_t *pos,
                         int64_t n_max, const int64_t *d_n, int32_t *out, int sm_count, cudaStream_t s) {
    if (n_max <= 0) return 0;
    fetch_sharded_kernel<<<stream_grid(n_max / 4 + 1, sm_count, 8), STREAM_THREADS, 0, s>>>(
        t, n_shards, shard_rows, pos, n_max, d_n, out);
    return 1;
}

int launch_aggregate(const int32_t *v, int64_t n_max, const int64_t *d_n, adb_agg *out,
                      adb_agg *scratch, unsigned int *ticket, int sm_count, cudaStream_t s) {
    int grid = stream_grid(n_max / 16 + 1, sm_count, 8);
    if (grid > kAggMaxBlocks) grid = kAggMaxBlocks;
    aggregate_kernel<<<grid, STREAM_THREADS, 0, s>>>(v, n_max < 0 ? 0 : n_max, d_n, out, scratch, ticket);
    return 1;
}

int launch_agg_combine(const adb_agg *parts, int32_t k, adb_agg *out, cudaStream_t s) {
    agg_combine_kernel<<<1, kWarp, 0, s>>>(parts, k, out);
    return 1;
}

int launch_ewise(const int32_t *a, const int32_t *b, int64_t n_max, const int64_t *d_n,
                  int32_t *out, bool subtract, int sm_count, cudaStream_t s) {
    if (n_max <= 0) return 0;
    const int grid = stream_grid(n_max / 8 + 1, sm_count, 8);
    if (subtract)
        ewise_kernel<true><<<grid, STREAM_THREADS, 0, s>>>(a, b, n_max, d_n, out);
    else
        ewise_kernel<false><<<grid, STREAM_THREADS, 0, s>>>(a, b, n_max, d_n, out);
    return 1;
}

int launch_synth_affine(int32_t *out, int64_t n, uint64_t first_row, uint64_t mul, uint64_t add, uint64_t modulus,
                        int sm_count, cudaStream_t s) {
    if (n <= 0) return 0;
    synth_affine_kernel<<<stream_grid(n / 4 + 1, sm_count, 16), STREAM_THREADS, 0, s>>>(out, n, first_row, mul, add, modulus);
    return 1;
}

int launch_scatter_value(int32_t *col, int64_t n_rows, const int32_t *pos, int64_t n, int32_t base, int32_t value,
                         int sm_count, cudaStream_t s) {
    if (n <= 0 || n_rows <= 0) return 0;
    scatter_value_kernel<<<stream_grid(n / 4 + 1, sm_count, 8), STREAM_THREADS, 0, s>>>(col, n_rows, pos, n, base, value);
    return 1;
}
int launch_mark_rows(uint32_t *dead, int64_t n_rows, const int32_t *pos, int64_t n, int32_t base, int sm_count,
                     cudaStream_t s) {
    if (n <= 0) return 0;
    mark_rows_kernel<<<stream_grid(n / 4 + 1, sm_count, 8), STREAM_THREADS, 0, s>>>(dead, n_rows, pos, n, base);
    return 1;
}
int launch_compact_rows(const int32_t *col, const uint32_t *dead, const uint32_t *dead_before, int64_t n_rows,
                        int32_t *out, int sm_count, cudaStream_t s) {
    if (n_rows <= 0) return 0;
    compact_rows_kernel<<<stream_grid(n_rows / 4 + 1, sm_count, 8), STREAM_THREADS, 0, s>>>(col, dead, dead_before,
                                                                                          n_rows, out);
    return 1;
}

int launch_narrow_u64(const unsigned long long *src, int64_t n, int32_t *dst, int sm_count, cudaStream_t s) {
    if (n <= 0) return 0;
    narrow_u64_kernel<<<stream_grid(n / 2 + 1, sm_count, 8), STREAM_THREADS, 0, s>>>(src, n, dst);
    return 1;
}

int launch_widen_i32(const int32_t *src, int64_t n, unsigned long long *dst, int sm_count, cudaStream_t s) {
    if (n <= 0) return 0;
    widen_i32_kernel<<<stream_grid(n / 2 + 1, sm_count, 8), STREAM_THREADS, 0, s>>>(src, n, dst);
    return 1;
}

int launch_iota(int32_t *out, int64_t n, int32_t first, int sm_count, cudaStream_t s) {
    if (n <= 0) return 0;
    iota_kernel<<<stream_grid(n / 4 + 1, sm_count, 8), STREAM_THREADS, 0, s>>>(out, n, first);
    return 1;
}

int launch_histogram(const int32_t *v, int64_t n, int32_t vmin, int32_t bin_size, unsigned long long *counts,
                     int sm_count, cudaStream_t s) {
    cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * HIST_BINS, s);
    if (n <= 0) return 0;
    histogram_kernel<<<stream_grid(n / 8 + 1, sm_count, 4), STREAM_THREADS, 0, s>>>(v, n, vmin, bin_size, counts);
    return 1;
}

int launch_synth_uniform(int32_t *out, int64_t n, uint64_t seed, uint64_t first_row, int32_t lo,
                          uint32_t span, int sm_count, cudaStream_t s) {
    if (n <= 0) return 0;
    synth_uniform_kernel<<<stream_grid(n / 4 + 1, sm_count, 16), STREAM_THREADS, 0, s>>>(
        out, n, seed, first_row, lo, span);
    return 1;
}

// Load this file's kernels now (CUDA loads them lazily, on first launch): a first launch that
// has to load code while another context's kernel spin-waits for this one can stall behind it.
void preload_gather_agg() {
    preload_one(reinterpret_cast<const void *>(&scatter_value_kernel));
    preload_one(reinterpret_cast<const void *>(&mark_rows_kernel));
    preload_one(reinterpret_cast<const void *>(&compact_rows_kernel));
    preload_one(reinterpret_cast<const void *>(&aggregate_kernel));
    { auto *fp = &ewise_kernel<true>; preload_one(reinterpret_cast<const void *>(fp)); }
    { auto *fp = &ewise_kernel<false>; preload_one(reinterpret_cast<const void *>(fp)); }
    preload_one(reinterpret_cast<const void *>(&fetch_kernel));
    preload_one(reinterpret_cast<const void *>(&fetch_sharded_kernel));
    preload_one(reinterpret_cast<const void *>(&histogram_kernel));
    preload_one(reinterpret_cast<const void *>(&iota_kernel));
    preload_one(reinterpret_cast<const void *>(&narrow_u64_kernel));
    preload_one(reinterpret_cast<const void *>(&synth_affine_kernel));
    preload_one(reinterpret_cast<const void *>(&synth_uniform_kernel));
    preload_one(reinterpret_cast<const void *>(&widen_i32_kernel));
    preload_one(reinterpret_cast<const void *>(&agg_combine_kernel));
    preload_one(reinterpret_cast<const void *>(&agg_export_kernel));
    preload_one(reinterpret_cast<const void *>(&agg_import_kernel));
}

}  // namespace adb
