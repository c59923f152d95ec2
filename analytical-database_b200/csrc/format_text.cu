// format_text.cu -- result text for print: int32 values -> "%d" joined by '\n', on the device.
//
// Replaces the sprintf loop of print for INT results (/root/reference/src/query.c:262-269:
// `sprintf("%d")` per tuple, "\n" between tuples, none after the last) -- SURVEY.md 8f
// rank 2: once a fetch of 10^8 values takes milliseconds, formatting them one sprintf at a
// time on the host is what a `print` waits for.  The values never leave HBM as integers;
// the finished text is downloaded instead (adb_download stages it through pinned lanes).
//
//   fmt_len_kernel    bytes needed by each 1024-value block (digits + sign + separator);
//   exclusive scan    of the block sizes (radix.cu) -> first byte of every block;
//   fmt_emit_kernel   every thread formats 4 consecutive values into a shared-memory tile at
//                     its scanned offset; the tile leaves as 16-byte rows.
#include "adb_common.cuh"

namespace adb {

constexpr int FMT_THREADS = 256;
constexpr int FMT_VPT = 4;                                   // values per thread
constexpr int FMT_BLOCK_VALUES = FMT_THREADS * FMT_VPT;      // 1024
constexpr int FMT_MAX_LEN = 12;                              // "-2147483648" + '\n'

__device__ __forceinline__ uint32_t fmt_digits(uint32_t u) {
    return 1u + (u >= 10u) + (u >= 100u) + (u >= 1000u) + (u >= 10000u) + (u >= 100000u) +
           (u >= 1000000u) + (u >= 10000000u) + (u >= 100000000u) + (u >= 1000000000u);
}
// length of "%d" of v plus one separator byte
__device__ __forceinline__ uint32_t fmt_len(int32_t v) {
    const uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    return fmt_digits(u) + (v < 0 ? 2u : 1u);
}

__device__ __forceinline__ void load4(const int32_t *__restrict__ val, int64_t i, int64_t n, int32_t (&v)[FMT_VPT],
                                      uint32_t (&len)[FMT_VPT]) {
    if (i + FMT_VPT <= n && (reinterpret_cast<uintptr_t>(val) & 15u) == 0) {
        const int4 x = ld_stream(reinterpret_cast<const int4 *>(val + i));
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    } else {
#pragma unroll
        for (int k = 0; k < FMT_VPT; ++k) v[k] = i + k < n ? ld_stream(val + i + k) : 0;
    }
#pragma unroll
    for (int k = 0; k < FMT_VPT; ++k) len[k] = i + k < n ? fmt_len(v[k]) : 0u;
}

__global__ void __launch_bounds__(FMT_THREADS)
fmt_len_kernel(const int32_t *__restrict__ val, int64_t n, uint32_t *__restrict__ block_len) {
    __shared__ uint32_t s_w[FMT_THREADS / kWarp];
    const int64_t i = ((int64_t)blockIdx.x * FMT_THREADS + threadIdx.x) * FMT_VPT;
    int32_t v[FMT_VPT];
    uint32_t len[FMT_VPT];
    load4(val, i, n, v, len);
    uint32_t c = warp_sum(len[0] + len[1] + len[2] + len[3]);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < FMT_THREADS / kWarp; ++w) t += s_w[w];
        block_len[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(FMT_THREADS)
fmt_emit_kernel(const int32_t *__restrict__ val, int64_t n, const uint32_t *__restrict__ block_off,
                unsigned char *__restrict__ text, uint64_t text_bytes) {
    __shared__ uint32_t s_w[FMT_THREADS / kWarp];
    __shared__ __align__(16) unsigned char s_tile[FMT_BLOCK_VALUES * FMT_MAX_LEN + 16];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = ((int64_t)blockIdx.x * FMT_THREADS + threadIdx.x) * FMT_VPT;
    int32_t v[FMT_VPT];
    uint32_t len[FMT_VPT];
    load4(val, i, n, v, len);
    const uint32_t mine = len[0] + len[1] + len[2] + len[3];
    const uint32_t incl = warp_incl_scan(mine, lane);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t wexcl = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < FMT_THREADS / kWarp; ++w) {
        if (w < warp) wexcl += s_w[w];
        total += s_w[w];
    }
    // the tile is laid out so that its byte `lead` is the block's first output byte and
    // global 16-byte rows line up with shared 16-byte rows
    const uint64_t g0 = block_off[blockIdx.x];
    const uint32_t lead = (uint32_t)(g0 & 15u);
    uint32_t o = lead + wexcl + incl - mine;
#pragma unroll
    for (int k = 0; k < FMT_VPT; ++k) {
        if (len[k] == 0) continue;
        const bool neg = v[k] < 0;
        uint32_t u = neg ? 0u - (uint32_t)v[k] : (uint32_t)v[k];
        uint32_t e = o + len[k] - 1;                         // separator slot
        s_tile[e] = '\n';
        do {
            s_tile[--e] = (unsigned char)('0' + u % 10u);
            u /= 10u;
        } while (u);
        if (neg) s_tile[--e] = '-';
        o += len[k];
    }
    __syncthreads();
    // [g0, g0 + total) clipped to text_bytes (the last value's separator is not part of the text)
    const uint64_t g_end = g0 + total < text_bytes ? g0 + total : text_bytes;
    const uint64_t row0 = g0 & ~(uint64_t)15;
    for (uint64_t r = row0 + (uint64_t)threadIdx.x * 16; r < g_end; r += (uint64_t)FMT_THREADS * 16) {
        const uint32_t so = (uint32_t)(r - row0);
        if (r >= g0 && r + 16 <= g_end) {
            *reinterpret_cast<uint4 *>(text + r) = *reinterpret_cast<const uint4 *>(s_tile + so);
        } else {
            for (uint32_t b = 0; b < 16; ++b)
                if (r + b >= g0 && r + b < g_end) text[r + b] = s_tile[so + b];
        }
    }
}

uint32_t fmt_blocks(int64_t n) { return (uint32_t)((n + FMT_BLOCK_VALUES - 1) / FMT_BLOCK_VALUES); }

int launch_fmt_len(const int32_t *val, int64_t n, uint32_t *block_len, cudaStream_t s) {
    fmt_len_kernel<<<fmt_blocks(n), FMT_THREADS, 0, s>>>(val, n, block_len);
    return 1;
}
int launch_fmt_emit(const int32_t *val, int64_t n, const uint32_t *block_off, unsigned char *text,
                    uint64_t text_bytes, cudaStream_t s) {
    fmt_emit_kernel<<<fmt_blocks(n), FMT_THREADS, 0, s>>>(val, n, block_off, text, text_bytes);
    return 1;
}

// Load this file's kernels now (CUDA loads them lazily, on first launch): a first launch that
// has to load code while another context's kernel spin-waits for this one can stall behind it.
void preload_format_text() {
    preload_one(reinterpret_cast<const void *>(&fmt_emit_kernel));
    preload_one(reinterpret_cast<const void *>(&fmt_len_kernel));
}

}  // namespace adb
