/*
 * adb_oracle.c -- CPU restatement of the reference column store's operator path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the CPU baseline.  The
 * engine (analytical-database_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  Every function below is diff-tested against the
 * reference's own objects (oracle/_ref/libref_*.so, built by oracle/Makefile
 * from /root/reference/src/{query,index,multimap,utils}.c where they lie) in
 * tests/test_oracle_vs_ref.py, and against the reference's project_tests
 * golden .dsl/.exp pairs in tests/test_golden_dsl.py.
 *
 * All arrays are flat int32; sizes are int64.  The reference keeps positions and
 * counters in `int` (src/query.c:40-43,94-95), so every entry point is only
 * defined for n < 2^31.  Citations are path:line under /root/reference.
 *
 * "oracle-undefined" marks inputs on which the reference reads out of bounds or
 * crashes; for those the restatement returns the result
 * that select_column_scan's predicate (low <= v < high) implies, and sets
 * *undefined_out when the caller passed one.
 */
#define _DEFAULT_SOURCE              /* strsep (load_db's tokeniser) under -std=c99 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---- range select over a base column: src/query.c:92-137 -------------------
 * Four specialised loops, one per combination of present bounds; positions are
 * emitted in ascending row order.  lo/hi == NULL means "bound absent"
 * (src/server.c:144-154).  Returns the hit count (Result.num_tuples). */
ORC_API int64_t orc_select_scan(const int32_t *data, int64_t n, const int32_t *lo,
                                const int32_t *hi, int32_t *out_pos) {
    int64_t h = 0;
    if (lo && hi) {                       /* query.c:97-104 */
        const int32_t l = *lo, u = *hi;
        for (int64_t r = 0; r < n; ++r)
            if (data[r] >= l && data[r] < u) out_pos[h++] = (int32_t)r;
    } else if (hi) {                      /* query.c:105-112 */
        const int32_t u = *hi;
        for (int64_t r = 0; r < n; ++r)
            if (data[r] < u) out_pos[h++] = (int32_t)r;
    } else if (lo) {                      /* query.c:113-120 */
        const int32_t l = *lo;
        for (int64_t r = 0; r < n; ++r)
            if (data[r] >= l) out_pos[h++] = (int32_t)r;
    } else {                              /* query.c:121-127: identity list */
        for (int64_t r = 0; r < n; ++r) out_pos[r] = (int32_t)r;
        h = n;
    }
    return h;
}

/* ---- range select over an intermediate (value, position) pair list:
 * src/query.c:38-86.  Emits the paired position, order preserved. */
ORC_API int64_t orc_select_result(const int32_t *val, const int32_t *pos, int64_t n,
                                  const int32_t *lo, const int32_t *hi, int32_t *out_pos) {
    int64_t h = 0;
    if (lo && hi) {                       /* query.c:45-52 */
        const int32_t l = *lo, u = *hi;
        for (int64_t i = 0; i < n; ++i)
            if (val[i] >= l && val[i] < u) out_pos[h++] = pos[i];
    } else if (hi) {                      /* query.c:53-60 */
        const int32_t u = *hi;
        for (int64_t i = 0; i < n; ++i)
            if (val[i] < u) out_pos[h++] = pos[i];
    } else if (lo) {                      /* query.c:61-68 */
        const int32_t l = *lo;
        for (int64_t i = 0; i < n; ++i)
            if (val[i] >= l) out_pos[h++] = pos[i];
    } else {                              /* query.c:69-75 */
        for (int64_t i = 0; i < n; ++i) out_pos[i] = pos[i];
        h = n;
    }
    return h;
}

/* ---- sorted-index range select: src/query.c:143-198 -------------------------
 * binary_search (query.c:143-160) returns the index of *some* element equal to
 * the target, else the index of the largest element below it.  Its `right`
 * cursor is a size_t, so a target below values[0] (or n == 0) underflows and
 * reads out of bounds: oracle-undefined.  In the defined domain the function
 * returns exactly what the reference returns, including its one quirk: when no
 * value lies in [low, high) but the first value >= low equals `high`, the
 * right-hand walk stops at right == left and one spurious position is emitted
 * (query.c:181-188; SURVEY.md A4 "low == high"). */
static int64_t orc_bsearch(const int32_t *a, int64_t n, int32_t target) {
    int64_t left = 0, right = n - 1;      /* caller guarantees target >= a[0], n > 0 */
    while (left <= right) {
        int64_t mid = (left + right) / 2;
        if (a[mid] == target) return mid;
        if (target < a[mid]) right = mid - 1; else left = mid + 1;
    }
    return right;
}

ORC_API int64_t orc_select_sorted_index(const int32_t *values, const uint64_t *positions,
                                        int64_t n, int32_t low, int32_t high,
                                        int32_t *out_pos, int32_t *undefined_out) {
    int64_t h = 0;
    if (undefined_out) *undefined_out = 0;
    if (n <= 0 || low < values[0] || high < values[0]) {
        /* oracle-undefined: reference SIGSEGVs (size_t underflow, query.c:145-153).
         * Defined here as the scan predicate applied to the index, in index order. */
        if (undefined_out) *undefined_out = 1;
        for (int64_t r = 0; r < n; ++r)
            if (values[r] >= low && values[r] < high) out_pos[h++] = (int32_t)positions[r];
        return h;
    }
    int64_t left = orc_bsearch(values, n, low);      /* query.c:172 */
    int64_t right = orc_bsearch(values, n, high);    /* query.c:173 */
    while (left > 0 && values[left] >= low) left--;  /* query.c:175-177 */
    if (values[left] != low) left++;                 /* query.c:178-180 */
    while (right > left && values[right] == high) right--;   /* query.c:181-183 */
    for (int64_t r = left; r <= right; ++r)          /* query.c:185-188 (size_t -> int) */
        out_pos[h++] = (int32_t)positions[r];
    return h;
}

/* ---- fetch (gather): src/query.c:223-243 ---- */
ORC_API void orc_fetch(const int32_t *data, const int32_t *pos, int64_t h, int32_t *out_val) {
    for (int64_t i = 0; i < h; ++i) out_val[i] = data[pos[i]];   /* query.c:229-231 */
}

/* ---- aggregates: src/query.c:306-437 ----
 * sum accumulates int32 into a C `long` (int64 on LP64), query.c:326-341, over a
 * Result or a whole Column -- the loop is the same, so one entry point. */
ORC_API int64_t orc_sum(const int32_t *v, int64_t n) {
    int64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) acc += v[i];
    return acc;
}
/* average: (double)sum / (double)num_tuples, query.c:308-314; n == 0 gives NaN. */
ORC_API double orc_avg(const int32_t *v, int64_t n) {
    return (double)orc_sum(v, n) / (double)n;
}
/* min / max seed with payload[0] (query.c:395,420); n == 0 is an out-of-bounds
 * read in the reference: oracle-undefined, defined here as INT32_MAX / INT32_MIN. */
ORC_API int32_t orc_min(const int32_t *v, int64_t n) {
    if (n <= 0) return INT32_MAX;
    int32_t m = v[0];
    for (int64_t i = 0; i < n; ++i) if (m > v[i]) m = v[i];     /* query.c:397-402 */
    return m;
}
ORC_API int32_t orc_max(const int32_t *v, int64_t n) {
    if (n <= 0) return INT32_MIN;
    int32_t m = v[0];
    for (int64_t i = 0; i < n; ++i) if (m < v[i]) m = v[i];     /* query.c:422-427 */
    return m;
}

/* ---- element-wise add / sub: src/query.c:356-390.  The reference adds signed
 * ints; with gcc on x86-64 that wraps two's-complement (SURVEY.md A7 [probe]),
 * restated here with unsigned arithmetic so the wrap is defined. */
ORC_API void orc_add(const int32_t *a, const int32_t *b, int64_t n, int32_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)((uint32_t)a[i] + (uint32_t)b[i]);
}
ORC_API void orc_sub(const int32_t *a, const int32_t *b, int64_t n, int32_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)((uint32_t)a[i] - (uint32_t)b[i]);
}

/* ---- batched shared scan: src/query.c:450-583 -------------------------------
 * One pass over the column, every row tested against every query's [low, high)
 * (query.c:472-479; has_low/has_high are ignored, a "null" bound is the 0 the
 * parser stored, parse.c:435-449).  Per query the three thread slices are
 * concatenated in slice order (query.c:563-574), i.e. ascending row order.  The
 * slices are cut by value range (query.c:506-521) and only tile the column when
 * 2*((max-min)/3) <= row_count; otherwise the reference reads past the column
 * (SURVEY.md A6): oracle-undefined, defined here as the full-column scan.
 * out_pos[q] must hold n ints; counts[q] receives the hit count. */
ORC_API void orc_shared_select(const int32_t *data, int64_t n, const int32_t *lows,
                               const int32_t *highs, int32_t q_count,
                               int32_t *const *out_pos, int64_t *counts) {
    for (int32_t q = 0; q < q_count; ++q) counts[q] = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int32_t v = data[r];
        for (int32_t q = 0; q < q_count; ++q)
            if (v >= lows[q] && v < highs[q]) out_pos[q][counts[q]++] = (int32_t)r;
    }
}

/* ---- hash join: src/query.c:652-696 + src/multimap.c:15-102 -----------------
 * Open-addressing multimap: size = first prime >= (int)(1.3*n1) by trial
 * division (multimap.c:15-38), slot = key % size (multimap.c:60-63), linear
 * probe until an empty slot (values_size == 0) or the key (multimap.c:65-71),
 * each slot owning a doubling value list (multimap.c:74-89).  Build on side one
 * in row order, probe with side two in row order; every match appends
 * (stored position, probe position) (query.c:664-681).  Output is therefore
 * probe-major, and within a key in build insertion order.
 * Negative keys index out of bounds (C remainder) and n1 == 0 divides by zero:
 * oracle-undefined; here negative keys use the non-negative remainder and an
 * empty build side yields no matches -- the pair list is what the equi-join
 * definition implies either way.
 * Returns the match count; *o1 / *o2 are malloc'd (release with orc_free). */
typedef struct { int32_t key; int32_t cnt, cap; int32_t *vals; } orc_slot;

static int orc_is_prime(int32_t x) {            /* multimap.c:15-27 */
    int32_t i;
    for (i = 2; i <= x / 2; ++i) if (x % i == 0) break;
    return i > x / 2;
}
static int orc_is_prime_fast(int32_t x) {       /* same predicate, sqrt bound */
    if (x < 4) return 1;                        /* multimap.c's loop accepts 0..3 */
    if (x % 2 == 0) return 0;
    for (int64_t i = 3; i * i <= x; i += 2) if (x % i == 0) return 0;
    return 1;
}
ORC_API int32_t orc_multimap_size(int32_t tuple_num, int32_t exact_trial_division) {
    int32_t s = (int32_t)(1.3 * tuple_num);     /* multimap.c:30-31 */
    while (!(exact_trial_division ? orc_is_prime(s) : orc_is_prime_fast(s))) s += 1;
    return s;
}

ORC_API int64_t orc_hash_join(const int32_t *v1, const int32_t *p1, int64_t n1,
                              const int32_t *v2, const int32_t *p2, int64_t n2,
                              int32_t **o1, int32_t **o2) {
    int64_t cap = 512, m = 0;                   /* PAGE_SIZE, query.c:654 */
    int32_t *a = malloc(cap * sizeof(int32_t)), *b = malloc(cap * sizeof(int32_t));
    *o1 = a; *o2 = b;
    if (n1 <= 0) return 0;                      /* oracle-undefined (key % 0) */
    const int32_t size = orc_multimap_size((int32_t)n1, 0);
    orc_slot *tab = calloc((size_t)size, sizeof(orc_slot));
    for (int64_t i = 0; i < n1; ++i) {          /* query.c:664-666 */
        int32_t idx = v1[i] % size; if (idx < 0) idx += size;
        while (tab[idx].cnt != 0 && tab[idx].key != v1[i]) idx = (idx + 1) % size;
        orc_slot *s = &tab[idx];
        s->key = v1[i];
        if (s->cnt == s->cap) {                 /* multimap.c:81-85 */
            s->cap = s->cap ? s->cap * 2 : 1;
            s->vals = realloc(s->vals, (size_t)s->cap * sizeof(int32_t));
        }
        s->vals[s->cnt++] = p1[i];
    }
    for (int64_t j = 0; j < n2; ++j) {          /* query.c:669-681 */
        int32_t idx = v2[j] % size; if (idx < 0) idx += size;
        int32_t steps = 0;                      /* a full table (n1 <= 3) never ends the
                                                 * reference's probe: oracle-undefined */
        while (tab[idx].cnt != 0 && tab[idx].key != v2[j] && steps++ < size) idx = (idx + 1) % size;
        const orc_slot *s = &tab[idx];
        if (s->cnt != 0 && s->key != v2[j]) continue;
        for (int32_t k = 0; k < s->cnt; ++k) {
            if (m == cap) {
                cap *= 2;
                a = realloc(a, cap * sizeof(int32_t));
                b = realloc(b, cap * sizeof(int32_t));
            }
            a[m] = s->vals[k]; b[m] = p2[j]; ++m;
        }
    }
    for (int32_t i = 0; i < size; ++i) free(tab[i].vals);
    free(tab);
    *o1 = a; *o2 = b;
    return m;
}

/* ---- nested-loop join: src/query.c:585-650.  Outer-major over side one. ---- */
ORC_API int64_t orc_nested_loop_join(const int32_t *v1, const int32_t *p1, int64_t n1,
                                     const int32_t *v2, const int32_t *p2, int64_t n2,
                                     int32_t **o1, int32_t **o2) {
    int64_t cap = 512, m = 0;
    int32_t *a = malloc(cap * sizeof(int32_t)), *b = malloc(cap * sizeof(int32_t));
    for (int64_t i = 0; i < n1; ++i)
        for (int64_t j = 0; j < n2; ++j)
            if (v1[i] == v2[j]) {               /* query.c:599-609 */
                if (m == cap) {
                    cap *= 2;
                    a = realloc(a, cap * sizeof(int32_t));
                    b = realloc(b, cap * sizeof(int32_t));
                }
                a[m] = p1[i]; b[m] = p2[j]; ++m;
            }
    *o1 = a; *o2 = b;
    return m;
}

ORC_API void orc_free(void *p) { free(p); }

/* ---- sorted index build: src/index.c:25-46,89-147 ---------------------------
 * (values, positions) start as (copy of column, identity) (index.c:89-101) and
 * are sorted by a Lomuto quicksort whose pivot is the last element and whose
 * comparison is strict `<` (index.c:33-46).  The tie order of equal values is
 * whatever that unstable sort leaves; this restatement performs the identical
 * swap sequence inside each partition, driving the recursion from an explicit
 * stack (sub-ranges are disjoint, so their processing order cannot change the
 * result). */
static int64_t orc_partition(int32_t *v, uint64_t *p, int64_t low, int64_t high) {
    const int32_t pivot = v[high];
    int64_t i = low - 1;
    for (int64_t j = low; j < high; ++j)
        if (v[j] < pivot) {
            ++i;
            int32_t tv = v[i]; v[i] = v[j]; v[j] = tv;
            uint64_t tp = p[i]; p[i] = p[j]; p[j] = tp;
        }
    int32_t tv = v[i + 1]; v[i + 1] = v[high]; v[high] = tv;
    uint64_t tp = p[i + 1]; p[i + 1] = p[high]; p[high] = tp;
    return i + 1;
}
ORC_API void orc_index_sort(const int32_t *data, int64_t n, int32_t *values, uint64_t *positions) {
    memcpy(values, data, (size_t)n * sizeof(int32_t));
    for (int64_t i = 0; i < n; ++i) positions[i] = (uint64_t)i;
    if (n < 2) return;
    int64_t cap = 64, top = 0;
    int64_t *stk = malloc(2 * cap * sizeof(int64_t));
    stk[0] = 0; stk[1] = n - 1; top = 1;
    while (top > 0) {
        --top;
        int64_t lo = stk[2 * top], hi = stk[2 * top + 1];
        while (lo < hi) {
            int64_t pv = orc_partition(values, positions, lo, hi);
            /* defer the larger side, iterate on the smaller: bounded stack */
            int64_t l0 = lo, l1 = pv - 1, r0 = pv + 1, r1 = hi;
            int64_t d0, d1;
            if (l1 - l0 > r1 - r0) { d0 = l0; d1 = l1; lo = r0; hi = r1; }
            else                   { d0 = r0; d1 = r1; lo = l0; hi = l1; }
            if (d0 < d1) {
                if (top == cap) { cap *= 2; stk = realloc(stk, 2 * cap * sizeof(int64_t)); }
                stk[2 * top] = d0; stk[2 * top + 1] = d1; ++top;
            }
        }
    }
    free(stk);
}

/* ---- clustered reorder of a sibling column: src/index.c:105-117.  The reference
 * copies the column onto the stack (a VLA of n ints, index.c:107) and overflows
 * beyond ~2 M rows; this restatement copies to the heap, same permutation. */
ORC_API void orc_reorder(int32_t *data, int64_t n, const uint64_t *sorted_positions) {
    int32_t *copy = malloc((size_t)n * sizeof(int32_t));
    memcpy(copy, data, (size_t)n * sizeof(int32_t));
    for (int64_t i = 0; i < n; ++i) data[i] = copy[sorted_positions[i]];
    free(copy);
}

/* ---- the north-star chain on one row range, for CPU timing -------------------
 * select (query.c:92) -> fetch (query.c:223) -> sum (query.c:325), every
 * intermediate materialised exactly as the operator API defines it. */
ORC_API int64_t orc_chain_select_fetch_sum(const int32_t *sel_col, const int32_t *fetch_col,
                                           int64_t n, const int32_t *lo, const int32_t *hi,
                                           int64_t *hits_out) {
    int32_t *pos = malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));   /* query.c:94 */
    int64_t h = orc_select_scan(sel_col, n, lo, hi, pos);
    int32_t *val = malloc((size_t)(h > 0 ? h : 1) * sizeof(int32_t));   /* query.c:226 */
    orc_fetch(fetch_col, pos, h, val);
    int64_t s = orc_sum(val, h);
    free(pos); free(val);
    if (hits_out) *hits_out = h;
    return s;
}

/* One restated chain per contiguous row range on `threads` host threads (the "all
 * host cores" CPU baseline when oracle/_ref is not available). */
#include <pthread.h>
typedef struct {
    const int32_t *sel, *fet; int64_t n; const int32_t *lo, *hi; int64_t sum, hits;
} orc_chain_job;
static void *orc_chain_worker(void *arg) {
    orc_chain_job *j = arg;
    j->sum = orc_chain_select_fetch_sum(j->sel, j->fet, j->n, j->lo, j->hi, &j->hits);
    return NULL;
}
ORC_API int64_t orc_chain_select_fetch_sum_mt(const int32_t *sel_col, const int32_t *fetch_col,
                                              int64_t n, const int32_t *lo, const int32_t *hi,
                                              int32_t threads, int64_t *hits_out) {
    if (threads < 1) threads = 1;
    pthread_t *tid = malloc((size_t)threads * sizeof *tid);
    orc_chain_job *jobs = malloc((size_t)threads * sizeof *jobs);
    int64_t per = (n + threads - 1) / threads, total = 0, hits = 0;
    for (int32_t t = 0; t < threads; ++t) {
        int64_t b = (int64_t)t * per, e = b + per > n ? n : b + per;
        if (b > n) b = e = n;
        jobs[t] = (orc_chain_job){sel_col + b, fetch_col + b, e - b, lo, hi, 0, 0};
        pthread_create(&tid[t], NULL, orc_chain_worker, &jobs[t]);
    }
    for (int32_t t = 0; t < threads; ++t) {
        pthread_join(tid[t], NULL);
        total += jobs[t].sum; hits += jobs[t].hits;
    }
    free(tid); free(jobs);
    if (hits_out) *hits_out = hits;
    return total;
}

/* ---- bulk load: the ingest loop of load_db, src/db_manager.c:304-318 -----------------
 * `text` holds the bytes of the CSV file.  fgets(line, MAX_LINE_SIZE = 1024, stream)
 * (db_manager.c:23,306) is restated over the buffer: up to 1023 bytes, stopping after a
 * '\n'.  The first `skip_lines` lines are consumed as load_db consumes the header
 * (db_manager.c:263).  Each further line is tokenised with strsep(",") and the first n_cols
 * tokens go through atoi (db_manager.c:309-311); `row` lives outside the loop, so a line
 * with fewer tokens keeps the previous line's values (oracle-undefined for the first line,
 * where the reference reads uninitialised stack: 0 here).  Every line becomes a row
 * (insert_row, db_manager.c:312).  out is column-major: out[c * rows_cap + r].
 * Returns the number of rows; with out == NULL it only counts. */
static int64_t orc_fgets(char *line, int size, const char *text, int64_t bytes, int64_t *cursor) {
    int64_t i = *cursor, n = 0;
    if (i >= bytes) return -1;                           /* EOF before any byte: NULL */
    while (n < size - 1 && i < bytes) {
        const char c = text[i++];
        line[n++] = c;
        if (c == '\n') break;
    }
    line[n] = 0;
    *cursor = i;
    return n;
}

ORC_API int64_t orc_csv_parse(const char *text, int64_t bytes, int32_t skip_lines, int32_t n_cols,
                              int32_t *out, int64_t rows_cap) {
    char line[1024];
    int64_t cursor = 0, rows = 0;
    int32_t *row = (int32_t *)calloc((size_t)(n_cols > 0 ? n_cols : 1), sizeof(int32_t));
    for (int32_t k = 0; k < skip_lines; ++k)
        if (orc_fgets(line, (int)sizeof line, text, bytes, &cursor) < 0) break;
    while (orc_fgets(line, (int)sizeof line, text, bytes, &cursor) >= 0) {
        char *temp = line, *token;
        int32_t index = 0;
        while ((token = strsep(&temp, ",")) != NULL && index < n_cols)      /* db_manager.c:309 */
            row[index++] = atoi(token);
        if (out && rows < rows_cap)
            for (int32_t c = 0; c < n_cols; ++c) out[(int64_t)c * rows_cap + rows] = row[c];
        ++rows;
    }
    free(row);
    return rows;
}

/* ---- print of one INT result: src/query.c:262-269 ---------------------------------------
 * "%d" per tuple, "\n" between tuples, nothing after the last.  out needs 12 bytes per
 * tuple (the reference allocates 11, query.c:253, and overruns on wide values -- SURVEY.md
 * A8).  Returns the text length (no NUL counted); an empty result gives "" (the reference
 * returns its uninitialised malloc(0): oracle-undefined). */
ORC_API int64_t orc_print_i32(const int32_t *v, int64_t n, char *out, int64_t cap) {
    int64_t index = 0;
    (void)cap;
    for (int64_t i = 0; i < n; ++i) {
        index += sprintf(&out[index], "%d", v[i]);
        if (i != n - 1) index += sprintf(&out[index], "\n");
    }
    return index;
}


/* ---- synthetic columns for the CPU arm of bench.py (not an operator: the counter-based
 * generator of analytical-database_b200/synth.py, restated in C so that the reference arm can
 * regenerate the whole 4 B-row table on the host in seconds) ------------------------------- */
static inline uint64_t orc_mix64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
typedef struct {
    int32_t *out;
    int64_t n;
    uint64_t seed, first_row;
    int32_t lo;
    uint32_t span;
} orc_synth_job;
static void *orc_synth_worker(void *arg) {
    orc_synth_job *j = arg;
    for (int64_t i = 0; i < j->n; ++i) {
        const uint64_t z = orc_mix64(j->seed, j->first_row + (uint64_t)i);
        j->out[i] = (int32_t)((uint32_t)j->lo + (uint32_t)(((z >> 32) * (uint64_t)j->span) >> 32));
    }
    return NULL;
}
ORC_API void orc_synth_uniform_mt(int32_t *out, int64_t n, uint64_t seed, uint64_t first_row, int32_t lo,
                                  uint32_t span, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    orc_synth_job jobs[256];
    for (int t = 0; t < threads; ++t) {
        const int64_t b = n * t / threads, e = n * (t + 1) / threads;
        jobs[t] = (orc_synth_job){out + b, e - b, seed, first_row + (uint64_t)b, lo, span};
        pthread_create(&tid[t], NULL, orc_synth_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
}
