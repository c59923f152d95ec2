/*
 * ref_harness.c -- flat-array adaptor around the UNMODIFIED reference operators.
 *
 * TEST INFRASTRUCTURE ONLY (see adb_oracle.c).  This file is compiled together
 * with /root/reference/src/{query,index,multimap,utils}.c -- read where they lie,
 * never copied -- into oracle/_ref/libref_O0.so / libref_O2.so by oracle/Makefile.
 * It only builds Column / Result / GeneralizedColumn / SelectOperator structs
 * (reference headers: src/include/cs165_api.h, db_manager.h, query.h) around the
 * caller's arrays and forwards to the reference function named in each comment.
 * Entry points mirror the orc_* signatures in adb_oracle.c one for one so the
 * tests can diff restatement against reference on the same inputs.
 */
#define _DEFAULT_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "cs165_api.h"
#include "db_manager.h"
#include "query.h"
#include "multimap.h"

/* non-static in src/query.c but absent from query.h */
Result *select_column_scan(Column *column, int *low_pointer, int *high_pointer, Status *ret_status);
Result *select_column_sorted_index(Column *column, int low, int high, Status *ret_status);
int get_proper_size(int tuple_num);     /* src/multimap.c:30 */

#define REF_API __attribute__((visibility("default")))

static Column mk_column(const int32_t *data, int64_t n) {
    Column c;
    memset(&c, 0, sizeof c);
    c.data = (int *)data;
    c.row_count = (size_t)n;
    return c;
}
static Result mk_result(const int32_t *payload, int64_t n) {
    Result r;
    r.num_tuples = (size_t)n;
    r.data_type = INT;
    r.payload = (void *)payload;
    return r;
}
static int64_t take(Result *r, int32_t *out) {
    int64_t h = (int64_t)r->num_tuples;
    if (h > 0) memcpy(out, r->payload, (size_t)h * sizeof(int32_t));
    free(r->payload);
    free(r);
    return h;
}

/* select_column_scan, src/query.c:92 */
REF_API int64_t ref_select_scan(const int32_t *data, int64_t n, const int32_t *lo,
                                const int32_t *hi, int32_t *out_pos) {
    Column c = mk_column(data, n);
    Status st;
    return take(select_column_scan(&c, (int *)lo, (int *)hi, &st), out_pos);
}

/* select_result, src/query.c:38 */
REF_API int64_t ref_select_result(const int32_t *val, const int32_t *pos, int64_t n,
                                  const int32_t *lo, const int32_t *hi, int32_t *out_pos) {
    Result v = mk_result(val, n), p = mk_result(pos, n);
    Status st;
    return take(select_result(&v, &p, (int *)lo, (int *)hi, &st), out_pos);
}

/* select_column via the index route, src/query.c:203-217 -> :165.  Callers must stay
 * inside the defined domain (low, high >= values[0], n > 0) or this crashes. */
REF_API int64_t ref_select_sorted_index(const int32_t *values, const uint64_t *positions,
                                        int64_t n, int32_t low, int32_t high,
                                        int32_t *out_pos, int32_t *undefined_out) {
    Column c = mk_column(NULL, n);
    ColumnIndex ix;
    ix.values = (int *)values;
    ix.positions = (size_t *)positions;
    c.index = &ix;
    c.has_index = true;
    Status st;
    int lo = low, hi = high;
    if (undefined_out) *undefined_out = 0;
    return take(select_column(&c, &lo, &hi, &st), out_pos);
}

/* fetch_column, src/query.c:223 */
REF_API void ref_fetch(const int32_t *data, const int32_t *pos, int64_t h, int32_t *out_val) {
    Column c = mk_column(data, 0);
    Result p = mk_result(pos, h);
    Status st;
    take(fetch_column(&c, &p, &st), out_val);
}

/* sum over a Result (src/query.c:329-335) and over a Column (:336-341) */
REF_API int64_t ref_sum(const int32_t *v, int64_t n) {
    Result r = mk_result(v, n);
    GeneralizedColumn g;
    g.column_type = RESULT;
    g.column_pointer.result = &r;
    Status st;
    Result *out = sum(&g, &st);
    int64_t s = *(long *)out->payload;
    free(out->payload); free(out);
    return s;
}
REF_API int64_t ref_sum_column(const int32_t *v, int64_t n) {
    Column c = mk_column(v, n);
    GeneralizedColumn g;
    g.column_type = COLUMN;
    g.column_pointer.column = &c;
    Status st;
    Result *out = sum(&g, &st);
    int64_t s = *(long *)out->payload;
    free(out->payload); free(out);
    return s;
}
/* average, src/query.c:306 */
REF_API double ref_avg(const int32_t *v, int64_t n) {
    Result r = mk_result(v, n);
    Status st;
    Result *out = average(&r, &st);
    double a = *(double *)out->payload;
    free(out->payload); free(out);
    return a;
}
/* min / max, src/query.c:392,417 (n must be > 0) */
REF_API int32_t ref_min(const int32_t *v, int64_t n) {
    Result r = mk_result(v, n);
    Status st;
    Result *out = min(&r, &st);
    int32_t m = *(int *)out->payload;
    free(out->payload); free(out);
    return m;
}
REF_API int32_t ref_max(const int32_t *v, int64_t n) {
    Result r = mk_result(v, n);
    Status st;
    Result *out = max(&r, &st);
    int32_t m = *(int *)out->payload;
    free(out->payload); free(out);
    return m;
}
/* add / sub, src/query.c:356,374 */
REF_API void ref_add(const int32_t *a, const int32_t *b, int64_t n, int32_t *out) {
    Result x = mk_result(a, n), y = mk_result(b, n);
    Status st;
    take(add(&x, &y, &st), out);
}
REF_API void ref_sub(const int32_t *a, const int32_t *b, int64_t n, int32_t *out) {
    Result x = mk_result(a, n), y = mk_result(b, n);
    Status st;
    take(sub(&x, &y, &st), out);
}

/* shared_select, src/query.c:496.  col_min / col_max feed the reference's value-range
 * slicing (query.c:506-521); the caller keeps 2*((max-min)/3) <= n.
 * Note the reference allocates q_count * n ints of output (query.c:556). */
REF_API void ref_shared_select(const int32_t *data, int64_t n, const int32_t *lows,
                               const int32_t *highs, int32_t q_count,
                               int32_t *const *out_pos, int64_t *counts,
                               int32_t col_min, int32_t col_max) {
    Column c = mk_column(data, n);
    c.min = col_min;
    c.max = col_max;
    SelectOperator *ops = calloc((size_t)q_count, sizeof(SelectOperator));
    for (int32_t q = 0; q < q_count; ++q) {
        ops[q].low = lows[q]; ops[q].high = highs[q];
        ops[q].has_low = ops[q].has_high = 1;
        ops[q].column = &c;
    }
    Status st;
    Result **res = shared_select(ops, q_count, &c, &st);
    for (int32_t q = 0; q < q_count; ++q) counts[q] = take(res[q], out_pos[q]);
    free(res);
    free(ops);
}

static int64_t take_pair(Result **res, int32_t **o1, int32_t **o2) {
    int64_t m = (int64_t)res[0]->num_tuples;
    *o1 = res[0]->payload; *o2 = res[1]->payload;
    free(res[0]); free(res[1]); free(res);
    return m;
}
/* hash_join, src/query.c:652 (keys >= 0, n1 >= 4) */
REF_API int64_t ref_hash_join(const int32_t *v1, const int32_t *p1, int64_t n1,
                              const int32_t *v2, const int32_t *p2, int64_t n2,
                              int32_t **o1, int32_t **o2) {
    Result a = mk_result(v1, n1), b = mk_result(p1, n1), c = mk_result(v2, n2), d = mk_result(p2, n2);
    Status st;
    return take_pair(hash_join(&a, &b, &c, &d, &st), o1, o2);
}
/* nested_loop_join, src/query.c:585 */
REF_API int64_t ref_nested_loop_join(const int32_t *v1, const int32_t *p1, int64_t n1,
                                     const int32_t *v2, const int32_t *p2, int64_t n2,
                                     int32_t **o1, int32_t **o2) {
    Result a = mk_result(v1, n1), b = mk_result(p1, n1), c = mk_result(v2, n2), d = mk_result(p2, n2);
    Status st;
    return take_pair(nested_loop_join(&a, &b, &c, &d, &st), o1, o2);
}
REF_API void ref_free(void *p) { free(p); }
REF_API int32_t ref_multimap_size(int32_t tuple_num, int32_t unused) {
    (void)unused;
    return get_proper_size(tuple_num);
}

/* build_unclustered_index's sort, src/index.c:140-143 (init_column_index + quicksort) */
REF_API void ref_index_sort(const int32_t *data, int64_t n, int32_t *values, uint64_t *positions) {
    Column c = mk_column(data, n);
    init_column_index(&c);
    quicksort(c.index->values, c.index->positions, 0, (int)n - 1);
    memcpy(values, c.index->values, (size_t)n * sizeof(int32_t));
    memcpy(positions, c.index->positions, (size_t)n * sizeof(uint64_t));
    free(c.index->values); free(c.index->positions); free(c.index);
}
/* reorder_column, src/index.c:105 (stack VLA: keep n below ~1 M) */
REF_API void ref_reorder(int32_t *data, int64_t n, const uint64_t *sorted_positions) {
    Column c = mk_column(data, n);
    reorder_column(&c, (size_t *)sorted_positions);
}

/* ---- the north-star chain for CPU timing: select_column_scan -> fetch_column -> sum,
 * each Result materialised by the reference's own malloc'ing operators. */
REF_API int64_t ref_chain_select_fetch_sum(const int32_t *sel_col, const int32_t *fetch_col,
                                           int64_t n, const int32_t *lo, const int32_t *hi,
                                           int64_t *hits_out) {
    Column cs = mk_column(sel_col, n), cf = mk_column(fetch_col, n);
    Status st;
    Result *pos = select_column_scan(&cs, (int *)lo, (int *)hi, &st);
    Result *val = fetch_column(&cf, pos, &st);
    GeneralizedColumn g;
    g.column_type = RESULT;
    g.column_pointer.result = val;
    Result *s = sum(&g, &st);
    int64_t out = *(long *)s->payload;
    if (hits_out) *hits_out = (int64_t)pos->num_tuples;
    free(pos->payload); free(pos); free(val->payload); free(val); free(s->payload); free(s);
    return out;
}

/* One reference instance per contiguous row range on `threads` host threads (the
 * "all host cores" figure of BASELINE.md section 3.3).  Partials are added here. */
typedef struct {
    const int32_t *sel, *fet; int64_t n; const int32_t *lo, *hi; int64_t sum, hits;
} chain_job;
static void *chain_worker(void *arg) {
    chain_job *j = arg;
    j->sum = ref_chain_select_fetch_sum(j->sel, j->fet, j->n, j->lo, j->hi, &j->hits);
    return NULL;
}
REF_API int64_t ref_chain_select_fetch_sum_mt(const int32_t *sel_col, const int32_t *fetch_col,
                                              int64_t n, const int32_t *lo, const int32_t *hi,
                                              int32_t threads, int64_t *hits_out) {
    if (threads < 1) threads = 1;
    pthread_t *tid = malloc((size_t)threads * sizeof *tid);
    chain_job *jobs = malloc((size_t)threads * sizeof *jobs);
    int64_t per = (n + threads - 1) / threads, total = 0, hits = 0;
    for (int32_t t = 0; t < threads; ++t) {
        int64_t b = (int64_t)t * per, e = b + per > n ? n : b + per;
        if (b > n) b = e = n;
        jobs[t] = (chain_job){sel_col + b, fetch_col + b, e - b, lo, hi, 0, 0};
        pthread_create(&tid[t], NULL, chain_worker, &jobs[t]);
    }
    for (int32_t t = 0; t < threads; ++t) {
        pthread_join(tid[t], NULL);
        total += jobs[t].sum; hits += jobs[t].hits;
    }
    free(tid); free(jobs);
    if (hits_out) *hits_out = hits;
    return total;
}

/* print, src/query.c:245-304, for one INT result.  Returns the strlen of the text; the text
 * itself is copied to out (capacity cap) when out != NULL.  The reference sizes its buffer at
 * 11 bytes per tuple (query.c:253): only call with values whose "%d\n" fits that on average. */
REF_API int64_t ref_print_i32(const int32_t *v, int64_t n, char *out, int64_t cap) {
    Result r = mk_result(v, n);
    Result *rs[1] = {&r};
    Status st;
    char *text = print(rs, 1, &st);
    int64_t len = 0;
    if (n > 0) {
        len = (int64_t)strlen(text);
        if (out && len <= cap) memcpy(out, text, (size_t)len);
    }
    free(text);
    return len;
}
