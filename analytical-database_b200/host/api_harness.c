/*
 * api_harness.c -- a C caller of the operator API, for measurement.
 *
 * The reference's operators are called from C (execute_DbOperator, src/server.c:137-435), a
 * few hundred nanoseconds per call; driving them from Python through ctypes costs microseconds
 * per call, which is noise for a 3 ms step on one GPU and 5 % of a 0.45 ms step on eight.  This
 * file is the dispatcher's s=select / f=fetch / a=sum(f) sequence (server.c:137-290) written the
 * way the dispatcher writes it -- one operator after the other, every Status checked, the long
 * read from the scalar Result, the three handles released the way client_context.c releases
 * them (free(payload), free(result), with the release hook) -- in a loop, timed with the wall
 * clock.  Built into libadb_harness.so next to libadb_query.so; bench.py's e2e leg calls it.
 * It computes nothing and includes no engine header: it only sees include/adb_query_api.h.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdlib.h>
#include <time.h>

#include "adb_query_api.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void drop(Result *r) {                      /* client_context.c:31-45 + the release hook */
    if (!r) return;
    adb_host_result_release(r);
    free(r->payload);
    free(r);
}

/* `steps` passes over n_pairs (select column, fetch column) pairs with the predicate
 * lo <= v < hi.  Returns 0 and the last pass's table-wide sum and hit count, plus the wall time
 * of all passes; -1 when an operator failed (adb_host_last_error()). */
int adb_harness_select_fetch_sum(Column **sel_cols, Column **fetch_cols, int n_pairs, int lo, int hi,
                                 int steps, long *sum_out, size_t *hits_out, double *seconds_out) {
    long total = 0;
    size_t hits = 0;
    const double t0 = now_s();
    for (int k = 0; k < steps; ++k) {
        total = 0;
        hits = 0;
        for (int i = 0; i < n_pairs; ++i) {
            Status st;
            int low = lo, high = hi;
            Result *s = select_column(sel_cols[i], &low, &high, &st);
            if (!s || st.code != OK) return -1;
            Result *f = fetch_column(fetch_cols[i], s, &st);
            if (!f || st.code != OK) { drop(s); return -1; }
            GeneralizedColumn gc;
            gc.column_type = RESULT;
            gc.column_pointer.result = f;
            Result *a = sum(&gc, &st);
            if (!a || st.code != OK) { drop(s); drop(f); return -1; }
            total += *(long *)a->payload;
            hits += s->num_tuples;
            drop(s);
            drop(f);
            drop(a);
        }
    }
    if (seconds_out) *seconds_out = now_s() - t0;
    if (sum_out) *sum_out = total;
    if (hits_out) *hits_out = hits;
    return 0;
}
