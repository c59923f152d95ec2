/*
 * query_shim.c -- the reference's operator API (src/include/query.h:20-44) on the B200
 * engine.  This file is the C host side of the drop-in: it replaces
 * /root/reference/src/query.c and src/multimap.c in the server's link line
 * (src/Makefile:62) and forwards every operator to the C-ABI of libadb_b200.so
 * (include/adb_engine.h).  It computes nothing itself and has no CPU fallback: if the
 * engine cannot start, every operator sets ret_status->code = ERROR and returns NULL,
 * which the dispatcher turns into a "Failed ..." reply (src/server.c:171-174).
 *
 * Where things live
 *   base columns    Column.data stays the host mmap the catalog owns; the first operator
 *                   that touches a column uploads it to HBM (int32 array) and the copy is
 *                   reused until the column's data pointer or row_count changes
 *                   (insert_row may re-mmap, src/db_manager.c:178-186).  A column with a
 *                   ColumnIndex also gets its (values, positions) uploaded -- exactly the
 *                   arrays src/index.c built, so tie order is the reference's -- plus the
 *                   implicit B+-tree for `btree` indexes.
 *   results         position lists and value vectors stay in HBM.  Result.payload is a
 *                   plain malloc block of max(16, 4*num_tuples) bytes, so the unchanged
 *                   plumbing may free() it (src/client_context.c:35,82) and may read
 *                   num_tuples ints from it (the dispatcher's no-op log loop,
 *                   src/server.c:177-181); a registry maps that address to the device
 *                   buffer.  The block's bytes are NOT the tuples unless ADB_SHIM_MIRROR=1
 *                   (then every result is also copied back to the host).  Scalars
 *                   (sum / avg / min / max) are ordinary host values, as in the reference.
 *   reclaiming HBM  adb_host_result_release() (two-line patch), the free() interposer
 *                   (host/free_interpose.c, no patch), or -- last resort -- the registry
 *                   notices malloc handing out a payload address again.
 *
 * Behaviour the reference leaves undefined (it reads out of bounds or crashes; SURVEY.md
 * appendix A) is defined here and listed in DESIGN.md: NULL bounds on an indexed column,
 * bounds below the index minimum, min/max/print of an empty result, add/sub of lists of
 * different length (ERROR), negative join keys and an empty join side.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <limits.h>
#include <malloc.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "adb_engine.h"
#include "adb_query_api.h"

/* ---- state ----------------------------------------------------------------------------- */
typedef struct DevColumn {
    const Column *key;
    const int *host_data;
    size_t rows;
    int32_t *d_data;
    int adopted;                /* d_data belongs to the caller (adb_host_column_adopt) */
    /* index */
    const int *host_ix_values;
    size_t ix_rows;
    int32_t *d_ix_values, *d_ix_positions;
    adb_index *ix;
} DevColumn;

typedef struct DevResult {
    void *payload;              /* key: the host block handed out as Result.payload */
    int32_t *d_ptr;
    size_t tuples;
} DevResult;

#define SLOT_EMPTY ((void *)0)
#define SLOT_TOMB ((void *)1)

static struct {
    int up, failed, mirror, lazy;
    DevColumn *cols;
    int ncols, capcols;
    DevResult *slots;           /* open addressing on payload address */
    size_t nslots, nused;       /* nused counts live + tombstones */
    volatile long nlive;
    adb_agg *d_agg;
    pthread_mutex_t mu;
    int mu_ready;
} S;

static __thread char t_err[384];

const char *adb_host_last_error(void) { return t_err; }
long adb_host_live_device_results(void) { return S.nlive; }

static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

static void lock(void) {
    if (!S.mu_ready) {
        pthread_mutexattr_t a;
        pthread_mutexattr_init(&a);
        pthread_mutexattr_settype(&a, PTHREAD_MUTEX_RECURSIVE);
        pthread_mutex_init(&S.mu, &a);
        pthread_mutexattr_destroy(&a);
        S.mu_ready = 1;
    }
    pthread_mutex_lock(&S.mu);
}
static void unlock(void) { pthread_mutex_unlock(&S.mu); }

/* Every operator must set ret_status->code on every path: the dispatcher's Status locals
 * are uninitialised (src/server.c:141,192). */
static void *op_fail(Status *st, const char *what) {
    if (t_err[0] == '\0') set_err("%s failed", what);
    if (st) {
        st->code = ERROR;
        st->error_message = t_err;
    }
    return NULL;
}
static void op_ok(Status *st) {
    if (st) {
        st->code = OK;
        st->error_message = NULL;
    }
}
#define CK(call)                                                  \
    do {                                                          \
        if ((call) != ADB_OK) {                                   \
            set_err("%s: %s", #call, adb_last_error());           \
            goto fail;                                            \
        }                                                         \
    } while (0)

/* ---- deferred select -------------------------------------------------
 * select_column over an un-indexed column returns as soon as the hit count is known: the
 * predicate pass has left its bitmap in the engine's scratch, the position buffer is
 * allocated, its contents are not written yet.  If the next thing that happens to the handle
 * is fetch_column and then an aggregate of that fetch -- the s=select / f=fetch / a=sum(f)
 * pattern of src/server.c:137-290 -- the aggregate call resolves all three with the chain's
 * fused second kernel (positions + gather + sum/min/max in one pass over the bitmap).  Any
 * other use of either handle, any other select, a column invalidation or a release first
 * writes what is pending (flush_pending), so the handles are indistinguishable from eager
 * ones to the unchanged plumbing.  At most one select is pending at a time.
 * ADB_SHIM_EAGER=1 (or the mirror mode) turns the deferral off. */
static struct {
    int active;
    const void *sel_payload;        /* registry key of the select handle */
    int32_t *sel_d;                 /* its position buffer, h entries */
    size_t h;
    const int32_t *d_col;           /* what was scanned, to redo the count if another user of */
    size_t rows;                    /* the engine overwrote the bitmap in the meantime        */
    int has_lo, has_hi, lo, hi;
    uint64_t generation;            /* adb_select_generation() right after the count */
    const void *fetch_payload;      /* fetch_column of that select, values not written yet */
    int32_t *fetch_d;
    const int32_t *d_fetch_col;
} P;

static int pending_recount(void) {
    if (adb_select_generation() == P.generation) return 0;
    int64_t h = -1;
    if (adb_select_count(P.d_col, (int64_t)P.rows, NULL, P.has_lo ? &P.lo : NULL,
                         P.has_hi ? &P.hi : NULL, NULL, &h) != ADB_OK) {
        set_err("deferred select: %s", adb_last_error());
        return -1;
    }
    if ((size_t)h != P.h) {
        set_err("deferred select: the column changed under a pending select (%zu hits, now %lld)",
                P.h, (long long)h);
        return -1;
    }
    return 0;
}

/* Write whatever is pending; afterwards every handle is an ordinary device result. */
static int flush_pending(void) {
    if (!P.active) return 0;
    P.active = 0;
    if (pending_recount()) return -1;
    adb_status s = P.fetch_payload
        ? adb_select_emit_fetch_agg(P.d_fetch_col, P.sel_d, P.fetch_d, S.d_agg, NULL)
        : adb_select_emit(NULL, 0, P.sel_d);
    if (s != ADB_OK) {
        set_err("deferred select: %s", adb_last_error());
        return -1;
    }
    return 0;
}

/* The plumbing is about to free (or has freed) this payload. */
static void pending_payload_gone(const void *payload) {
    if (!P.active) return;
    if (payload == P.fetch_payload) {           /* nobody can read those values any more */
        P.fetch_payload = NULL;
        P.fetch_d = NULL;
    } else if (payload == P.sel_payload) {
        if (P.fetch_payload) flush_pending();   /* the fetch still needs the selected rows */
        else P.active = 0;
    }
}

/* ---- engine lifecycle ------------------------------------------------------------------ */
int adb_host_init(int device) {
    if (S.up) return 0;
    if (adb_init(device) != ADB_OK) {
        set_err("adb_init(%d): %s", device, adb_last_error());
        S.failed = 1;
        return -1;
    }
    void *p = NULL;
    if (adb_alloc(&p, sizeof(adb_agg)) != ADB_OK) {
        set_err("adb_alloc: %s", adb_last_error());
        return -1;
    }
    S.d_agg = (adb_agg *)p;
    const char *m = getenv("ADB_SHIM_MIRROR");
    S.mirror = m && m[0] && m[0] != '0';
    /* Result payloads are plain malloc blocks of 4 * num_tuples bytes that nobody writes
     * (the plumbing only frees them): above glibc's mmap threshold every handle costs an
     * mmap + munmap pair (~10 us for a 20 MB list).  Serve them from the heap instead (up to 32 MB) and
     * keep the freed space.  ADB_SHIM_NO_MALLOPT=1 leaves the process's malloc policy alone. */
    const char *nm = getenv("ADB_SHIM_NO_MALLOPT");
    if (!(nm && nm[0] && nm[0] != '0')) {
        mallopt(M_MMAP_THRESHOLD, 32 << 20);         /* glibc's ceiling on 64-bit */
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
    }
    const char *eager = getenv("ADB_SHIM_EAGER");
    S.lazy = !S.mirror && !(eager && eager[0] && eager[0] != '0');
    S.up = 1;
    S.failed = 0;
    return 0;
}

static int ensure_up(void) {
    if (S.up) return 0;
    const char *d = getenv("ADB_DEVICE");
    return adb_host_init(d ? atoi(d) : 0);
}

static void dev_column_drop(DevColumn *c) {
    if (P.active && c->d_data && (c->d_data == P.d_col || c->d_data == P.d_fetch_col)) flush_pending();
    if (c->ix) adb_index_destroy(c->ix);
    if (c->d_ix_values) adb_free(c->d_ix_values);
    if (c->d_ix_positions) adb_free(c->d_ix_positions);
    if (c->d_data && !c->adopted) adb_free(c->d_data);
    const Column *key = c->key;
    memset(c, 0, sizeof *c);
    c->key = key;
}

void adb_host_shutdown(void) {
    if (!S.up) return;
    lock();
    P.active = 0;
    for (size_t i = 0; i < S.nslots; ++i)
        if (S.slots[i].payload != SLOT_EMPTY && S.slots[i].payload != SLOT_TOMB)
            adb_free(S.slots[i].d_ptr);
    DevResult *old = S.slots;
    S.slots = NULL;
    S.nslots = S.nused = 0;
    S.nlive = 0;
    for (int i = 0; i < S.ncols; ++i) dev_column_drop(&S.cols[i]);
    DevColumn *oldc = S.cols;
    S.cols = NULL;
    S.ncols = S.capcols = 0;
    adb_free(S.d_agg);
    S.d_agg = NULL;
    S.up = 0;
    unlock();
    free(old);
    free(oldc);
    adb_shutdown();
}

/* ---- base columns ----------------------------------------------------------------------- */
static DevColumn *dev_column_slot(Column *column) {
    if (!column) {
        set_err("NULL column");
        return NULL;
    }
    if (column->row_count >= ((size_t)1 << 31)) {
        set_err("column of %zu rows: positions are int (src/query.c:94-95), shard it below 2^31",
                column->row_count);
        return NULL;
    }
    DevColumn *c = NULL;
    for (int i = 0; i < S.ncols; ++i)
        if (S.cols[i].key == column) c = &S.cols[i];
    if (!c) {
        if (S.ncols == S.capcols) {
            int cap = S.capcols ? 2 * S.capcols : 16;
            DevColumn *n = realloc(S.cols, (size_t)cap * sizeof *n);
            if (!n) {
                set_err("out of host memory");
                return NULL;
            }
            S.cols = n;
            S.capcols = cap;
        }
        c = &S.cols[S.ncols++];
        memset(c, 0, sizeof *c);
        c->key = column;
    }
    return c;
}

static DevColumn *dev_column(Column *column) {
    DevColumn *c = dev_column_slot(column);
    if (!c) return NULL;
    if (c->d_data && c->host_data == column->data && c->rows == column->row_count) return c;
    dev_column_drop(c);
    void *p = NULL;
    if (adb_alloc(&p, 4 * column->row_count) != ADB_OK ||
        adb_upload(p, column->data, 4 * column->row_count) != ADB_OK) {
        set_err("column upload: %s", adb_last_error());
        if (p) adb_free(p);
        return NULL;
    }
    c->d_data = p;
    c->host_data = column->data;
    c->rows = column->row_count;
    return c;
}

int adb_host_column_upload(Column *column) {
    if (ensure_up()) return -1;
    return dev_column(column) ? 0 : -1;
}

/* The column's rows are already in HBM (a GPU-side loader put them there: SURVEY.md 8f
 * rank 1).  The shim uses d_data as is until the column's data pointer or row_count
 * changes; the buffer stays the caller's. */
int adb_host_column_adopt(Column *column, const void *d_data) {
    if (ensure_up()) return -1;
    DevColumn *c = dev_column_slot(column);
    if (!c || !d_data) return -1;
    dev_column_drop(c);
    c->d_data = (int32_t *)d_data;
    c->adopted = 1;
    c->host_data = column->data;
    c->rows = column->row_count;
    return 0;
}

void adb_host_column_invalidate(Column *column) {
    for (int i = 0; i < S.ncols; ++i)
        if (S.cols[i].key == column) dev_column_drop(&S.cols[i]);
}

/* Upload the ColumnIndex the reference built (src/index.c:89-101,119-146).  Its positions
 * are size_t on the host and are truncated to int when emitted (src/query.c:187), so the
 * device copy is int32. */
static int dev_index(DevColumn *c, Column *column) {
    if (!column->index || !column->index->values || !column->index->positions) {
        set_err("column '%s' is flagged clustered/has_index but carries no ColumnIndex", column->name);
        return -1;
    }
    if (c->ix && c->host_ix_values == column->index->values && c->ix_rows == column->row_count) return 0;
    if (c->ix) adb_index_destroy(c->ix);
    if (c->d_ix_values) adb_free(c->d_ix_values);
    if (c->d_ix_positions) adb_free(c->d_ix_positions);
    c->ix = NULL;
    c->d_ix_values = c->d_ix_positions = NULL;
    const size_t n = column->row_count;
    int32_t *pos32 = malloc(n ? 4 * n : 4);
    if (!pos32) {
        set_err("out of host memory");
        return -1;
    }
    for (size_t i = 0; i < n; ++i) pos32[i] = (int32_t)column->index->positions[i];
    void *dv = NULL, *dp = NULL;
    int rc = -1;
    if (adb_alloc(&dv, 4 * n) == ADB_OK && adb_alloc(&dp, 4 * n) == ADB_OK &&
        adb_upload(dv, column->index->values, 4 * n) == ADB_OK &&
        adb_upload(dp, pos32, 4 * n) == ADB_OK &&
        adb_index_create(dv, dp, (int64_t)n, /*with_btree=*/!column->sorted, &c->ix) == ADB_OK) {
        c->d_ix_values = dv;
        c->d_ix_positions = dp;
        c->host_ix_values = column->index->values;
        c->ix_rows = n;
        rc = 0;
    } else {
        set_err("index upload: %s", adb_last_error());
        if (dv) adb_free(dv);
        if (dp) adb_free(dp);
    }
    free(pos32);
    return rc;
}

/* ---- device-resident results ------------------------------------------------------------- */
static size_t slot_of(const void *p, size_t nslots) {
    uint64_t x = (uint64_t)(uintptr_t)p;
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 29;
    return (size_t)x & (nslots - 1);
}

static DevResult *registry_find(const void *payload) {
    if (!S.nslots || !payload) return NULL;
    for (size_t i = slot_of(payload, S.nslots);; i = (i + 1) & (S.nslots - 1)) {
        if (S.slots[i].payload == SLOT_EMPTY) return NULL;
        if (S.slots[i].payload == payload) return &S.slots[i];
    }
}

static int registry_grow(void) {
    const size_t n = S.nslots ? 2 * S.nslots : 256;
    DevResult *ns = calloc(n, sizeof *ns), *old = S.slots;
    if (!ns) return -1;
    size_t used = 0;
    for (size_t i = 0; i < S.nslots; ++i) {
        void *p = old[i].payload;
        if (p == SLOT_EMPTY || p == SLOT_TOMB) continue;
        size_t j = slot_of(p, n);
        while (ns[j].payload != SLOT_EMPTY) j = (j + 1) & (n - 1);
        ns[j] = old[i];
        ++used;
    }
    S.slots = ns;
    S.nslots = n;
    S.nused = used;
    free(old);
    return 0;
}

/* Detach the device buffer registered under `payload` (if any) and return it. */
static int32_t *registry_take(const void *payload) {
    int32_t *d = NULL;
    lock();
    DevResult *r = registry_find(payload);
    if (r) {
        d = r->d_ptr;
        r->payload = SLOT_TOMB;
        r->d_ptr = NULL;
        --S.nlive;
    }
    unlock();
    return d;
}

void adb_host_payload_freed(void *payload) {
    if (S.nlive <= 0 || !payload) return;
    pending_payload_gone(payload);
    int32_t *d = registry_take(payload);
    if (d) adb_free(d);
}

void adb_host_result_release(Result *result) {
    if (result) adb_host_payload_freed(result->payload);
}

/* Wrap a device buffer of `tuples` int32 as a Result the plumbing can own. */
static Result *new_dev_result(int32_t *d_ptr, size_t tuples) {
    Result *r = malloc(sizeof *r);
    size_t bytes = 4 * tuples < 16 ? 16 : 4 * tuples;
    void *payload = malloc(bytes);
    if (!r || !payload) {
        free(r);
        free(payload);
        adb_free(d_ptr);
        set_err("out of host memory for a %zu-tuple result", tuples);
        return NULL;
    }
    if (S.mirror) {
        if (tuples && adb_download(payload, d_ptr, 4 * tuples) != ADB_OK) {
            set_err("result mirror: %s", adb_last_error());
            free(r);
            free(payload);
            adb_free(d_ptr);
            return NULL;
        }
    }
    /* malloc returned an address we still hold a buffer for: that payload was freed by the
     * plumbing without telling us -- reclaim its HBM now.  No engine call is made with the
     * registry lock held (with the free() interposer other threads' frees wait on it). */
    if (P.active && (payload == P.sel_payload || payload == P.fetch_payload))
        pending_payload_gone(payload);          /* may still need the dead handle's buffer */
    int32_t *dead = NULL;
    lock();
    DevResult *stale = registry_find(payload);
    if (stale) {
        dead = stale->d_ptr;
        stale->d_ptr = d_ptr;
        stale->tuples = tuples;
    } else {
        if (4 * (S.nused + 1) > 3 * S.nslots && registry_grow()) {
            unlock();
            free(r);
            free(payload);
            adb_free(d_ptr);
            set_err("out of host memory");
            return NULL;
        }
        size_t j = slot_of(payload, S.nslots);
        while (S.slots[j].payload != SLOT_EMPTY && S.slots[j].payload != SLOT_TOMB)
            j = (j + 1) & (S.nslots - 1);
        if (S.slots[j].payload == SLOT_EMPTY) ++S.nused;
        S.slots[j].payload = payload;
        S.slots[j].d_ptr = d_ptr;
        S.slots[j].tuples = tuples;
        ++S.nlive;
    }
    unlock();
    if (dead) adb_free(dead);
    r->num_tuples = tuples;
    r->data_type = INT;
    r->payload = payload;
    return r;
}

static int32_t *alloc_i32(size_t n) {
    void *p = NULL;
    if (adb_alloc(&p, 4 * n) != ADB_OK) {
        set_err("adb_alloc(%zu): %s", 4 * n, adb_last_error());
        return NULL;
    }
    return p;
}

/* An operator input: the device buffer behind a Result.  A Result whose payload is not in
 * the registry is an ordinary host array (built by a test or by foreign code); it is staged
 * into a temporary device buffer that the caller releases with unstage(). */
typedef struct Staged {
    int32_t *d;
    int temp;
} Staged;

static int stage(const Result *r, Staged *out) {
    out->d = NULL;
    out->temp = 0;
    if (!r) {
        set_err("NULL result operand");
        return -1;
    }
    if (r->num_tuples >= ((size_t)1 << 31)) {
        set_err("result of %zu tuples exceeds the int position domain (src/query.c:40-43)", r->num_tuples);
        return -1;
    }
    if (flush_pending()) return -1;             /* an operand is about to be read */
    lock();
    DevResult *e = registry_find(r->payload);
    int32_t *d = e ? e->d_ptr : NULL;
    size_t have = e ? e->tuples : 0;
    unlock();
    if (d) {
        if (have < r->num_tuples) {
            set_err("result claims %zu tuples but its device buffer holds %zu", r->num_tuples, have);
            return -1;
        }
        out->d = d;
        return 0;
    }
    if (r->num_tuples && !r->payload) {
        set_err("result operand has no payload");
        return -1;
    }
    out->d = alloc_i32(r->num_tuples);
    if (!out->d) return -1;
    out->temp = 1;
    if (r->num_tuples && adb_upload(out->d, r->payload, 4 * r->num_tuples) != ADB_OK) {
        set_err("operand upload: %s", adb_last_error());
        adb_free(out->d);
        out->d = NULL;
        return -1;
    }
    return 0;
}
static void unstage(Staged *s) {
    if (s->temp && s->d) adb_free(s->d);   /* stream-ordered: safe right after the launch */
    s->d = NULL;
}

int adb_host_result_to_host(const Result *result, void *dst) {
    if (!result) return -1;
    if (result->num_tuples == 0) return 0;
    if (flush_pending()) return -1;
    lock();
    DevResult *e = registry_find(result->payload);
    int32_t *d = e ? e->d_ptr : NULL;
    unlock();
    if (!d) {                               /* scalar or foreign host payload */
        size_t w = result->data_type == INT || result->data_type == FLOAT ? 4 : 8;
        memcpy(dst, result->payload, w * result->num_tuples);
        return 0;
    }
    if (adb_download(dst, d, 4 * result->num_tuples) != ADB_OK) {
        set_err("adb_download: %s", adb_last_error());
        return -1;
    }
    return 0;
}

/* ---- selects ------------------------------------------------------------------------------ */
#ifndef ADB_WITH_REFERENCE_HEADERS
/* src/index.c:180-185: the cost model is the constant `true`. */
bool should_use_index(Column *column, int low, int high) {
    (void)column; (void)low; (void)high;
    return true;
}
#endif

void log_result(Result *result) { (void)result; }        /* src/query.c:26-28: returns at once */

/* select_column_sorted_index (src/query.c:165-198) through the uploaded index. */
static Result *select_index_path(Column *column, DevColumn *c, int *low, int *high, Status *st) {
    int64_t h = 0;
    int32_t *out = NULL;
    if (dev_index(c, column)) return op_fail(st, "select_column");
    const int use_btree = !column->sorted;             /* create_index(... sorted=false) = btree */
    CK(adb_select_index_count(c->ix, use_btree, low, high, NULL, &h));
    out = alloc_i32((size_t)h);
    if (!out) goto fail;
    CK(adb_select_index_emit(c->ix, out));
    {
        Result *r = new_dev_result(out, (size_t)h);
        if (!r) return op_fail(st, "select_column");
        op_ok(st);
        return r;
    }
fail:
    if (out) adb_free(out);
    return op_fail(st, "select_column");
}

/* src/query.c:203-220: clustered or indexed columns go through the index, others scan. */
Result *select_column(Column *column, int *low, int *high, Status *ret_status) {
    t_err[0] = '\0';
    if (ensure_up()) return op_fail(ret_status, "select_column");
    DevColumn *c = dev_column(column);
    if (!c) return op_fail(ret_status, "select_column");
    if (column->clustered ||
        (column->has_index && should_use_index(column, low ? *low : 0, high ? *high : 0)))
        return select_index_path(column, c, low, high, ret_status);
    int64_t h = 0;
    int32_t *out = NULL;
    if (flush_pending()) goto fail;                 /* this select takes over the bitmap */
    CK(adb_select_count(c->d_data, (int64_t)c->rows, NULL, low, high, NULL, &h));
    out = alloc_i32((size_t)h);
    if (!out) goto fail;
    const int defer = S.lazy && h > 0;
    if (!defer) CK(adb_select_emit(NULL, 0, out));
    {
        Result *r = new_dev_result(out, (size_t)h);
        if (!r) return op_fail(ret_status, "select_column");
        if (defer) {                                /* positions are written by whoever needs them first */
            memset(&P, 0, sizeof P);
            P.sel_payload = r->payload;
            P.sel_d = out;
            P.h = (size_t)h;
            P.d_col = c->d_data;
            P.rows = c->rows;
            P.has_lo = low != NULL;
            P.has_hi = high != NULL;
            P.lo = low ? *low : 0;
            P.hi = high ? *high : 0;
            P.generation = adb_select_generation();
            P.active = 1;
        }
        op_ok(ret_status);
        return r;
    }
fail:
    if (out) adb_free(out);
    return op_fail(ret_status, "select_column");
}

/* src/query.c:38-86: predicate over a fetched value vector, emits the paired positions. */
Result *select_result(Result *column, Result *position, int *low_pointer, int *high_pointer,
                      Status *ret_status) {
    t_err[0] = '\0';
    Staged v = {0}, p = {0};
    int32_t *out = NULL;
    int64_t h = 0;
    if (ensure_up() || stage(column, &v) || stage(position, &p)) goto fail;
    if (position->num_tuples < column->num_tuples) {
        set_err("select: %zu values but only %zu positions", column->num_tuples, position->num_tuples);
        goto fail;
    }
    CK(adb_select_count(v.d, (int64_t)column->num_tuples, NULL, low_pointer, high_pointer, NULL, &h));
    out = alloc_i32((size_t)h);
    if (!out) goto fail;
    CK(adb_select_emit(p.d, 0, out));
    unstage(&v);
    unstage(&p);
    {
        Result *r = new_dev_result(out, (size_t)h);
        if (!r) return op_fail(ret_status, "select_result");
        op_ok(ret_status);
        return r;
    }
fail:
    if (out) adb_free(out);
    unstage(&v);
    unstage(&p);
    return op_fail(ret_status, "select_result");
}

/* src/query.c:450-583: query_count range selects in one pass; reads .low/.high only and
 * ignores has_low/has_high and indexes, exactly as query.c:474 does. */
Result **shared_select(SelectOperator *operators, int query_count, Column *column,
                       Status *ret_status) {
    t_err[0] = '\0';
    Result **results = NULL;
    int32_t **outs = NULL;
    int32_t *lows = NULL, *highs = NULL;
    int64_t *counts = NULL;
    int made = 0;
    if (ensure_up()) goto fail;
    if (query_count < 1 || query_count > ADB_MAX_BATCH || !operators) {
        set_err("shared_select: query_count %d outside [1, %d] (the dispatcher chunks batches "
                "to 150, src/server.c:366-371)", query_count, ADB_MAX_BATCH);
        goto fail;
    }
    DevColumn *c = dev_column(column);
    if (!c || flush_pending()) goto fail;
    results = calloc((size_t)query_count, sizeof *results);
    outs = calloc((size_t)query_count, sizeof *outs);
    lows = malloc(sizeof *lows * (size_t)query_count);
    highs = malloc(sizeof *highs * (size_t)query_count);
    counts = malloc(sizeof *counts * (size_t)query_count);
    if (!results || !outs || !lows || !highs || !counts) {
        set_err("out of host memory");
        goto fail;
    }
    for (int q = 0; q < query_count; ++q) {
        lows[q] = operators[q].low;
        highs[q] = operators[q].high;
    }
    CK(adb_shared_select_count(c->d_data, (int64_t)c->rows, lows, highs, query_count, counts));
    int64_t cap = 1;
    for (int q = 0; q < query_count; ++q) {
        outs[q] = alloc_i32((size_t)counts[q]);
        if (!outs[q]) goto fail;
        if (counts[q] > cap) cap = counts[q];
    }
    CK(adb_shared_select_emit(outs, cap));
    for (int q = 0; q < query_count; ++q) {
        int32_t *d = outs[q];
        outs[q] = NULL;                             /* ownership moves into the Result */
        results[q] = new_dev_result(d, (size_t)counts[q]);
        if (!results[q]) goto fail;
        made = q + 1;
    }
    free(outs); free(lows); free(highs); free(counts);
    op_ok(ret_status);
    return results;
fail:
    for (int q = 0; q < made; ++q) {
        adb_host_result_release(results[q]);
        free(results[q]->payload);
        free(results[q]);
    }
    if (outs)
        for (int q = 0; q < query_count; ++q)
            if (outs[q]) adb_free(outs[q]);
    free(results); free(outs); free(lows); free(highs); free(counts);
    return op_fail(ret_status, "shared_select");
}

/* ---- fetch --------------------------------------------------------------------------------- */
/* src/query.c:223-243: values[i] = column->data[position[i]]. */
Result *fetch_column(Column *column, Result *position_result, Status *ret_status) {
    t_err[0] = '\0';
    Staged p = {0};
    int32_t *out = NULL;
    if (ensure_up()) goto fail;
    DevColumn *c = dev_column(column);
    if (!c) goto fail;
    /* fetch of the pending select: nothing is launched yet -- the values are written together
     * with the positions, by the aggregate that usually follows or by the first other reader */
    if (P.active && position_result && position_result->payload == P.sel_payload && !P.fetch_payload &&
        position_result->num_tuples == P.h && c->rows >= P.rows) {
        out = alloc_i32(P.h);
        if (!out) goto fail;
        Result *r = new_dev_result(out, P.h);
        if (!r) return op_fail(ret_status, "fetch_column");
        if (P.active) {                             /* (new_dev_result may have flushed) */
            P.fetch_payload = r->payload;
            P.fetch_d = out;
            P.d_fetch_col = c->d_data;
        } else if (adb_fetch(c->d_data, P.sel_d, (int64_t)P.h, NULL, 0, out) != ADB_OK) {
            set_err("adb_fetch: %s", adb_last_error());
            adb_host_result_release(r);
            free(r->payload);
            free(r);
            return op_fail(ret_status, "fetch_column");
        }
        op_ok(ret_status);
        return r;
    }
    if (stage(position_result, &p)) goto fail;
    const size_t n = position_result->num_tuples;
    out = alloc_i32(n);
    if (!out) goto fail;
    CK(adb_fetch(c->d_data, p.d, (int64_t)n, NULL, 0, out));
    unstage(&p);
    {
        Result *r = new_dev_result(out, n);
        if (!r) return op_fail(ret_status, "fetch_column");
        op_ok(ret_status);
        return r;
    }
fail:
    if (out) adb_free(out);
    unstage(&p);
    return op_fail(ret_status, "fetch_column");
}

/* ---- aggregates ------------------------------------------------------------------------------ */
static int aggregate_result(const Result *r, adb_agg *h) {
    Staged v = {0};
    if (ensure_up()) return -1;
    /* aggregate of the pending fetch of the pending select: the chain's fused second kernel
     * writes both handles and the aggregate in one pass */
    if (P.active && r && P.fetch_payload && r->payload == P.fetch_payload && r->num_tuples == P.h &&
        r->data_type == INT && adb_select_generation() == P.generation) {
        P.active = 0;
        if (adb_select_emit_fetch_agg(P.d_fetch_col, P.sel_d, P.fetch_d, S.d_agg, h) != ADB_OK) {
            set_err("adb_select_emit_fetch_agg: %s", adb_last_error());
            return -1;
        }
        return 0;
    }
    if (stage(r, &v)) return -1;
    adb_status s = adb_aggregate(v.d, (int64_t)r->num_tuples, NULL, S.d_agg, h);
    if (s != ADB_OK) set_err("adb_aggregate: %s", adb_last_error());
    unstage(&v);
    return s == ADB_OK ? 0 : -1;
}

static Result *scalar_result(DataType t, const void *value, size_t width, Status *st, const char *what) {
    Result *r = malloc(sizeof *r);
    void *p = malloc(width);
    if (!r || !p) {
        free(r);
        free(p);
        set_err("out of host memory");
        return op_fail(st, what);
    }
    /* malloc handed out an address we still hold a device result for: the plumbing freed that
     * payload without telling us (no release hook, no interposer) -- it is dead, drop it */
    adb_host_payload_freed(p);
    memcpy(p, value, width);
    r->num_tuples = 1;
    r->data_type = t;
    r->payload = p;
    op_ok(st);
    return r;
}

/* src/query.c:306-323: long sum, then (double)sum / (double)num_tuples on the host -- the
 * same two conversions and one division, so the double is bit-identical (empty -> -nan). */
Result *average(Result *column, Status *ret_status) {
    t_err[0] = '\0';
    adb_agg a;
    if (aggregate_result(column, &a)) return op_fail(ret_status, "average");
    long s = (long)a.sum;
    double avg = (double)s / (double)column->num_tuples;
    return scalar_result(DOUBLE, &avg, sizeof avg, ret_status, "average");
}

/* src/query.c:325-354: long sum over a Result or over a whole base column. */
Result *sum(GeneralizedColumn *column, Status *ret_status) {
    t_err[0] = '\0';
    adb_agg a;
    if (!column) {
        set_err("NULL operand");
        return op_fail(ret_status, "sum");
    }
    if (column->column_type == RESULT) {
        if (aggregate_result(column->column_pointer.result, &a)) return op_fail(ret_status, "sum");
    } else {
        if (ensure_up()) return op_fail(ret_status, "sum");
        DevColumn *c = dev_column(column->column_pointer.column);
        if (!c) return op_fail(ret_status, "sum");
        if (adb_aggregate(c->d_data, (int64_t)c->rows, NULL, S.d_agg, &a) != ADB_OK) {
            set_err("adb_aggregate: %s", adb_last_error());
            return op_fail(ret_status, "sum");
        }
    }
    long s = (long)a.sum;
    return scalar_result(LONG, &s, sizeof s, ret_status, "sum");
}

/* src/query.c:392-437.  The reference seeds with payload[0]; on an empty input that is an
 * out-of-bounds read (oracle-undefined) -- here min is INT_MAX and max INT_MIN. */
Result *min(Result *column, Status *ret_status) {
    t_err[0] = '\0';
    adb_agg a;
    if (aggregate_result(column, &a)) return op_fail(ret_status, "min");
    int v = a.min;
    return scalar_result(INT, &v, sizeof v, ret_status, "min");
}
Result *max(Result *column, Status *ret_status) {
    t_err[0] = '\0';
    adb_agg a;
    if (aggregate_result(column, &a)) return op_fail(ret_status, "max");
    int v = a.max;
    return scalar_result(INT, &v, sizeof v, ret_status, "max");
}

/* ---- add / sub -------------------------------------------------------------------------------- */
/* src/query.c:356-390: length is column_one's; the reference does not check column_two's
 * length (it would read out of bounds) -- here a shorter column_two is an ERROR. */
static Result *ewise(Result *one, Result *two, int subtract, Status *st) {
    t_err[0] = '\0';
    const char *what = subtract ? "sub" : "add";
    Staged a = {0}, b = {0};
    int32_t *out = NULL;
    if (ensure_up() || stage(one, &a) || stage(two, &b)) goto fail;
    if (two->num_tuples < one->num_tuples) {
        set_err("%s: operands of %zu and %zu tuples", what, one->num_tuples, two->num_tuples);
        goto fail;
    }
    const size_t n = one->num_tuples;
    out = alloc_i32(n);
    if (!out) goto fail;
    CK((subtract ? adb_sub : adb_add)(a.d, b.d, (int64_t)n, NULL, out));
    unstage(&a);
    unstage(&b);
    {
        Result *r = new_dev_result(out, n);
        if (!r) return op_fail(st, what);
        op_ok(st);
        return r;
    }
fail:
    if (out) adb_free(out);
    unstage(&a);
    unstage(&b);
    return op_fail(st, what);
}
Result *add(Result *column_one, Result *column_two, Status *ret_status) {
    return ewise(column_one, column_two, 0, ret_status);
}
Result *sub(Result *column_one, Result *column_two, Status *ret_status) {
    return ewise(column_one, column_two, 1, ret_status);
}

/* ---- joins --------------------------------------------------------------------------------------- */
/* src/query.c:585-696.  results[0] lists side-one positions, results[1] side-two positions;
 * hash join is probe-major over side two, nested-loop outer-major over side one.  The
 * caller frees the two-element array (src/server.c:432). */
static Result **join(Result *v1, Result *p1, Result *v2, Result *p2, int nested, Status *st) {
    t_err[0] = '\0';
    const char *what = nested ? "nested_loop_join" : "hash_join";
    Staged a = {0}, b = {0}, c = {0}, d = {0};
    int32_t *o1 = NULL, *o2 = NULL;
    Result **results = NULL;
    int64_t m = 0;
    if (ensure_up() || stage(v1, &a) || stage(p1, &b) || stage(v2, &c) || stage(p2, &d)) goto fail;
    if (p1->num_tuples < v1->num_tuples || p2->num_tuples < v2->num_tuples) {
        set_err("%s: fewer positions than values", what);
        goto fail;
    }
    CK((nested ? adb_nested_loop_join_count : adb_hash_join_count)(
        a.d, b.d, (int64_t)v1->num_tuples, c.d, d.d, (int64_t)v2->num_tuples, &m));
    o1 = alloc_i32((size_t)m);
    o2 = alloc_i32((size_t)m);
    if (!o1 || !o2) goto fail;
    CK(adb_join_emit(o1, o2));
    unstage(&a); unstage(&b); unstage(&c); unstage(&d);
    results = malloc(2 * sizeof *results);
    if (!results) {
        set_err("out of host memory");
        goto fail;
    }
    results[0] = new_dev_result(o1, (size_t)m);
    o1 = NULL;
    results[1] = results[0] ? new_dev_result(o2, (size_t)m) : NULL;
    if (results[0]) o2 = NULL;
    if (!results[0] || !results[1]) {
        if (results[0]) {
            adb_host_result_release(results[0]);
            free(results[0]->payload);
            free(results[0]);
        }
        goto fail;
    }
    op_ok(st);
    return results;
fail:
    if (o1) adb_free(o1);
    if (o2) adb_free(o2);
    unstage(&a); unstage(&b); unstage(&c); unstage(&d);
    free(results);
    return op_fail(st, what);
}
Result **nested_loop_join(Result *column_one, Result *position_one, Result *column_two,
                          Result *position_two, Status *ret_status) {
    return join(column_one, position_one, column_two, position_two, 1, ret_status);
}
Result **hash_join(Result *column_one, Result *position_one, Result *column_two,
                   Result *position_two, Status *ret_status) {
    return join(column_one, position_one, column_two, position_two, 0, ret_status);
}

/* ---- print ------------------------------------------------------------------------------------------ */
#define PRINT_ON_DEVICE_MIN 4096       /* shorter results: the download + host loop is cheaper than three launches */
typedef struct Text {
    char *s;
    size_t len, cap;
} Text;
static int text_room(Text *t, size_t extra) {
    if (t->len + extra + 1 <= t->cap) return 0;
    size_t cap = t->cap ? t->cap : 64;
    while (cap < t->len + extra + 1) cap *= 2;
    char *n = realloc(t->s, cap);
    if (!n) return -1;
    t->s = n;
    t->cap = cap;
    return 0;
}
static void text_i64(Text *t, long long v) {          /* "%d" / "%ld" without the printf cost */
    char buf[24];
    int k = 0;
    unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
    do {
        buf[k++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) buf[k++] = '-';
    while (k) t->s[t->len++] = buf[--k];
}

/* src/query.c:245-304: results are rendered one after another (column-major), values of a
 * result separated by '\n', results separated by ','; ints "%d", longs "%ld", floats and
 * doubles "%.2f".  Device-resident payloads are copied to the host first.  The reference
 * sizes its buffer at 11 bytes per tuple and overruns it for wide values;
 * this one grows.  An all-empty print returns "" (the reference returns uninitialised
 * bytes). */
char *print(Result **results, int result_num, Status *ret_status) {
    t_err[0] = '\0';
    Text t = {0};
    void *host = NULL;
    if (flush_pending()) return op_fail(ret_status, "print");
    if (text_room(&t, 16)) goto oom;
    t.s[0] = '\0';
    for (int i = 0; i < result_num; ++i) {
        Result *r = results[i];
        if (i > 0) {
            if (text_room(&t, 1)) goto oom;
            t.s[t.len++] = ',';
        }
        const size_t n = r->num_tuples;
        const size_t w = r->data_type == INT || r->data_type == FLOAT ? 4 : 8;
        if (r->data_type == INT && n >= PRINT_ON_DEVICE_MIN) {
            /* a long device-resident INT result is formatted where it lives; the text, not
             * the integers, crosses PCIe and no sprintf runs */
            lock();
            DevResult *e = registry_find(r->payload);
            const int32_t *d = e ? e->d_ptr : NULL;
            unlock();
            if (d) {
                int64_t bytes = 0;
                void *d_text = NULL;
                if (adb_format_i32_count(d, (int64_t)n, &bytes) != ADB_OK) goto dev_fail;
                if (text_room(&t, (size_t)bytes)) goto oom;
                if (adb_alloc(&d_text, (size_t)bytes) != ADB_OK) goto dev_fail;
                if (adb_format_i32_emit(d_text) != ADB_OK ||
                    adb_download(t.s + t.len, d_text, (size_t)bytes) != ADB_OK) {
                    adb_free(d_text);
                    goto dev_fail;
                }
                adb_free(d_text);
                t.len += (size_t)bytes;
                continue;
            }
        }
        host = malloc(n ? n * w : 1);
        if (!host) goto oom;
        if (adb_host_result_to_host(r, host)) {
            free(host);
            free(t.s);
            return op_fail(ret_status, "print");
        }
        for (size_t k = 0; k < n; ++k) {
            if (text_room(&t, 400)) goto oom;       /* "%.2f" of a double: at most 312 chars */
            if (r->data_type == INT) text_i64(&t, ((int *)host)[k]);
            else if (r->data_type == LONG) text_i64(&t, ((long *)host)[k]);
            else if (r->data_type == FLOAT) t.len += (size_t)sprintf(t.s + t.len, "%.2f", ((float *)host)[k]);
            else t.len += (size_t)sprintf(t.s + t.len, "%.2f", ((double *)host)[k]);
            if (k != n - 1) t.s[t.len++] = '\n';
        }
        free(host);
        host = NULL;
    }
    t.s[t.len] = '\0';
    op_ok(ret_status);
    return t.s;
dev_fail:
    set_err("print: %s", adb_last_error());
    free(t.s);
    return op_fail(ret_status, "print");
oom:
    free(host);
    free(t.s);
    set_err("out of host memory");
    return op_fail(ret_status, "print");
}
