/*
 * adb_query_api.h -- the reference column store's operator API, as the engine's C host
 * side (analytical-database_b200/host/query_shim.c) implements it.
 *
 * The drop-in replaces /root/reference/src/query.c and src/multimap.c: it exports the
 * thirteen functions of src/include/query.h:20-44 with the same names, argument meaning,
 * ownership and error behaviour, so src/server.c:137-435 (the dispatcher), src/parse.c and
 * src/client_context.c link against it unchanged.
 *
 * Two ways to compile the host side:
 *   -DADB_WITH_REFERENCE_HEADERS -I<reference>/src/include
 *        the reference's own cs165_api.h / db_manager.h / query.h supply the types (this is
 *        how the drop-in server in oracle/_ref/dropin is built);
 *   default
 *        the layout-compatible declarations below stand in, so the host side builds and is
 *        tested where the reference tree is absent (the GPU box).  tests/test_host_layout.py
 *        checks every sizeof / offsetof below against the reference's headers.
 * Only the fields the operator path touches are spelled out; catalog-side pointers are
 * opaque.
 */
#ifndef ADB_QUERY_API_H
#define ADB_QUERY_API_H

#ifdef ADB_WITH_REFERENCE_HEADERS
#include "cs165_api.h"
#include "db_manager.h"
#include "query.h"
#else

#include <stdbool.h>
#include <stddef.h>

/* src/include/cs165_api.h:34-46 */
#define MAX_SIZE_NAME 64
#define HANDLE_MAX_SIZE 64
#define LONG_INT_LENGTH 10

/* src/include/cs165_api.h:58-63 -- what a Result's payload holds */
typedef enum DataType { INT, LONG, FLOAT, DOUBLE } DataType;

/* src/include/cs165_api.h:65-68 -- sorted copy of a column + the row of every value */
typedef struct ColumnIndex {
    int *values;
    size_t *positions;
} ColumnIndex;

struct Node;        /* src/include/btree.h:7-13 (stub B-tree, never populated) */
struct Histogram;   /* src/include/cs165_api.h:71-75 */

/* src/include/cs165_api.h:77-92 -- a base column: 128 bytes */
typedef struct Column {
    char name[MAX_SIZE_NAME];
    int *data;                   /* row_count valid ints (mmap'd file) */
    int fd;
    size_t row_count;
    bool sorted;
    bool clustered;
    bool has_index;
    ColumnIndex *index;
    struct Node *btree_node;
    struct Histogram *histogram;
    int max;
    int min;
} Column;

/* src/include/cs165_api.h:152-163 */
typedef enum StatusCode { OK, ERROR } StatusCode;
typedef struct Status {
    StatusCode code;
    char *error_message;
} Status;

/* src/include/cs165_api.h:179-183 -- an intermediate bound to a handle */
typedef struct Result {
    size_t num_tuples;
    DataType data_type;
    void *payload;
} Result;

/* src/include/cs165_api.h:188-206 */
typedef enum GeneralizedColumnType { RESULT, COLUMN } GeneralizedColumnType;
typedef union GeneralizedColumnPointer {
    Result *result;
    Column *column;
} GeneralizedColumnPointer;
typedef struct GeneralizedColumn {
    GeneralizedColumnType column_type;
    GeneralizedColumnPointer column_pointer;
} GeneralizedColumn;

/* src/include/db_manager.h:51-54,95-108 -- one range select as the parser fills it in;
 * shared_select reads .low / .high only (src/query.c:474) */
typedef enum SelectType { COLUMN_SELECT, RESULT_SELECT } SelectType;
struct Db;
struct Table;
struct Comparator;
typedef struct SelectOperator {
    SelectType select_type;
    char handle[HANDLE_MAX_SIZE];
    int low;
    int high;
    int has_low;
    int has_high;
    struct Db *db;
    struct Table *table;
    Column *column;
    Result *col_result;
    Result *pos_result;
    struct Comparator *comparator;
} SelectOperator;

/* ---- the operator API: src/include/query.h:20-44 ------------------------------------ */
Result *select_result(Result *column, Result *position, int *low_pointer, int *high_pointer,
                      Status *ret_status);
Result *select_column(Column *column, int *low, int *high, Status *ret_status);
Result *fetch_column(Column *column, Result *position_result, Status *ret_status);
char *print(Result **result, int result_num, Status *ret_status);
Result *average(Result *column, Status *ret_status);
Result *sum(GeneralizedColumn *column, Status *ret_status);
Result *add(Result *column_one, Result *column_two, Status *ret_status);
Result *sub(Result *column_one, Result *column_two, Status *ret_status);
Result *min(Result *column, Status *ret_status);
Result *max(Result *column, Status *ret_status);
Result **shared_select(SelectOperator *operators, int query_count, Column *column,
                       Status *ret_status);
Result **nested_loop_join(Result *column_one, Result *position_one, Result *column_two,
                          Result *position_two, Status *ret_status);
Result **hash_join(Result *column_one, Result *position_one, Result *column_two,
                   Result *position_two, Status *ret_status);
void log_result(Result *result);
/* src/include/cs165_api.h:94; defined by src/index.c:180-185 in the drop-in build and by
 * the shim itself (same constant `true`) in the standalone build */
bool should_use_index(Column *column, int low, int high);

#endif /* ADB_WITH_REFERENCE_HEADERS */

/* ---- hooks the engine adds (SURVEY.md section 8b "new hooks"); all optional ----------
 * The shim initialises the engine on first use (device ADB_DEVICE, default 0) and uploads a
 * column the first time an operator touches it, re-uploading when its data pointer or
 * row_count changed (insert_row may re-mmap, src/db_manager.c:178-186), so the unchanged
 * server needs none of these; they exist for hosts that want explicit control. */
#ifdef __cplusplus
extern "C" {
#endif
int adb_host_init(int device);                    /* call from main(), src/server.c:616 */
/* One process, `gpus` GPUs (SURVEY.md 8b `engine_init(int ngpus)`): every column is row-range
 * sharded over the GPUs, every operator of this header fans out over them and still returns
 * ONE Result (see the head of host/query_shim.c).  Devices ADB_DEVICE .. ADB_DEVICE+gpus-1,
 * wrapping around when the box has fewer (contexts then share a device).  Without an explicit
 * call the shim initialises itself on first use with ADB_GPUS (default 1) GPUs. */
int adb_host_init_multi(int gpus);
int adb_host_gpus(void);                          /* 0 before initialisation */
void adb_host_shutdown(void);                     /* call from shutdown_server(), src/server.c:40 */
int adb_host_column_upload(Column *column);       /* after load_db + build_index, src/server.c:120-125 */
int adb_host_column_adopt(Column *column, const void *d_data);   /* rows already in HBM (GPU-side load) */
/* the same with several GPUs: d_shards[g] (on GPU g) holds rows [g*shard_rows, (g+1)*shard_rows),
 * shard_rows a multiple of 32 */
int adb_host_column_adopt_shards(Column *column, const void *const *d_shards, size_t shard_rows);
void adb_host_column_invalidate(Column *column);  /* after insert_row, src/server.c:250 */
/* Index build on the engine -- what build_unclustered_index / build_clustered_index do
 * (src/index.c:105-146): cols[0 .. n_cols) are the table's columns in declaration order,
 * cols[which] the indexed one (its sorted / clustered flags say what to build).  Fills
 * cols[which]->index (plain malloc, as init_column_index does), permutes the siblings' host data
 * in place for a clustered index, and installs the device-side index.  Ties come out in
 * ascending row order (stable radix sort; the reference's quicksort leaves another order,
 * SURVEY.md A3 -- identical for unique keys).  host/index_shim.c builds build_index(Db*) on it. */
int adb_host_index_build(Column **cols, int n_cols, int which);
/* Updates and deletes of milestone 5: what a parser branch for
 * `relational_update(db.tbl.col, positions, value)` / `relational_delete(db.tbl, positions)` would
 * call (the reference's parser has none, src/parse.c:876-960; semantics =
 * project_tests/data_generation_scripts/milestone5.py:123-262).  cols[0 .. n_cols) are the table's
 * columns; the host arrays (Column.data, row_count) are brought up to date, unclustered indexes
 * rebuilt on the engine; a table with a clustered index is refused.  0 on success. */
int adb_host_relational_update(Column **cols, int n_cols, int which, Result *positions, int value);
int adb_host_relational_delete(Column **cols, int n_cols, Result *positions);
/* build_histogram's counts (src/index.c:63-84) computed on the device */
int adb_host_column_histogram(Column *column, int bin_size, unsigned long counts[100]);
/* Device-resident results: Result.payload of a position list / value vector is a small
 * malloc'd descriptor, so the plumbing's free(payload) (src/client_context.c:35,82) stays
 * valid; the HBM buffer behind it is returned to the engine by adb_host_result_release()
 * (the two-line patch) or, with no patch at all, by the free() interposer
 * in host/free_interpose.c. */
void adb_host_result_release(Result *result);
void adb_host_payload_freed(void *payload);       /* what the interposer calls */
/* release + free(payload) + free(Result) for n Results at once (what free_client_context does
 * per handle, src/client_context.c:76-90, with the release hook applied) */
void adb_host_results_drop(Result **results, int n);
/* Copy a result's tuples to host memory (what print does); returns 0 on success. */
int adb_host_result_to_host(const Result *result, void *dst);
const char *adb_host_last_error(void);
long adb_host_live_device_results(void);          /* diagnostics: descriptors not yet released */
void adb_host_profile_dump(void);                 /* ADB_SHIM_PROFILE=1: wall time per host phase, to stderr */
#ifdef __cplusplus
}
#endif

#endif /* ADB_QUERY_API_H */
