/*
 * Sequential CPU rulebook construction — plain-C restatement of what spconv 2.1's CPU path
 * (`ops.get_indice_pairs` with a host hash map; third-party, not in /root/reference) does for the call
 * sites spconv_backbone.py:89,92-93,113,563-564.  TEST INFRASTRUCTURE: the oracle's fast rulebook and the
 * "CPU indexing" half of the CPU baseline bench.py times.  Same canonical form as oracle/rulebook.py
 * (which it is tested against): neighbour tables, lowest row wins on duplicate coordinates, centre offset
 * of a submanifold conv is the identity, regular-conv outputs sorted by linear index.
 *
 * build: gcc -O2 -shared -fPIC -o librulebook_ref.so rulebook_ref.c   (oracle/cbuild.py)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 3

typedef struct {
    int64_t* keys;
    int32_t* vals;
    uint64_t mask;
} table_t;

static uint64_t mix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

static int table_init(table_t* t, int64_t n) {
    uint64_t s = 1024;
    while (s < (uint64_t)(2 * n + 2)) s <<= 1;
    t->keys = (int64_t*)malloc(s * sizeof(int64_t));
    t->vals = (int32_t*)malloc(s * sizeof(int32_t));
    if (!t->keys || !t->vals) return -1;
    memset(t->keys, 0xFF, s * sizeof(int64_t)); /* -1 = empty */
    t->mask = s - 1;
    return 0;
}
static void table_free(table_t* t) { free(t->keys); free(t->vals); }

/* insert-if-absent (first insertion wins); returns the stored value */
static int32_t table_put(table_t* t, int64_t key, int32_t val) {
    uint64_t s = mix((uint64_t)key) & t->mask;
    for (;;) {
        if (t->keys[s] == -1) { t->keys[s] = key; t->vals[s] = val; return val; }
        if (t->keys[s] == key) return t->vals[s];
        s = (s + 1) & t->mask;
    }
}
static int32_t table_get(const table_t* t, int64_t key) {
    uint64_t s = mix((uint64_t)key) & t->mask;
    for (;;) {
        if (t->keys[s] == -1) return -1;
        if (t->keys[s] == key) return t->vals[s];
        s = (s + 1) & t->mask;
    }
}

static void decode_k(int ndim, const int32_t* ksize, int k, int* off) {
    for (int d = ndim - 1; d >= 0; --d) { off[d] = k % ksize[d]; k /= ksize[d]; }
}

/* nbr [K, n] */
int ref_subm_rulebook(const int32_t* idx, int n, int ndim, const int32_t* shape, const int32_t* ksize,
                      const int32_t* dil, int32_t* nbr) {
    table_t t;
    if (table_init(&t, n)) return -1;
    int K = 1, centre = 0;
    for (int d = 0; d < ndim; ++d) { K *= ksize[d]; }
    for (int d = 0; d < ndim; ++d) centre = centre * ksize[d] + ksize[d] / 2;
    for (int r = 0; r < n; ++r) {
        const int32_t* p = idx + (size_t)r * (1 + ndim);
        int64_t key = p[0];
        for (int d = 0; d < ndim; ++d) key = key * shape[d] + p[1 + d];
        table_put(&t, key, r);
    }
    for (int k = 0; k < K; ++k) {
        int off[MAXD];
        decode_k(ndim, ksize, k, off);
        int32_t* row = nbr + (size_t)k * n;
        for (int r = 0; r < n; ++r) {
            if (k == centre) { row[r] = r; continue; }
            const int32_t* p = idx + (size_t)r * (1 + ndim);
            int64_t key = p[0];
            int ok = 1;
            for (int d = 0; d < ndim; ++d) {
                int v = p[1 + d] + (off[d] - ksize[d] / 2) * dil[d];
                if (v < 0 || v >= shape[d]) { ok = 0; break; }
                key = key * shape[d] + v;
            }
            row[r] = ok ? table_get(&t, key) : -1;
        }
    }
    table_free(&t);
    return 0;
}

static int cmp_i64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

static int64_t out_cell(int ndim, const int32_t* p, const int* off, const int32_t* oshape, const int32_t* stride,
                        const int32_t* pad, const int32_t* dil) {
    int64_t lin = p[0];
    for (int d = 0; d < ndim; ++d) {
        int num = p[1 + d] + pad[d] - off[d] * dil[d];
        if (num < 0 || num % stride[d]) return -1;
        int o = num / stride[d];
        if (o >= oshape[d]) return -1;
        lin = lin * oshape[d] + o;
    }
    return lin;
}

/* Phase 1: returns the number of output rows and keeps the sorted unique cells in *cells_out (caller frees
 * through ref_free).  Phase 2 fills the tables. */
int64_t ref_conv_rulebook_count(const int32_t* idx, int n, int ndim, const int32_t* oshape, const int32_t* ksize,
                                const int32_t* stride, const int32_t* pad, const int32_t* dil, int64_t** cells_out) {
    int K = 1;
    for (int d = 0; d < ndim; ++d) K *= ksize[d];
    table_t seen;
    if (table_init(&seen, (int64_t)n * 8 + 16)) return -1;
    int64_t cap = (int64_t)n * 8 + 16, m = 0;
    int64_t* cells = (int64_t*)malloc(cap * sizeof(int64_t));
    for (int r = 0; r < n; ++r) {
        const int32_t* p = idx + (size_t)r * (1 + ndim);
        for (int k = 0; k < K; ++k) {
            int off[MAXD];
            decode_k(ndim, ksize, k, off);
            int64_t lin = out_cell(ndim, p, off, oshape, stride, pad, dil);
            if (lin < 0) continue;
            table_put(&seen, lin, 1);
        }
    }
    /* collect the keys of the set */
    for (uint64_t s = 0; s <= seen.mask; ++s)
        if (seen.keys[s] != -1) {
            if (m == cap) { cap *= 2; cells = (int64_t*)realloc(cells, cap * sizeof(int64_t)); }
            cells[m++] = seen.keys[s];
        }
    table_free(&seen);
    qsort(cells, (size_t)m, sizeof(int64_t), cmp_i64);
    *cells_out = cells;
    return m;
}

int ref_conv_rulebook_fill(const int32_t* idx, int n, int ndim, const int32_t* oshape, const int32_t* ksize,
                           const int32_t* stride, const int32_t* pad, const int32_t* dil, const int64_t* cells,
                           int64_t m, int32_t* out_idx, int32_t* nbr_fwd, int32_t* nbr_bwd) {
    int K = 1;
    for (int d = 0; d < ndim; ++d) K *= ksize[d];
    table_t rank;
    if (table_init(&rank, m)) return -1;
    for (int64_t i = 0; i < m; ++i) {
        table_put(&rank, cells[i], (int32_t)i);
        int64_t lin = cells[i];
        int32_t* o = out_idx + (size_t)i * (1 + ndim);
        for (int d = ndim - 1; d >= 0; --d) { o[1 + d] = (int32_t)(lin % oshape[d]); lin /= oshape[d]; }
        o[0] = (int32_t)lin;
    }
    memset(nbr_fwd, 0xFF, (size_t)K * m * sizeof(int32_t));
    for (int k = 0; k < K; ++k) {
        int off[MAXD];
        decode_k(ndim, ksize, k, off);
        for (int r = 0; r < n; ++r) {
            const int32_t* p = idx + (size_t)r * (1 + ndim);
            int64_t lin = out_cell(ndim, p, off, oshape, stride, pad, dil);
            int32_t orow = lin < 0 ? -1 : table_get(&rank, lin);
            nbr_bwd[(size_t)k * n + r] = orow;
            if (orow >= 0) nbr_fwd[(size_t)k * m + orow] = r;
        }
    }
    table_free(&rank);
    return 0;
}

void ref_free(void* p) { free(p); }
