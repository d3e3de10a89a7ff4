// Plan executor: one C-ABI call runs a whole backbone forward (or backward) — sm_100a.
//
// The reference drives its backbone layer by layer from Python (spconv_backbone.py:609-699 VirConvL8x.forward, :207-229
// NRConvBlock.forward, :339-535 VirConv8x.forward): ~100 operator calls and ~220 kernel launches per training step,
// which on a B200 costs more host time (5.0 ms) than the kernels take to run (3.9 ms; profiles/step_anatomy_r1.txt).
// Here the host side of that loop is native: Python describes the layer graph ONCE as a small op list (the "plan"),
// and vc_exec_forward / vc_exec_backward walk it, carving every activation, rulebook and scratch buffer out of one
// caller-owned arena and enqueueing the same kernels the per-operator entry points launch.
//
//   * two streams: index ops (rulebooks, voxel->pixel projection; they depend on coordinates only) run on the SIDE
//     stream and run ahead of the feature ops (conv+BN+ReLU, concat) on the MAIN stream; the four data-dependent row
//     counts of the strided convs are read back on the side stream, so the host never waits for feature kernels;
//   * cross-stream dependencies are CUDA events from a small library-owned pool; the arena is bump-allocated and never
//     reused inside a step, so there are no memory hazards between the streams;
//   * all tensor-core weight images of the step are produced by one launch (tc_prep_images);
//   * backward walks the list in reverse with gradient accumulation rules (first contribution writes, later ones add).
// The state a forward leaves behind for its backward (buffer addresses, row counts) is a plain host struct in a
// caller-owned blob; the library keeps nothing between calls except the event pool.
#include <cuda_bf16.h>

#include <new>

#include "common.cuh"

namespace vc {
namespace {

constexpr int MAX_OPS = 256, MAX_F = 160, MAX_I = 64, MAX_RB = 64, MAX_L = 96, OPI = 24, OPF = 8;
constexpr uint64_t STATE_MAGIC = 0x5643455845433031ULL;  // "VCEXEC01"
enum { OP_SUBM_RB = 1, OP_CONV_RB = 2, OP_INDEX2UV = 3, OP_CBR = 4, OP_CAT = 5 };
enum { G_EMPTY = 0, G_EXT = 1, G_OWN = 2 };
// op int fields
enum { F_KIND = 0, F_STREAM = 1, F_A = 2, F_B = 3, F_C = 4, F_NDIM = 5, F_KS = 6, F_ST = 9, F_PD = 12, F_DL = 15, F_CIN = 18,
       F_COUT = 19, F_LAYER = 20, F_X0 = 21, F_X1 = 22, F_X2 = 23 };
// layer pointer table columns
enum { P_W = 0, P_GAMMA, P_BETA, P_RM, P_RV, P_NBT, P_DW, P_DGAMMA, P_DBETA, P_COLS };

struct ISet {
    int32_t* idx;
    const int* n_dev;      // static mode: the row count lives on the device and `n` is the capacity of the buffers
    int n, ndim, shape[3];
    int ev;  // event to wait on when consumed from the other stream (-1: none), stream that produced it
    int prod;
};
struct FSlot {
    float* f32;
    void* bf16;
    float* grad;
    const int* n_dev;      // static mode (see ISet)
    int rows, c, grad_state, ev, prod;
};
struct RBk {
    int32_t *nbr, *nbr_bwd, *pair_num;
    const int *n_in_dev, *n_out_dev;   // static mode (see ISet); table pitches are n_out (nbr) and n_in (nbr_bwd)
    int K, n_in, n_out, subm, unique, ev, prod;
};
constexpr int CTR_PER_LAYER = 8;   // forward, dgrad, up to 6 wgrad passes (wgrad2_passes: at most 4 today)

struct Layer {
    float *x, *y, *stats;
    double* sums;          // [4*cout]: forward sums, backward sums
    int* tile_ctr;         // [CTR_PER_LAYER] tile-scheduler counters of the persistent kernels (forward, dgrad, wgrad passes), zeroed with the sums
    float* wscratch;       // [K][cin][cout] fp32 accumulator of the persistent wgrad kernel (zeroed with the sums), or NULL
    void *wimg_fwd, *wimg_dgrad;
    int in_slot, out_slot, rb, use_tc, use_tc_w, cin, cout, need_dgrad;   // use_tc: conv fwd / gather dgrad; use_tc_w: wgrad / scatter dgrad
};
struct State {
    uint64_t magic;
    size_t used;           // arena bytes in use after the last call
    int n_ops, n_layers, training, precision, batch_size, is_static;
    ISet iset[MAX_I];
    FSlot f[MAX_F];
    RBk rb[MAX_RB];
    Layer layer[MAX_L];
};

struct Arena {
    char* base;
    size_t cap, used;
    bool failed;
    void* alloc(size_t bytes) {
        size_t a = (used + 255) & ~(size_t)255;
        if (a + bytes > cap) {
            failed = true;
            used = a + bytes;   // keep counting: the error message reports what would have been needed so far
            return nullptr;
        }
        used = a + bytes;
        return base + a;
    }
};

// library-owned event pool (cross-stream ordering only; timing disabled).  An event may be re-recorded as soon as the
// wait on its previous recording has been ENQUEUED, so a small round-robin pool is enough.
constexpr int EV_POOL = 512;   // > the events of one forward / backward call: the cursor restarts at every call
// (process-wide, not thread_local: autograd runs backward on its own thread; one process drives one device)
cudaEvent_t g_ev[EV_POOL];
int g_ev_dev = -1, g_ev_next = 0;

int ev_init() {
    int dev = 0;
    VC_CUDA(cudaGetDevice(&dev));
    if (g_ev_dev == dev) return VC_OK;
    if (g_ev_dev >= 0)
        for (int i = 0; i < EV_POOL; ++i) cudaEventDestroy(g_ev[i]);
    for (int i = 0; i < EV_POOL; ++i) VC_CUDA(cudaEventCreateWithFlags(&g_ev[i], cudaEventDisableTiming));
    g_ev_dev = dev;
    g_ev_next = 0;
    return VC_OK;
}

// optional per-kernel timing (bench.py's roofline pass): event pairs around the conv launches
struct TimeRec {
    cudaEvent_t a, b;
    int kind, layer;   // kind: 0 conv fwd f32, 1 conv fwd tc, 2 dgrad f32, 3 dgrad tc, 4 dgrad scatter, 5 wgrad f32, 6 wgrad tc, 7 scatter tc
};
constexpr int T_MAX = 4096;
TimeRec g_t[T_MAX];
int g_t_n = 0, g_t_created = 0;
bool g_timing = false;

struct Timed {
    int idx;
    cudaStream_t st;
    Timed(int kind, int layer, cudaStream_t s) : idx(-1), st(s) {
        if (!g_timing || g_t_n >= T_MAX) return;
        if (g_t_n >= g_t_created) {
            if (cudaEventCreate(&g_t[g_t_n].a) != cudaSuccess || cudaEventCreate(&g_t[g_t_n].b) != cudaSuccess) return;
            g_t_created = g_t_n + 1;
        }
        idx = g_t_n++;
        g_t[idx].kind = kind;
        g_t[idx].layer = layer;
        cudaEventRecord(g_t[idx].a, st);
    }
    ~Timed() {
        if (idx >= 0) cudaEventRecord(g_t[idx].b, st);
    }
};

__global__ void __launch_bounds__(256) add_kernel(const float4* a, const float4* b, float4* c, size_t n4) {   // c may alias a or b
    pdl_wait();
    pdl_launch_dependents();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 x = a[i], y = b[i];
        c[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
}

// backward of the channel concat: da = dcat[:, :ca] (+ a_add), db = dcat[:, ca:] (+ b_add); add pointers may alias
// the outputs (in-place accumulation) or be NULL
__global__ void __launch_bounds__(256) cat2_bwd_kernel(const float4* __restrict__ dcat, int n, int ca4, int cb4,
                                                       float4* da, const float4* a_add, float4* db, const float4* b_add) {
    pdl_wait();
    pdl_launch_dependents();
    const int c4 = ca4 + cb4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * c4, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const size_t r = i / c4;
        const int c = (int)(i % c4);
        float4 v = dcat[i];
        if (c < ca4) {
            const size_t o = r * ca4 + c;
            if (a_add) { float4 w = a_add[o]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
            da[o] = v;
        } else {
            const size_t o = r * cb4 + (c - ca4);
            if (b_add) { float4 w = b_add[o]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
            db[o] = v;
        }
    }
}

int ew_grid(size_t n4) {
    size_t b = (n4 + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    return (int)(b < 1 ? 1 : b);
}

int launch_add(const float* a, const float* b, float* c, size_t n, cudaStream_t st) {
    if (n == 0) return VC_OK;
    VC_LAUNCH_CHAIN(add_kernel, dim3(ew_grid(n / 4)), dim3(256), 0, st, (const float4*)a, (const float4*)b, (float4*)c, n / 4);
    return VC_OK;
}

bool tc_ok(int c) { return c == 16 || c == 32 || c == 64; }

struct Ctx {
    const int32_t* oi;
    const float* of;
    int n_ops;
    const uint64_t* lp;
    const float* lf;
    int n_layers;
    State* S;
    Arena A;
    cudaStream_t st[3];   // forward: main, index stream, second index stream (image branch); backward: main, wgrad
    int32_t* err;
    const int* op(int i) const { return oi + (size_t)i * OPI; }
    template <class T>
    T* P(int layer, int col) const { return reinterpret_cast<T*>(lp[(size_t)layer * P_COLS + col]); }
};

#define VC_TRY(expr)              \
    do {                          \
        int rc__ = (expr);        \
        if (rc__) return rc__;    \
    } while (0)

#define VC_ALLOC(var, type, bytes)                                                                     \
    type var = (type)C.A.alloc(bytes);                                                                 \
    if (!var && (bytes) > 0) {                                                                         \
        set_error("plan executor: arena too small (%zu bytes given, > %zu needed)", C.A.cap, C.A.used); \
        return VC_ERR_WORKSPACE;                                                                       \
    }

// make `consumer` stream wait for a resource produced on the other stream
int wait_for(Ctx& C, int& ev, int prod, int consumer) {
    if (ev >= 0 && C.st[prod] != C.st[consumer]) VC_CUDA(cudaStreamWaitEvent(C.st[consumer], g_ev[ev], 0));
    return VC_OK;
}
int record_on(Ctx& C, int s) {   // -> event id recorded at the current tail of stream s
    if (C.st[0] == C.st[1] && C.st[0] == C.st[2]) return -1;
    int e = g_ev_next;
    g_ev_next = (g_ev_next + 1) % EV_POOL;
    if (cudaEventRecord(g_ev[e], C.st[s]) != cudaSuccess) return -1;
    return e;
}

int check_plan(const int32_t* oi, int n_ops, int n_layers) {
    VC_CHECK_ARG(oi && n_ops > 0 && n_ops <= MAX_OPS && n_layers >= 0 && n_layers <= MAX_L, "bad plan size");
    for (int i = 0; i < n_ops; ++i) {
        const int* o = oi + (size_t)i * OPI;
        VC_CHECK_ARG(o[F_KIND] >= OP_SUBM_RB && o[F_KIND] <= OP_CAT, "op %d: unknown kind %d", i, o[F_KIND]);
        VC_CHECK_ARG(o[F_STREAM] >= 0 && o[F_STREAM] <= 2, "op %d: bad stream", i);
        const int lim_a = (o[F_KIND] == OP_CBR || o[F_KIND] == OP_CAT) ? MAX_F : MAX_I;
        VC_CHECK_ARG(o[F_A] >= 0 && o[F_A] < lim_a, "op %d: slot a out of range", i);
        if (o[F_KIND] == OP_CBR) {
            VC_CHECK_ARG(o[F_B] >= 0 && o[F_B] < MAX_F && o[F_C] >= 0 && o[F_C] < MAX_RB, "op %d: bad slots", i);
            VC_CHECK_ARG(o[F_LAYER] >= 0 && o[F_LAYER] < n_layers, "op %d: bad layer", i);
        } else if (o[F_KIND] == OP_CAT) {
            VC_CHECK_ARG(o[F_B] >= 0 && o[F_B] < MAX_F && o[F_C] >= 0 && o[F_C] < MAX_F, "op %d: bad slots", i);
        } else if (o[F_KIND] == OP_SUBM_RB) {
            VC_CHECK_ARG(o[F_C] >= 0 && o[F_C] < MAX_RB, "op %d: bad rulebook id", i);
        } else if (o[F_KIND] == OP_CONV_RB) {
            VC_CHECK_ARG(o[F_B] >= 0 && o[F_B] < MAX_I && o[F_C] >= 0 && o[F_C] < MAX_RB, "op %d: bad ids", i);
        } else {
            VC_CHECK_ARG(o[F_B] >= 0 && o[F_B] < MAX_I, "op %d: bad index-set id", i);
        }
    }
    return VC_OK;
}

}  // namespace
}  // namespace vc

using namespace vc;

extern "C" size_t vc_exec_state_bytes(void) { return sizeof(State); }

extern "C" int vc_exec_timing(int enable) {
    g_timing = enable != 0;
    g_t_n = 0;
    return VC_OK;
}

// ms_out[i], kind_layer_out[2*i..] for the records since vc_exec_timing(1); synchronises the device.  Returns the count.
extern "C" int vc_exec_timing_read(float* ms_out, int32_t* kind_layer_out, int max_records) {
    if (cudaDeviceSynchronize() != cudaSuccess) return VC_ERR_CUDA;
    int n = g_t_n < max_records ? g_t_n : max_records;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_t[i].a, g_t[i].b);
        ms_out[i] = ms;
        kind_layer_out[2 * i] = g_t[i].kind;
        kind_layer_out[2 * i + 1] = g_t[i].layer;
    }
    return n;
}

extern "C" int vc_exec_query(const void* state, int what, int id, long long* out) {
    const State* S = (const State*)state;
    VC_CHECK_ARG(S && S->magic == STATE_MAGIC && out, "not an executor state");
    switch (what) {
        case 0:
            out[0] = (long long)S->used;
            return VC_OK;
        case 1: {
            VC_CHECK_ARG(id >= 0 && id < MAX_F, "slot id");
            const FSlot& f = S->f[id];
            out[0] = (long long)(uintptr_t)f.f32; out[1] = (long long)(uintptr_t)f.bf16; out[2] = f.rows; out[3] = f.c;
            return VC_OK;
        }
        case 2: {
            VC_CHECK_ARG(id >= 0 && id < MAX_I, "index-set id");
            const ISet& s = S->iset[id];
            out[0] = (long long)(uintptr_t)s.idx; out[1] = s.n; out[2] = s.ndim; out[3] = s.shape[0]; out[4] = s.shape[1];
            out[5] = s.shape[2]; out[6] = (long long)(uintptr_t)s.n_dev;
            return VC_OK;
        }
        case 3: {
            VC_CHECK_ARG(id >= 0 && id < MAX_RB, "rulebook id");
            const RBk& r = S->rb[id];
            out[0] = (long long)(uintptr_t)r.nbr; out[1] = (long long)(uintptr_t)r.nbr_bwd;
            out[2] = (long long)(uintptr_t)r.pair_num; out[3] = r.K; out[4] = r.n_in; out[5] = r.n_out; out[6] = r.subm;
            out[7] = r.unique;
            return VC_OK;
        }
        case 4: {
            VC_CHECK_ARG(id >= 0 && id < MAX_L, "layer id");
            const Layer& l = S->layer[id];
            out[0] = (long long)(uintptr_t)l.x; out[1] = (long long)(uintptr_t)l.y; out[2] = (long long)(uintptr_t)l.stats;
            out[3] = l.use_tc; out[4] = l.rb; out[5] = l.in_slot; out[6] = l.out_slot;
            return VC_OK;
        }
    }
    set_error("vc_exec_query: unknown selector %d", what);
    return VC_ERR_INVALID;
}

extern "C" int vc_exec_forward(const int32_t* ops_i, const float* ops_f, int n_ops, const uint64_t* layer_ptrs,
                               const float* layer_f, int n_layers, const float* feats0, int c0, const int32_t* idx0, int n0,
                               const int32_t* shape0, int batch_size, const float* proj_params, int training, int precision,
                               int want_pair_num, void* arena, size_t arena_bytes, int32_t* pinned_host, int32_t* err_flag,
                               void* state, size_t state_bytes, vc_stream_t main_stream, vc_stream_t side_stream,
                               int side_waits_main, const int32_t* caps, const int32_t* n0_dev, int32_t* overflow_flag,
                               vc_stream_t side2_stream) {
    // Static mode (n0_dev != NULL; CUDA-graph capturable: no host read of a device value, every buffer sized from host-side
    // capacities): n0 is the capacity of the input buffers, *n0_dev the number of valid rows, caps[i] the row capacity of
    // index set i (the strided convs' outputs); every data-dependent row count stays in device memory, kernels clamp to the
    // capacities and report an overflow through *overflow_flag (the largest row count that did not fit).
    const bool is_static = n0_dev != nullptr;
    VC_CHECK_ARG(!is_static || (caps && overflow_flag), "static mode needs capacities and an overflow flag");
    // Phased execution (static mode only; bits 8-9 of want_pair_num): 1 = enqueue ONLY the index operators (voxel -> rulebooks,
    // projection), 2 = ONLY the feature operators (weight images, conv + BN + ReLU, concat).  Both phases walk the whole plan with
    // the same deterministic bump allocation, so they agree on every address and leave the same state blob: graph.GraphedStep
    // captures them as two graphs and runs the index graph of step t+1 beside the feature graph of step t.
    const int phase = (want_pair_num >> 8) & 3;
    want_pair_num &= 0xff;
    VC_CHECK_ARG(phase == 0 || (is_static && phase <= 2), "phased execution needs static mode");
    const bool do_idx = phase != 2, do_feat = phase != 1;
    VC_TRY(check_plan(ops_i, n_ops, n_layers));
    VC_CHECK_ARG(state && state_bytes >= sizeof(State), "state blob too small (%zu < %zu)", state_bytes, sizeof(State));
    VC_CHECK_ARG(feats0 && idx0 && n0 > 0 && shape0 && batch_size > 0 && arena && pinned_host && ops_f && layer_ptrs && layer_f,
                 "null / empty argument");
    VC_CHECK_ARG(precision == 0 || precision == 1, "precision must be 0 (fp32) or 1 (bf16)");
    VC_TRY(ev_init());
    State* S = new (state) State();
    memset(S, 0, sizeof(State));
    S->n_ops = n_ops; S->n_layers = n_layers; S->training = training; S->precision = precision; S->batch_size = batch_size;
    S->is_static = is_static;
    for (int i = 0; i < MAX_I; ++i) S->iset[i].ev = -1;
    for (int i = 0; i < MAX_F; ++i) S->f[i].ev = -1;
    for (int i = 0; i < MAX_RB; ++i) S->rb[i].ev = -1;

    Ctx C;
    C.oi = ops_i; C.of = ops_f; C.n_ops = n_ops; C.lp = layer_ptrs; C.lf = layer_f; C.n_layers = n_layers; C.S = S;
    C.A = Arena{(char*)arena, arena_bytes, 0, false};
    C.st[0] = (cudaStream_t)main_stream;
    C.st[1] = side_stream ? (cudaStream_t)side_stream : (cudaStream_t)main_stream;
    C.st[2] = (side_stream && side2_stream) ? (cudaStream_t)side2_stream : C.st[1];   // image-branch index ops (index2uv + 2-D rulebooks)
    C.err = err_flag;
    const bool two = C.st[0] != C.st[1];
    g_ev_next = 0;

    S->iset[0].idx = const_cast<int32_t*>(idx0); S->iset[0].n = n0; S->iset[0].ndim = 3; S->iset[0].n_dev = n0_dev;
    for (int d = 0; d < 3; ++d) S->iset[0].shape[d] = shape0[d];
    S->f[0].f32 = const_cast<float*>(feats0); S->f[0].rows = n0; S->f[0].c = c0; S->f[0].n_dev = n0_dev;

    // rulebook meta known from the plan alone (needed to choose the dgrad weight images)
    int rb_subm[MAX_RB] = {0}, rb_unique[MAX_RB] = {0}, rb_K[MAX_RB] = {0};
    for (int i = 0; i < n_ops; ++i) {
        const int* o = C.op(i);
        if (o[F_KIND] == OP_SUBM_RB || o[F_KIND] == OP_CONV_RB) {
            int K = 1;
            for (int d = 0; d < o[F_NDIM]; ++d) K *= o[F_KS + d];
            rb_K[o[F_C]] = K;
            rb_subm[o[F_C]] = o[F_KIND] == OP_SUBM_RB;
            rb_unique[o[F_C]] = o[F_KIND] == OP_CONV_RB ? 1 : o[F_X0];
        }
    }

    // BatchNorm accumulators + tile-scheduler counters (+ the wgrad accumulators of a training step) of every layer: one
    // region, one memset
    size_t sums_doubles = 0, n_cbr = 0, wg_floats = 0;
    for (int i = 0; i < n_ops; ++i)
        if (C.op(i)[F_KIND] == OP_CBR) {
            const int* o = C.op(i);
            sums_doubles += 4 * (size_t)o[F_COUT];
            ++n_cbr;
            if (training && precision == 1 && g_tc_variant == 1 && tc2_ch_ok(o[F_CIN]) && tc2_ch_ok(o[F_COUT]))
                wg_floats += ((size_t)rb_K[o[F_C]] * o[F_CIN] * o[F_COUT] + 63) / 64 * 64;
        }
    const size_t zero_bytes = sums_doubles * 8 + n_cbr * CTR_PER_LAYER * sizeof(int) + wg_floats * 4;
    VC_ALLOC(sums_all, double*, zero_bytes);
    int* ctr_all = reinterpret_cast<int*>(sums_all + sums_doubles);
    float* wg_all = reinterpret_cast<float*>(ctr_all + n_cbr * CTR_PER_LAYER);
    if (zero_bytes && do_feat) VC_CUDA(cudaMemsetAsync(sums_all, 0, zero_bytes, C.st[0]));
    size_t wg_cur = 0;
    size_t sums_cur = 0, ctr_cur = 0;

    // tensor-core weight images of every layer, one launch
    for (int i = 0; i < n_ops; ++i) {
        const int* o = C.op(i);
        if (o[F_KIND] != OP_CBR) continue;
        Layer& L = S->layer[o[F_LAYER]];
        L.cin = o[F_CIN]; L.cout = o[F_COUT]; L.in_slot = o[F_A]; L.out_slot = o[F_B]; L.rb = o[F_C];
        L.need_dgrad = o[F_X0];
        L.use_tc = precision == 1 && tc_conv_ch_ok(L.cin) && tc_conv_ch_ok(L.cout);
        L.use_tc_w = precision == 1 && tc_ok(L.cin) && tc_ok(L.cout);
        L.sums = sums_all + sums_cur;
        sums_cur += 4 * (size_t)L.cout;
        L.tile_ctr = ctr_all + CTR_PER_LAYER * ctr_cur++;
        L.wscratch = nullptr;
        if (training && precision == 1 && g_tc_variant == 1 && tc2_ch_ok(L.cin) && tc2_ch_ok(L.cout)) {
            L.wscratch = wg_all + wg_cur;
            wg_cur += ((size_t)rb_K[L.rb] * L.cin * L.cout + 63) / 64 * 64;
        }
    }
    if (precision == 1) {
        TcPrepTable T;
        T.n = 0;
        for (int i = 0; i < n_ops; ++i) {
            const int* o = C.op(i);
            if (o[F_KIND] != OP_CBR) continue;
            Layer& L = S->layer[o[F_LAYER]];
            if (!L.use_tc) continue;
            const int K = rb_K[L.rb];
            // dgrad image: mirrored for a submanifold table used as its own transpose, plain for a strided conv's
            // nbr_bwd table and for the many-to-one image-branch table (tensor-core scatter, round-1 image layout)
            const bool many_to_one = rb_subm[L.rb] && !rb_unique[L.rb];
            const bool dgrad_tc = training && L.need_dgrad && (!many_to_one || L.use_tc_w);
            for (int mode = 0; mode < (dgrad_tc ? 2 : 1); ++mode) {
                const int layout = (mode == 1 && many_to_one) ? 0 : g_tc_variant;
                VC_ALLOC(img, void*, tc_image_bytes(L.cin, L.cout, K, layout));
                (mode == 0 ? L.wimg_fwd : L.wimg_dgrad) = img;
                if (T.n == TC_PREP_MAX) {
                    if (do_feat) VC_TRY(tc_prep_images(T, C.st[0]));
                    T.n = 0;
                }
                TcPrepEntry& e = T.e[T.n++];
                e.w = C.P<const float>(o[F_LAYER], P_W); e.img = img; e.cin = L.cin; e.cout = L.cout; e.K = K; e.mode = mode;
                e.mirror = mode == 1 && rb_subm[L.rb] && !many_to_one;
                e.layout = layout;
            }
        }
        if (do_feat) VC_TRY(tc_prep_images(T, C.st[0]));
    }

    // the side stream starts after everything already queued on main (the caller's inputs) — unless the caller vouches
    // that idx0 / proj_params are already valid for the side stream and the arena is safe to write from it
    // (side_waits_main = 0): the index pipeline of this step then overlaps whatever main is still running (the
    // previous step's backward)
    if (two && side_waits_main) {
        int e = record_on(C, 0);
        if (e >= 0) {
            VC_CUDA(cudaStreamWaitEvent(C.st[1], g_ev[e], 0));
            if (C.st[2] != C.st[1]) VC_CUDA(cudaStreamWaitEvent(C.st[2], g_ev[e], 0));
        }
    }
    int last_ev[3] = {-1, -1, -1};     // latest event of each side stream (joined into main at the end)
    auto side_ev = [&](int s_) -> int {
        if (s_ == 0) return -1;
        return last_ev[s_] = record_on(C, s_);
    };
    int n_syncs = 0;
    static thread_local void* chain_ws[MAX_OPS];
    static thread_local size_t chain_wsb[MAX_OPS];
    static thread_local bool chain_done[MAX_OPS];
    for (int i = 0; i < n_ops; ++i) chain_done[i] = false;
    // count + row-count read-back + output indices of strided conv op j (its input index set must exist)
    auto count_emit = [&](int j) -> int {
        const int* o = C.op(j);
        const int s = two ? o[F_STREAM] : 0;
        cudaStream_t st = C.st[s];
        ISet& I = S->iset[o[F_A]];
        ISet& O = S->iset[o[F_B]];
        RBk& R = S->rb[o[F_C]];
        VC_CHECK_ARG(I.ndim == o[F_NDIM], "op %d: index set %d has the wrong ndim", j, o[F_A]);
        VC_TRY(wait_for(C, I.ev, I.prod, s));
        int32_t oshape[3] = {0, 0, 0};
        VC_TRY(vc_conv_out_shape(I.ndim, I.shape, o + F_KS, o + F_ST, o + F_PD, o + F_DL, oshape));
        const size_t wsb = vc_conv_rulebook_ws_bytes(I.ndim, batch_size, oshape);
        VC_ALLOC(ws, void*, wsb);
        VC_ALLOC(n_dev, int32_t*, 4);
        int n_out;
        if (is_static) {
            n_out = caps[o[F_B]];
            VC_CHECK_ARG(n_out > 0, "op %d: static mode needs a capacity for index set %d", j, o[F_B]);
            if (do_idx)
                VC_TRY(conv_rulebook_count_dev(I.idx, I.n, I.n_dev, I.ndim, batch_size, I.shape, o + F_KS, o + F_ST, o + F_PD, o + F_DL,
                                               n_dev, n_out, overflow_flag, ws, wsb, st));
        } else {
            VC_TRY(conv_rulebook_count_dev(I.idx, I.n, nullptr, I.ndim, batch_size, I.shape, o + F_KS, o + F_ST, o + F_PD, o + F_DL,
                                           n_dev, 0, nullptr, ws, wsb, st));
            VC_CUDA(cudaMemcpyAsync(pinned_host + (n_syncs & 7), n_dev, 4, cudaMemcpyDeviceToHost, st));
            // the tensor-core kernels' pipeline-timeout flag rides along with the first read-back of a step (it reports on the
            // kernels of EARLIER steps: a wedged mbarrier wait must not go unnoticed in production — ADVICE r1)
            if (n_syncs == 0 && C.err != nullptr) VC_CUDA(cudaMemcpyAsync(pinned_host + 15, C.err, 4, cudaMemcpyDeviceToHost, st));
            VC_CUDA(cudaStreamSynchronize(st));        // the one data-dependent size per strided conv
            n_out = pinned_host[n_syncs & 7];
            if (n_syncs == 0 && C.err != nullptr && pinned_host[15] != 0) {
                set_error("a tensor-core kernel's pipeline wait timed out in an earlier step (error flag %d): its results are invalid",
                          pinned_host[15]);
                return VC_ERR_PIPELINE;
            }
            ++n_syncs;
        }
        R.K = rb_K[o[F_C]]; R.n_in = I.n; R.n_out = n_out; R.subm = 0; R.unique = 1;
        R.n_in_dev = I.n_dev; R.n_out_dev = is_static ? n_dev : nullptr;
        O.n = n_out; O.ndim = I.ndim; O.n_dev = R.n_out_dev;
        for (int d = 0; d < 3; ++d) O.shape[d] = oshape[d];
        VC_ALLOC(oidx, int32_t*, (size_t)(n_out > 0 ? n_out : 1) * (1 + I.ndim) * 4);
        O.idx = oidx;
        if (is_static && do_idx) VC_CUDA(cudaMemsetAsync(oidx, 0xFF, (size_t)n_out * (1 + I.ndim) * 4, st));   // defined tail (-1) for the published indices
        if (do_idx)
            VC_TRY(conv_rulebook_fill_phases(I.idx, I.n, I.n_dev, I.ndim, batch_size, I.shape, o + F_KS, o + F_ST, o + F_PD, o + F_DL, n_out, O.idx,
                                             nullptr, nullptr, nullptr, ws, wsb, st, 1));
        O.prod = s;
        O.ev = side_ev(s);
        chain_ws[j] = ws; chain_wsb[j] = wsb; chain_done[j] = true;
        return VC_OK;
    };

    for (int i = 0; i < n_ops; ++i) {
        const int* o = C.op(i);
        const float* fo = C.of + (size_t)i * OPF;
        const int s = two ? o[F_STREAM] : 0;
        cudaStream_t st = C.st[s];
        switch (o[F_KIND]) {
            case OP_SUBM_RB: {
                ISet& I = S->iset[o[F_A]];
                RBk& R = S->rb[o[F_C]];
                VC_CHECK_ARG(I.idx && I.ndim == o[F_NDIM], "op %d: index set %d not built / wrong ndim", i, o[F_A]);
                VC_TRY(wait_for(C, I.ev, I.prod, s));
                R.K = rb_K[o[F_C]]; R.n_in = R.n_out = I.n; R.subm = 1; R.unique = o[F_X0];
                R.n_in_dev = R.n_out_dev = I.n_dev;
                VC_ALLOC(nbr, int32_t*, (size_t)R.K * I.n * 4);
                R.nbr = nbr;
                if (want_pair_num) {
                    VC_ALLOC(pn, int32_t*, (size_t)R.K * 4);
                    R.pair_num = pn;
                }
                const size_t wsb = vc_subm_rulebook_ws_bytes(I.n);
                VC_ALLOC(ws, void*, wsb);
                if (do_idx) VC_TRY(subm_rulebook_dev(I.idx, I.n, I.n_dev, I.ndim, batch_size, I.shape, o + F_KS, o + F_DL, R.nbr, R.pair_num, ws, wsb, st));
                R.prod = s;
                R.ev = side_ev(s);
                break;
            }
            case OP_CONV_RB: {
                // First strided conv reached: run the count -> (host reads the row count) -> emit-indices chain of EVERY
                // strided conv of the plan now, back to back.  Each link only needs the previous link's output indices,
                // and it is the only part of the forward the host has to wait for; the neighbour tables, submanifold
                // rulebooks and projections queue up behind it without further synchronisation.
                if (!chain_done[i]) {
                    if (is_static) {
                        // no host read to wait for: keep the natural order (count -> indices -> tables per stage), so the
                        // tables of an early stage are not queued behind the counting kernels of all later ones
                        VC_TRY(count_emit(i));
                    } else {
                        for (int j = i; j < n_ops; ++j)
                            if (C.op(j)[F_KIND] == OP_CONV_RB && S->iset[C.op(j)[F_A]].idx) VC_TRY(count_emit(j));
                    }
                }
                VC_CHECK_ARG(chain_done[i], "op %d: index set %d not built", i, o[F_A]);
                ISet& I = S->iset[o[F_A]];
                ISet& O = S->iset[o[F_B]];
                RBk& R = S->rb[o[F_C]];
                VC_TRY(wait_for(C, I.ev, I.prod, s));
                VC_TRY(wait_for(C, O.ev, O.prod, s));
                VC_ALLOC(nbr, int32_t*, (size_t)R.K * (R.n_out > 0 ? R.n_out : 1) * 4);
                VC_ALLOC(nbr_bwd, int32_t*, (size_t)R.K * I.n * 4);
                R.nbr = nbr; R.nbr_bwd = nbr_bwd;
                if (want_pair_num) {
                    VC_ALLOC(pn, int32_t*, (size_t)R.K * 4);
                    R.pair_num = pn;
                }
                if (do_idx)
                    VC_TRY(conv_rulebook_fill_phases(I.idx, I.n, I.n_dev, I.ndim, batch_size, I.shape, o + F_KS, o + F_ST, o + F_PD, o + F_DL,
                                                     R.n_out, O.idx, R.nbr, R.nbr_bwd, R.pair_num, chain_ws[i], chain_wsb[i], st, 2));
                R.prod = s;
                R.ev = side_ev(s);
                break;
            }
            case OP_INDEX2UV: {
                ISet& I = S->iset[o[F_A]];
                ISet& O = S->iset[o[F_B]];
                VC_CHECK_ARG(I.idx && I.ndim == 3 && proj_params, "op %d: index2uv needs a built 3-D index set and projection params", i);
                VC_TRY(wait_for(C, I.ev, I.prod, s));
                VC_ALLOC(uv, int32_t*, (size_t)I.n * 3 * 4);
                O.idx = uv; O.n = I.n; O.ndim = 2; O.shape[0] = o[F_KS]; O.shape[1] = o[F_KS + 1]; O.shape[2] = 0; O.n_dev = I.n_dev;
                if (do_idx) VC_TRY(index2uv_dev(I.idx, I.n, I.n_dev, batch_size, proj_params, fo, o[F_X0], o[F_X1], o[F_X2], uv, st));
                O.prod = s;
                O.ev = side_ev(s);
                break;
            }
            case OP_CBR: {
                const int li = o[F_LAYER];
                Layer& L = S->layer[li];
                FSlot& X = S->f[L.in_slot];
                FSlot& Y = S->f[L.out_slot];
                RBk& R = S->rb[L.rb];
                VC_CHECK_ARG(X.f32 && R.nbr && X.c == L.cin && X.rows == R.n_in, "op %d: input slot / rulebook mismatch (rows %d vs %d, c %d vs %d)",
                             i, X.rows, R.n_in, X.c, L.cin);
                VC_TRY(wait_for(C, R.ev, R.prod, s));
                VC_TRY(wait_for(C, X.ev, X.prod, s));
                const size_t elems = (size_t)R.n_out * L.cout;
                VC_ALLOC(x, float*, elems * 4);
                VC_ALLOC(y, float*, elems * 4);
                VC_ALLOC(stats, float*, (size_t)4 * L.cout * 4);
                void* yb = nullptr;
                if (precision == 1) {
                    VC_ALLOC(yb_, void*, elems * 2);
                    yb = yb_;
                }
                L.x = x; L.y = y; L.stats = stats;
                double* sums = training ? L.sums : nullptr;
                if (L.use_tc && R.n_out > 0) {
                    if (!X.bf16) {   // no producer wrote a shadow (network input): cast once
                        VC_ALLOC(xb, void*, (size_t)X.rows * X.c * 2);
                        if (do_feat) VC_TRY(vc_cast_f32_bf16(X.f32, xb, (long long)X.rows * X.c, st));
                        X.bf16 = xb;
                    }
                    Timed t(1, li, st);
                    // (flag bit 1: the rulebook was built on another stream and joined by an event, so the kernel may load its first
                    //  table slices before its griddepcontrol.wait — conv_tc2.cu)
                    if (do_feat)
                        VC_TRY(tc_conv_with_image(L.cin, L.cout, X.bf16, L.wimg_fwd, R.nbr, R.n_out, x, R.n_out, R.n_out_dev, R.K, sums, C.err, st,
                                                  nullptr, L.tile_ctr, (phase == 2 || C.st[R.prod] != C.st[s]) ? 2 : 0));
                } else {
                    const size_t wsb = vc_conv_ws_bytes(L.cin, L.cout, R.K);
                    VC_ALLOC(ws, void*, wsb);
                    Timed t(0, li, st);
                    if (do_feat) VC_TRY(vc_conv_fwd_f32(X.f32, C.P<const float>(li, P_W), R.nbr, x, R.n_out, L.cin, L.cout, R.K, sums, ws, wsb, st));
                }
                if (do_feat)
                    VC_TRY(bn_apply_relu_dev(x, L.sums, R.n_out, R.n_out_dev, L.cout, C.P<const float>(li, P_GAMMA), C.P<const float>(li, P_BETA),
                                             C.P<float>(li, P_RM), C.P<float>(li, P_RV), C.P<long long>(li, P_NBT), C.lf[2 * li + 1],
                                             C.lf[2 * li], training, y, L.cout, yb, L.cout, stats, 1, is_static, st));
                Y.f32 = y; Y.bf16 = yb; Y.rows = R.n_out; Y.c = L.cout; Y.prod = s; Y.n_dev = R.n_out_dev;
                Y.ev = side_ev(s);
                break;
            }
            case OP_CAT: {
                FSlot& Aa = S->f[o[F_A]];
                FSlot& Bb = S->f[o[F_B]];
                FSlot& O = S->f[o[F_C]];
                VC_CHECK_ARG(Aa.f32 && Bb.f32 && Aa.rows == Bb.rows, "op %d: concat inputs not built / row mismatch", i);
                VC_TRY(wait_for(C, Aa.ev, Aa.prod, s));
                VC_TRY(wait_for(C, Bb.ev, Bb.prod, s));
                const size_t elems = (size_t)Aa.rows * (Aa.c + Bb.c);
                VC_ALLOC(out, float*, elems * 4);
                void* ob = nullptr;
                if (precision == 1) {
                    VC_ALLOC(ob_, void*, elems * 2);
                    ob = ob_;
                }
                if (do_feat) VC_TRY(cat2_dev(Aa.f32, Bb.f32, out, ob, Aa.rows, Aa.n_dev, Aa.c, Bb.c, st));
                O.f32 = out; O.bf16 = ob; O.rows = Aa.rows; O.c = Aa.c + Bb.c; O.prod = s; O.n_dev = Aa.n_dev;
                O.ev = side_ev(s);
                break;
            }
        }
    }
    for (int s_ = 1; s_ <= 2; ++s_)
        if (two && last_ev[s_] >= 0) VC_CUDA(cudaStreamWaitEvent(C.st[0], g_ev[last_ev[s_]], 0));   // join
    S->used = C.A.used;
    S->magic = STATE_MAGIC;
    return VC_OK;
}

namespace vc {
namespace {

// hand a gradient contribution of shape [rows, c] to slot F: `write(dst)` must enqueue kernels that OVERWRITE dst
template <class W>
int contribute(Ctx& C, FSlot& F, cudaStream_t st, W&& write) {
    const size_t n = (size_t)F.rows * F.c;
    if (F.grad_state == G_EMPTY) {
        VC_ALLOC(buf, float*, n * 4);
        VC_TRY(write(buf));
        F.grad = buf;
    } else if (F.grad_state == G_EXT) {
        VC_ALLOC(buf, float*, n * 4);
        VC_TRY(write(buf));
        VC_TRY(launch_add(buf, F.grad, buf, n, st));
        F.grad = buf;
    } else {
        VC_ALLOC(tmp, float*, n * 4);
        VC_TRY(write(tmp));
        VC_TRY(launch_add(F.grad, tmp, F.grad, n, st));
    }
    F.grad_state = G_OWN;
    return VC_OK;
}

// contribution by a kernel that can ADD an existing [rows, c] matrix in its epilogue (the tensor-core gather dgrad):
// `write(dst, addend)` overwrites dst with result + addend (addend may be NULL or alias dst) — no add kernel, no temporary
template <class W>
int contribute_fused(Ctx& C, FSlot& F, W&& write) {
    const size_t n = (size_t)F.rows * F.c;
    if (F.grad_state == G_OWN) {
        VC_TRY(write(F.grad, (const float*)F.grad));
    } else {
        VC_ALLOC(buf, float*, n * 4);
        VC_TRY(write(buf, F.grad_state == G_EXT ? (const float*)F.grad : nullptr));
        F.grad = buf;
    }
    F.grad_state = G_OWN;
    return VC_OK;
}

// contribution by a kernel that ACCUMULATES with atomics (the scatter dgrads): `scatter(dst)` adds into dst, which must
// hold the sum so far — zeros, a copy of the outside gradient, or the slot's own buffer (then nothing is staged at all)
template <class W>
int contribute_scatter(Ctx& C, FSlot& F, cudaStream_t st, W&& scatter) {
    const size_t n = (size_t)F.rows * F.c;
    if (F.grad_state != G_OWN) {
        VC_ALLOC(buf, float*, n * 4);
        if (F.grad_state == G_EXT)
            VC_CUDA(cudaMemcpyAsync(buf, F.grad, n * 4, cudaMemcpyDeviceToDevice, st));
        else
            VC_CUDA(cudaMemsetAsync(buf, 0, n * 4, st));
        F.grad = buf;
        F.grad_state = G_OWN;
    }
    return scatter(F.grad);
}

}  // namespace
}  // namespace vc

extern "C" int vc_exec_backward(const int32_t* ops_i, const float* ops_f, int n_ops, const uint64_t* layer_ptrs,
                                const float* layer_f, int n_layers, const int32_t* pub_slots, const uint64_t* ext_grads,
                                int n_pub, void* arena, size_t arena_bytes, int32_t* err_flag, void* state,
                                vc_stream_t stream_, vc_stream_t wgrad_stream_) {
    VC_TRY(check_plan(ops_i, n_ops, n_layers));
    State* S = (State*)state;
    VC_CHECK_ARG(S && S->magic == STATE_MAGIC && S->n_ops == n_ops && S->n_layers == n_layers, "state does not belong to this plan");
    VC_CHECK_ARG(arena && layer_ptrs && (n_pub == 0 || (pub_slots && ext_grads)), "null argument");
    Ctx C;
    C.oi = ops_i; C.of = ops_f; C.n_ops = n_ops; C.lp = layer_ptrs; C.lf = layer_f; C.n_layers = n_layers; C.S = S;
    C.A = Arena{(char*)arena, arena_bytes, S->used, false};
    C.st[0] = (cudaStream_t)stream_;
    C.st[1] = wgrad_stream_ ? (cudaStream_t)wgrad_stream_ : (cudaStream_t)stream_;
    C.st[2] = C.st[1];
    C.err = err_flag;
    g_ev_next = 0;
    cudaStream_t st = C.st[0];
    cudaStream_t wst = C.st[1];      // weight gradients: off the dgrad chain's critical path
    const bool two = st != wst;
    if (two) VC_TRY(ev_init());
    int last_w_ev = -1;
    const int training = S->training;
    WgradFinTable fin;          // persistent wgrad kernels accumulate into scratch images; ONE transposing launch at the end
    fin.n = 0;

    for (int i = 0; i < MAX_F; ++i) { S->f[i].grad = nullptr; S->f[i].grad_state = G_EMPTY; }
    for (int i = 0; i < n_pub; ++i) {
        VC_CHECK_ARG(pub_slots[i] > 0 && pub_slots[i] < MAX_F, "published slot id");
        if (ext_grads[i]) {
            FSlot& F = S->f[pub_slots[i]];
            VC_CHECK_ARG(F.grad_state == G_EMPTY, "slot %d published twice", pub_slots[i]);
            F.grad = reinterpret_cast<float*>(ext_grads[i]);
            F.grad_state = G_EXT;
        }
    }

    for (int i = n_ops - 1; i >= 0; --i) {
        const int* o = C.op(i);
        if (o[F_KIND] == OP_CAT) {
            FSlot& Aa = S->f[o[F_A]];
            FSlot& Bb = S->f[o[F_B]];
            FSlot& O = S->f[o[F_C]];
            if (O.grad_state == G_EMPTY) continue;
            float *da, *db;
            const float *a_add = nullptr, *b_add = nullptr;
            if (Aa.grad_state == G_OWN) { da = Aa.grad; a_add = Aa.grad; }
            else { VC_ALLOC(t, float*, (size_t)Aa.rows * Aa.c * 4); da = t; a_add = Aa.grad_state == G_EXT ? Aa.grad : nullptr; }
            if (Bb.grad_state == G_OWN) { db = Bb.grad; b_add = Bb.grad; }
            else { VC_ALLOC(t, float*, (size_t)Bb.rows * Bb.c * 4); db = t; b_add = Bb.grad_state == G_EXT ? Bb.grad : nullptr; }
            if (O.rows > 0) {
                VC_LAUNCH_CHAIN(cat2_bwd_kernel, dim3(ew_grid((size_t)O.rows * O.c / 4)), dim3(256), 0, st, (const float4*)O.grad,
                                O.rows, Aa.c / 4, Bb.c / 4, (float4*)da, (const float4*)a_add, (float4*)db, (const float4*)b_add);
            }
            Aa.grad = da; Aa.grad_state = G_OWN;
            Bb.grad = db; Bb.grad_state = G_OWN;
            continue;
        }
        if (o[F_KIND] != OP_CBR) continue;
        const int li = o[F_LAYER];
        Layer& L = S->layer[li];
        FSlot& X = S->f[L.in_slot];
        FSlot& Y = S->f[L.out_slot];
        RBk& R = S->rb[L.rb];
        float* dw = C.P<float>(li, P_DW);
        float* dgamma = C.P<float>(li, P_DGAMMA);
        float* dbeta = C.P<float>(li, P_DBETA);
        const size_t wn = (size_t)R.K * L.cin * L.cout;
        if (Y.grad_state == G_EMPTY || R.n_out == 0) {   // nothing flowed back into this layer
            VC_CUDA(cudaMemsetAsync(dw, 0, wn * 4, st));
            VC_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)L.cout * 4, st));
            VC_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)L.cout * 4, st));
            continue;
        }
        const size_t elems = (size_t)R.n_out * L.cout;
        VC_ALLOC(dx, float*, elems * 4);
        void* dxb = nullptr;
        if (S->precision == 1) {
            VC_ALLOC(t, void*, elems * 2);
            dxb = t;
        }
        VC_TRY(bn_relu_bwd_dev(Y.grad, L.cout, L.x, C.P<const float>(li, P_GAMMA), L.stats, dx, dxb, dgamma, dbeta, R.n_out, R.n_out_dev,
                               L.cout, training, L.sums + 2 * L.cout, S->is_static, st));
        // wgrad (needs dx / its bf16 shadow: ordered after the BN backward through an event when on its own stream)
        if (two) {
            int e = record_on(C, 0);
            if (e >= 0) VC_CUDA(cudaStreamWaitEvent(wst, g_ev[e], 0));
        }
        if (L.wscratch != nullptr) {
            Timed t(6, li, wst);
            VC_TRY((g_wgrad_variant ? tc3_wgrad : tc2_wgrad)(L.cin, L.cout, X.bf16, dxb, R.nbr, R.n_out, L.wscratch, R.n_out, R.n_out_dev,
                                                              R.K, C.err, wst, L.tile_ctr + 2));
            if (fin.n == WGRAD_FIN_MAX) {
                VC_TRY(wgrad_finalize(fin, wst));
                fin.n = 0;
            }
            WgradFinEntry& fe = fin.e[fin.n++];
            fe.scratch = L.wscratch; fe.dw = dw; fe.cin = L.cin; fe.cout = L.cout; fe.K = R.K; fe.first = 0;
        } else if (L.use_tc_w) {
            const size_t wsb = vc_conv_wgrad_tc_ws_bytes(R.n_out, L.cin, L.cout, R.K);
            VC_ALLOC(ws, void*, wsb);
            Timed t(6, li, wst);
            VC_TRY(vc_conv_wgrad_tc(X.bf16, dxb, R.nbr, dw, R.n_out, L.cin, L.cout, R.K, ws, wsb, C.err, wst));
        } else {
            const size_t wsb = vc_conv_wgrad_ws_bytes(R.n_out, L.cin, L.cout, R.K);
            VC_ALLOC(ws, void*, wsb);
            Timed t(5, li, wst);
            VC_TRY(vc_conv_wgrad_f32(X.f32, dx, R.nbr, dw, R.n_out, L.cin, L.cout, R.K, ws, wsb, wst));
        }
        if (two) last_w_ev = record_on(C, 1);
        // dgrad
        if (!L.need_dgrad || L.in_slot == 0) continue;
        if (R.subm && !R.unique && L.use_tc_w && L.wimg_dgrad) {
            VC_TRY(contribute_scatter(C, X, st, [&](float* dst) -> int {
                Timed t(7, li, st);
                return tc_scatter_with_image(L.cout, L.cin, dxb, L.wimg_dgrad, R.nbr, dst, R.n_out, R.K, C.err, st);
            }));
        } else if (R.subm && !R.unique) {
            const size_t wsb = vc_conv_ws_bytes(L.cin, L.cout, R.K);
            VC_ALLOC(ws, void*, wsb);
            VC_TRY(contribute_scatter(C, X, st, [&](float* dst) -> int {
                Timed t(4, li, st);
                return vc_conv_dgrad_scatter_f32(dx, C.P<const float>(li, P_W), R.nbr, dst, R.n_out, L.cin, L.cout, R.K, ws, wsb, st);
            }));
        } else if (L.use_tc && L.wimg_dgrad) {
            const int32_t* table = R.subm ? R.nbr : R.nbr_bwd;
            VC_TRY(contribute_fused(C, X, [&](float* dst, const float* addend) -> int {
                Timed t(3, li, st);
                return tc_conv_with_image(L.cout, L.cin, dxb, L.wimg_dgrad, table, R.n_in, dst, R.n_in, R.n_in_dev, R.K, nullptr, C.err, st,
                                          addend, L.tile_ctr + 1, (R.subm ? 0 : 1) | 2);      // (tables from the forward call: early loads allowed)
            }));
        } else {
            const size_t wsb = vc_conv_ws_bytes(L.cin, L.cout, R.K);
            VC_ALLOC(ws, void*, wsb);
            const int32_t* table = R.subm ? R.nbr : R.nbr_bwd;
            VC_TRY(contribute(C, X, st, [&](float* dst) -> int {
                Timed t(2, li, st);
                return vc_conv_dgrad_f32(dx, C.P<const float>(li, P_W), table, dst, R.n_in, L.cin, L.cout, R.K, R.subm ? 1 : 0, ws, wsb,
                                         st);
            }));
        }
    }
    if (fin.n > 0) {
        VC_TRY(wgrad_finalize(fin, wst));
        if (two) last_w_ev = record_on(C, 1);
    }
    if (two && last_w_ev >= 0) VC_CUDA(cudaStreamWaitEvent(st, g_ev[last_w_ev], 0));   // join: gradients complete on `stream`
    S->used = C.A.used;
    return VC_OK;
}
