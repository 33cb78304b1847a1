// vspectra_api.cu -- C-ABI of the weight-spectrum path (see include/vspectra.h).
//
// Host side of the drop-in boundary: validates shapes, buckets the ragged batch by
// shape class (a ViT has three classes, a six-scenario sweep nine), lays
// out the device workspace and launches the three stages per class on the caller's
// stream.  No global device state; the only process-wide datum is the launch counter.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "eig_kernels.cuh"
#include "gram_f64.cuh"
#include "gram_i8.cuh"
#include "dgemm_dmma.cuh"

using namespace vsp;

namespace {

std::atomic<int64_t> g_launches{0};
thread_local std::string t_cuda_error;

inline bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    t_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define VSP_CUDA(call)                                 \
    do {                                               \
        if (!cuda_ok((call), #call)) return VSP_E_CUDA; \
    } while (0)

// Development aid (VSP_KERNEL_TIMING=1): one CUDA event after every launch of an execution, per-launch times
// printed to stderr after a stream synchronisation.  Off by default: no events, no synchronisation.
struct LaunchTimer {
    bool on = false;
    std::vector<std::pair<std::string, cudaEvent_t>> ev;
    void start(cudaStream_t st) {
        static const bool enabled = std::getenv("VSP_KERNEL_TIMING") != nullptr;
        on = enabled;
        tick("begin", st);
    }
    void tick(const char* name, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        ev.emplace_back(name, e);
    }
    void report(cudaStream_t st) {
        if (!on) return;
        cudaStreamSynchronize(st);
        for (auto& pe : ev) cudaEventSynchronize(pe.second);
        std::string line = "[vsp timing]";
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
            char buf[96];
            std::snprintf(buf, sizeof buf, " %s=%.3f", ev[i].first.c_str(), ms);
            line += buf;
        }
        std::fprintf(stderr, "%s\n", line.c_str());
        for (auto& pe : ev) cudaEventDestroy(pe.second);
        ev.clear();
    }
};
thread_local LaunchTimer t_timer;

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int64_t round_up64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

struct ShapeClass {
    int n;       // order of the Gram matrix
    int full;    // 1: eigensolve in global memory
    int begin;   // first item (sorted order)
    int count;
    int npad, split;
    int refine_slots;             // FP64 pool of the ill-conditioned re-solve
    int64_t refine_slot_doubles;  // max K*n over the class
    int64_t refine_off;           // byte offset of the pool in the refine region
    int64_t refine_items_off;     // byte offset of the slot -> item table
    int refine_counter;           // index of this class's slot counter
    int refine_B;                 // CTAs per cluster of the re-solve (0: one-CTA kernel)
    bool refine_shared = false;   // re-solve entry point that leaves room for bisection CTAs on its SMs (n <= 256, large classes)
    mutable int refine_launch_slots = 0;  // clusters per launch: refine_slots capped by what is resident at once (first execution)
    int refine_kmax;              // largest contraction length of the class
    int refine_xs_cap;            // doubles of shared memory for a CTA's share of X
};

int validate(int32_t count, const int32_t* rows, const int32_t* cols, const int64_t* ld) {
    if (count < 0) return VSP_E_ARG;
    if (count > 0 && (!rows || !cols)) return VSP_E_ARG;
    for (int i = 0; i < count; ++i) {
        if (rows[i] < 1 || cols[i] < 1) return VSP_E_ARG;
        if (ld && ld[i] < cols[i]) return VSP_E_ARG;
        if (std::min(rows[i], cols[i]) > VSP_MAX_N) return VSP_E_UNSUPPORTED;
    }
    return VSP_OK;
}

}  // namespace

struct vsp_plan {
    int count = 0;
    int dtype = VSP_F32;
    vsp_opts opts{};
    std::vector<ItemDesc> items;  // sorted by shape class
    std::vector<int> order;       // sorted position -> caller index
    std::vector<ShapeClass> classes;
    std::vector<I8Class> i8classes;  // fp32 inputs: (n, Kp) groups sharing one digit-plane tensor
    int64_t i8_bytes = 0;            // digit planes + row exponents, after the FP64 region
    int64_t refine_bytes = 0;        // slot counters (first 1 KB) + rounding flags + FP64 pools, after the int8 region
    int64_t refine_zero_bytes = 0;   // head of that region that is zeroed before every execution
    int gram_method = 1;             // 1: tcgen05 int8 split (fp32 inputs), 0: FP64 CUDA cores
    int64_t ws_doubles = 0;
    int64_t sv_total = 0;
    ItemDesc* d_items = nullptr;
    int device = -1;
    int sm_count = 148;              // persistent kernels: one CTA per SM
    // the re-solve of ill-conditioned items runs beside the bisection kernel on a forked stream
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<PFN_encodeTiled>(f);
    }();
    return fn;
}

// digit planes of one Gram class as a 2-D uint8 tensor [count*6*n rows][kp bytes], 64-byte swizzle
int make_plane_map(CUtensorMap* map, void* base, const I8Class& c, int box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        t_cuda_error = "cuTensorMapEncodeTiled entry point not available";
        return VSP_E_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)c.kp, (cuuint64_t)c.count * kDigits * c.n};
    const cuuint64_t strides[1] = {(cuuint64_t)c.kp};
    const cuuint32_t box[2] = {(cuuint32_t)kI8ChunkK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        t_cuda_error = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
        return VSP_E_CUDA;
    }
    return VSP_OK;
}

// stage 1 on the tensor cores for every (n, Kp) group of this shape class: slices, then the persistent MMA kernel.
// (Running a later group's slices on a side stream beside an earlier group's MMAs was measured and gains nothing:
// the MMA CTA fills its SM)
int launch_gram_i8(const vsp_plan* p, const ShapeClass& c, double* ws, unsigned char* wsb, int* inexact, cudaStream_t st) {
    for (const I8Class& g : p->i8classes) {
        if (g.n != c.n) continue;
        dim3 sgrid(g.count, (g.n + 31) / 32);
        slice_i8_kernel<<<sgrid, 256, 0, st>>>(p->d_items, g, wsb, inexact);
        g_launches++;
        t_timer.tick("slice_i8", st);
        if (!cuda_ok(cudaGetLastError(), "slice_i8_kernel")) return VSP_E_CUDA;
        CUtensorMap tmA, tmB;
        int rc = make_plane_map(&tmA, wsb + g.slice_off, g, kI8TileM);
        if (rc == VSP_OK) rc = make_plane_map(&tmB, wsb + g.slice_off, g, kI8TileN);
        if (rc != VSP_OK) return rc;
        if (!cuda_ok(cudaFuncSetAttribute(gram_i8_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SmemBytes),
                     "cudaFuncSetAttribute(gram_i8_mma_kernel)"))
            return VSP_E_CUDA;
        static const int mma_grid = std::getenv("VSP_GRAM_GRID") ? std::atoi(std::getenv("VSP_GRAM_GRID")) : 0;  // experiments
        const int mgrid = std::min(g.mtiles * g.count, mma_grid > 0 ? mma_grid : p->sm_count);  // persistent: a CTA per SM
        gram_i8_mma_kernel<<<mgrid, kI8Threads, kI8SmemBytes, st>>>(p->d_items, g, wsb, ws, tmA, tmB);
        g_launches++;
        t_timer.tick("gram_i8_mma", st);
        if (!cuda_ok(cudaGetLastError(), "gram_i8_mma_kernel")) return VSP_E_CUDA;
    }
    return VSP_OK;
}

template <typename TIn>
int launch_gram(const vsp_plan* p, const ShapeClass& c, double* ws, cudaStream_t st) {
    if (c.n <= 32) {
        dim3 grid(c.count, 1);
        gram_f64_kernel<TIn, 32, 32><<<grid, 256, 0, st>>>(p->d_items, c.begin, ws);
    } else {
        const int nt = (c.n + 63) / 64;
        dim3 grid(c.count, nt * (nt + 1) / 2);
        gram_f64_kernel<TIn, 64, 16><<<grid, 256, 0, st>>>(p->d_items, c.begin, ws);
    }
    g_launches++;
    return cuda_ok(cudaGetLastError(), "gram_f64_kernel") ? VSP_OK : VSP_E_CUDA;
}

}  // namespace

extern "C" {

int vsp_version(void) { return VSP_VERSION; }

const char* vsp_error_string(int code) {
    switch (code) {
        case VSP_OK: return "ok";
        case VSP_E_ARG: return "invalid argument";
        case VSP_E_UNSUPPORTED: return "unsupported shape or dtype";
        case VSP_E_WORKSPACE: return "workspace too small";
        case VSP_E_CUDA: return "CUDA runtime error";
        case VSP_E_ALLOC: return "host allocation failed";
        default: return "unknown error";
    }
}

const char* vsp_last_cuda_error(void) { return t_cuda_error.c_str(); }

int64_t vsp_kernel_launch_count(void) { return g_launches.load(); }
void vsp_reset_kernel_launch_count(void) { g_launches.store(0); }

int vsp_sv_offsets(int32_t count, const int32_t* rows, const int32_t* cols, int64_t* sv_offsets) {
    const int rc = validate(count, rows, cols, nullptr);
    if (rc != VSP_OK) return rc;
    if (!sv_offsets) return VSP_E_ARG;
    int64_t off = 0;
    for (int i = 0; i < count; ++i) {
        sv_offsets[i] = off;
        off += std::min(rows[i], cols[i]);
    }
    sv_offsets[count] = off;
    return VSP_OK;
}

// doubles of one item's Gram region: the packed triangle (or the full matrix) and, for the orders the bandwidth-8
// reduction takes (sbr8.cuh), its compact band output behind it
static bool use_sbr8() {
    static const bool on = std::getenv("VSP_NO_SBR8") == nullptr;  // experiments: round-1 kernels for every order
    return on;
}
static int gram_layout(int n) {
    if (n > kSmemMaxN) return kGramFull;
    return (use_sbr8() && sbr8_order(n) <= kSbr8MaxN) ? kGramTiled : kGramPacked;
}
static int64_t item_gram_doubles(int n, int layout) {
    if (layout == kGramFull) return round_up64((int64_t)n * n, 4);
    if (layout == kGramTiled) return sbr8_band_off(n) + round_up64(sbr8_band_doubles(n), 4);
    return round_up64(poff(n), 4);
}

static int plan_layout(vsp_plan* p, int32_t count, const int32_t* rows, const int32_t* cols, const int64_t* ld,
                       int32_t dtype, const vsp_opts* opts);
static int64_t plan_total_bytes(const vsp_plan* plan);

// Upper bound for any plan over these shapes: the layout of the fp32 plan (the fp64 plan has the same FP64 and
// re-solve regions and no digit planes), computed by the same code that lays out a real plan.
int64_t vsp_workspace_bytes(int32_t count, const int32_t* rows, const int32_t* cols) {
    const int rc = validate(count, rows, cols, nullptr);
    if (rc != VSP_OK) return rc;
    vsp_plan tmp;
    const int lrc = plan_layout(&tmp, count, rows, cols, nullptr, VSP_F32, nullptr);
    if (lrc != VSP_OK) return lrc;
    return plan_total_bytes(&tmp) + 256;  // + alignment slack inside the caller's buffer
}

// host-side part of a plan: shape classes, workspace layout (no CUDA calls)
static int plan_layout(vsp_plan* p, int32_t count, const int32_t* rows, const int32_t* cols, const int64_t* ld,
                       int32_t dtype, const vsp_opts* opts) {
    p->count = count;
    p->dtype = dtype;
    if (opts) {
        p->opts = *opts;
    } else {
        p->opts.fit_start = p->opts.fit_end = p->opts.hill_k = -1;
        p->opts.want_sv = -1;
        p->opts.refine = -1;
        p->opts.dist_k = 0;
        p->opts.clauset = 0;
    }
    if (p->opts.clauset != 1) p->opts.clauset = 0;
    if (p->opts.dist_k < 0) p->opts.dist_k = 0;
    // caller-order SV offsets
    std::vector<int64_t> sv_off(count + 1, 0);
    for (int i = 0; i < count; ++i) sv_off[i + 1] = sv_off[i] + std::min(rows[i], cols[i]);
    p->sv_total = sv_off[count];
    // bucket by Gram order
    p->order.resize(count);
    for (int i = 0; i < count; ++i) p->order[i] = i;
    std::stable_sort(p->order.begin(), p->order.end(), [&](int a, int b) {
        const int na = std::min(rows[a], cols[a]), nb = std::min(rows[b], cols[b]);
        if (na != nb) return na < nb;
        return std::max(rows[a], cols[a]) < std::max(rows[b], cols[b]);
    });
    p->items.resize(count);
    int64_t off = 0;
    for (int s = 0; s < count; ++s) {
        const int i = p->order[s];
        ItemDesc& it = p->items[s];
        std::memset(&it, 0, sizeof(it));
        it.rows = rows[i];
        it.cols = cols[i];
        it.ld = ld ? ld[i] : cols[i];
        it.n = std::min(rows[i], cols[i]);
        it.kdim = std::max(rows[i], cols[i]);
        it.trans = rows[i] > cols[i] ? 1 : 0;
        it.item = i;
        it.full = gram_layout(it.n);
        it.sv_off = sv_off[i];
        it.gram_off = off;
        off += item_gram_doubles(it.n, it.full);
        it.de_off = off;
        off += round_up64(2 * (int64_t)it.n + MISC_COUNT, 4);
        if (p->classes.empty() || p->classes.back().n != it.n) {
            ShapeClass c{};
            c.n = it.n;
            c.full = it.full;
            c.begin = s;
            c.count = 0;
            c.npad = round_up(it.n, 32);
            c.split = c.full == kGramFull ? 1 : std::max(1, std::min(4, 384 / c.npad));
            p->classes.push_back(c);
        }
        p->classes.back().count++;
    }
    p->ws_doubles = off;
    {   // tuning / fallback switch: VSP_GRAM=f64 selects the FP64 CUDA-core Gram for fp32 inputs too
        const char* e = std::getenv("VSP_GRAM");
        p->gram_method = (e && std::string(e) == "f64") ? 0 : 1;
    }
    // The exactness argument of the int8 split (gram_i8.cuh: every level sum fits int32, the Horner halves fit
    // 2^53) holds for contraction lengths up to kI8MaxK; a plan with a longer one takes the FP64 Gram kernel.
    for (int s = 0; s < count; ++s)
        if (p->items[s].kdim > kI8MaxK) p->gram_method = 0;
    if (dtype == VSP_F32 && p->gram_method == 1) {
        int64_t boff = 0;
        for (int s = 0; s < count; ++s) {
            const ItemDesc& it = p->items[s];
            const int kp = round_up(it.kdim, kI8ChunkK);
            if (p->i8classes.empty() || p->i8classes.back().n != it.n || p->i8classes.back().kp != kp) {
                I8Class c{};
                c.n = it.n;
                c.kp = kp;
                c.begin = s;
                c.count = 0;
                c.mtiles = (it.n + kI8TileM - 1) / kI8TileM;  // <= 32 for n <= VSP_MAX_N
                const int mtl = c.mtiles - 1, tail_rows = it.n - mtl * kI8TileM;
                static const bool no_xt = std::getenv("VSP_GRAM_NO_XT") != nullptr;  // experiments
                c.xt = (!no_xt && mtl >= 1 && tail_rows <= kI8TileN) ? 2 * mtl : -1;  // gram_i8.cuh: I8Class::xt
                for (int mt = 0; mt < c.mtiles; ++mt) {
                    const int last_col = std::min(it.n - 1, mt * kI8TileM + kI8TileM - 1);
                    c.nt_count[mt] = (unsigned char)(last_col / kI8TileN + 1);
                    if (c.xt >= 0) c.nt_count[mt] = (unsigned char)(mt == mtl ? 1 : 2 * mt + 3);
                }
                p->i8classes.push_back(c);
            }
            p->i8classes.back().count++;
        }
        for (I8Class& c : p->i8classes) {
            c.slice_off = boff;
            boff += round_up64((int64_t)c.count * kDigits * c.n * c.kp, 1024);
            c.exp_off = boff;
            boff += round_up64((int64_t)c.count * c.n * 4, 1024);
        }
        p->i8_bytes = boff;
    }
    if (const char* e = std::getenv("VSP_REFINE")) {  // experiments: VSP_REFINE=0 disables the re-solve
        if (std::string(e) == "0") p->opts.refine = 0;
    }
    if (p->opts.refine != 0) {
        int64_t roff = 1024;  // the slot counters live in the first KB,
        roff += round_up64((int64_t)count * 4, 1024);  // the per-item "stage 1 rounded an element" flags behind them
        p->refine_zero_bytes = roff;
        int idx = 0;
        for (ShapeClass& c : p->classes) {
            c.refine_slots = 0;
            if (c.n > kRefineMaxN || idx >= 256) continue;
            int64_t kn = 0;
            for (int s = c.begin; s < c.begin + c.count; ++s)
                kn = std::max<int64_t>(kn, (int64_t)p->items[s].kdim * p->items[s].n);
            c.refine_slot_doubles = round_up64(kn, 4);
            {   // cluster re-solve (refine_cluster.cuh): B CTAs per flagged matrix; VSP_REFINE_OLD=1 keeps the one-CTA kernel
                static const bool old_refine = std::getenv("VSP_REFINE_OLD") != nullptr;
                int kmax = 0;
                int64_t min_share = INT64_MAX;
                // n <= 256: THREE CTAs per matrix.  A 192 x 192 share is then exactly the 96 KB shared-memory cap, and the
                // co-residency limit (clusters do not span GPCs) is ~46 clusters instead of ~33: the Scenario-A sweep flags
                // 36 matrices per step, which four-CTA clusters served in two rounds (re-solve 3.6 ms, three-CTA 2.3 ms)
                // n > 256: a step sweeps the CTA's share of X in L2 three times, so the step time follows the share: sixteen
                // CTAs (non-portable cluster size, one GPC each: 8 resident clusters; n = 768: 14.8 ms per matrix) while the
                // flagged matrices of a random-init class (~4 %) fit one round, else eight (22 ms, but 25 % fewer SM-ms)
                // n <= 256, small class (a single checkpoint, the short chunks at the ends of a host sweep): the bisection
                // kernel is over in a fraction of a millisecond, so the re-solve of a flagged matrix IS the latency of the
                // call: all registers and four CTAs (1.5 instead of 2.2 ms per matrix)
                c.refine_shared = c.n <= 256 && c.count > 2048;  // (1 116 matrices: bisection 0.8 ms against 1.9 ms of shared re-solve)
                int B = c.n <= 256 ? (c.refine_shared ? 3 : 4) : (c.count <= 200 ? kRcMaxCluster : 8);
                if (const char* e = std::getenv("VSP_REFINE_B")) B = std::max(1, std::min(kRcMaxCluster, std::atoi(e)));  // experiments
                for (int s = c.begin; s < c.begin + c.count; ++s) {
                    kmax = std::max(kmax, p->items[s].kdim);
                    min_share = std::min<int64_t>(min_share, (int64_t)((c.n + B - 1) / B) * p->items[s].kdim);
                }
                c.refine_B = (old_refine || c.n < 16) ? 0 : B;
                c.refine_kmax = kmax;
                // a CTA's share of X stays in shared memory when it is small (the square matrices of ViT-Tiny: 74 KB),
                // which leaves room for bisection CTAs on the same SM; larger shares work from the L2-resident pool
                c.refine_xs_cap = (min_share <= 12288) ? 12288 : 0;
                if (const char* e = std::getenv("VSP_REFINE_XS")) c.refine_xs_cap = std::atoi(e);  // experiments
                if (c.refine_B > 0)
                    c.refine_slot_doubles = round_up64(refine_cluster_slot_doubles(kmax, c.n, c.refine_B), 4);
            }
            // pool buffers = CTAs of the re-solve launch (1024 threads: one per SM); more flagged items than
            // buffers are served in rounds, so the pool never limits how many items can be re-solved
            c.refine_slots = std::min(c.count, std::max(4, std::min(c.count / 8, 2 * 148)));
            // cluster re-solve: no more clusters than can be resident at once (each loops over the work list)
            if (c.refine_B > 0) c.refine_slots = std::min(c.refine_slots, std::max(4, 148 / c.refine_B));
            c.refine_counter = idx++;
            c.refine_items_off = roff;  // work list: one entry per item of the class
            roff += round_up64((int64_t)c.count * 4, 1024);
            c.refine_off = roff;
            roff += round_up64(c.refine_slot_doubles * 8 * c.refine_slots, 1024);
        }
        p->refine_bytes = roff;
    }
    return VSP_OK;
}

int vsp_plan_create(int32_t count, const int32_t* rows, const int32_t* cols, const int64_t* ld, int32_t dtype,
                    const vsp_opts* opts, vsp_plan** out_plan) {
    if (!out_plan) return VSP_E_ARG;
    *out_plan = nullptr;
    if (dtype != VSP_F32 && dtype != VSP_F64) return VSP_E_UNSUPPORTED;
    const int rc = validate(count, rows, cols, ld);
    if (rc != VSP_OK) return rc;
    vsp_plan* p = new (std::nothrow) vsp_plan();
    if (!p) return VSP_E_ALLOC;
    const int lrc = plan_layout(p, count, rows, cols, ld, dtype, opts);
    if (lrc != VSP_OK) {
        delete p;
        return lrc;
    }
    if (count > 0) {
        if (!cuda_ok(cudaGetDevice(&p->device), "cudaGetDevice") ||
            !cuda_ok(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, p->device), "cudaDeviceGetAttribute") ||
            !cuda_ok(cudaMalloc(&p->d_items, sizeof(ItemDesc) * (size_t)count), "cudaMalloc(items)") ||
            !cuda_ok(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking), "cudaStreamCreate(side)") ||
            !cuda_ok(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming), "cudaEventCreate") ||
            !cuda_ok(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming), "cudaEventCreate")) {
            vsp_plan_destroy(p);
            return VSP_E_CUDA;
        }
    }
    *out_plan = p;
    return VSP_OK;
}

static int64_t plan_f64_bytes(const vsp_plan* plan) { return round_up64(plan->ws_doubles * (int64_t)sizeof(double), 1024); }
static int64_t plan_total_bytes(const vsp_plan* plan) {
    return plan_f64_bytes(plan) + plan->i8_bytes + plan->refine_bytes + 2048;
}

int64_t vsp_plan_workspace_bytes(const vsp_plan* plan) { return plan ? plan_total_bytes(plan) : VSP_E_ARG; }
int64_t vsp_plan_sv_count(const vsp_plan* plan) { return plan ? plan->sv_total : VSP_E_ARG; }

void vsp_plan_destroy(vsp_plan* plan) {
    if (!plan) return;
    if (plan->d_items) cudaFree(plan->d_items);
    if (plan->side) cudaStreamDestroy(plan->side);
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    if (plan->ev_join) cudaEventDestroy(plan->ev_join);
    delete plan;
}

static int execute_impl(vsp_plan* p, const void* const* d_ptrs, double* d_sv, vsp_record* d_records,
                        void* d_workspace, int64_t workspace_bytes, void* stream, std::vector<cudaEvent_t>* evs,
                        double* d_dist = nullptr) {
    if (!p) return VSP_E_ARG;
    if (p->count == 0) return VSP_OK;
    if (!d_ptrs || !d_records || !d_workspace) return VSP_E_ARG;
    if (p->opts.want_sv != 0 && !d_sv) return VSP_E_ARG;
    if ((p->opts.dist_k > 0 || p->opts.clauset > 0) && !d_dist) return VSP_E_ARG;  // such a plan needs vsp_plan_execute_dist
    // align the workspace to 256 bytes inside the caller's buffer
    uintptr_t base = reinterpret_cast<uintptr_t>(d_workspace);
    uintptr_t aligned = (base + 255) & ~uintptr_t(255);
    if ((int64_t)(aligned - base) + plan_f64_bytes(p) + p->i8_bytes + p->refine_bytes > workspace_bytes)
        return VSP_E_WORKSPACE;
    double* ws = reinterpret_cast<double*>(aligned);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* refine_base = reinterpret_cast<unsigned char*>(ws) + plan_f64_bytes(p) + p->i8_bytes;
    if (p->refine_bytes > 0) VSP_CUDA(cudaMemsetAsync(refine_base, 0, (size_t)p->refine_zero_bytes, st));  // counters, flags
    int* inexact = p->refine_bytes > 0 ? reinterpret_cast<int*>(refine_base + 1024) : nullptr;

    for (int s = 0; s < p->count; ++s) {
        const void* ptr = d_ptrs[p->order[s]];
        if (!ptr) return VSP_E_ARG;
        p->items[s].ptr = ptr;
    }
    VSP_CUDA(cudaMemcpyAsync(p->d_items, p->items.data(), sizeof(ItemDesc) * (size_t)p->count,
                             cudaMemcpyHostToDevice, st));
    t_timer.start(st);

    // per-device attribute; cheap enough to set on every call (one process may drive several GPUs)
    VSP_CUDA(cudaFuncSetAttribute(tridiag_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VSP_CUDA(cudaFuncSetAttribute(bisect_metrics_kernel<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VSP_CUDA(cudaFuncSetAttribute(bisect_metrics_kernel<128, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    VSP_CUDA(cudaFuncSetAttribute(bisect_metrics_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));

    auto mark = [&]() -> int {
        if (!evs) return VSP_OK;
        cudaEvent_t ev;
        if (!cuda_ok(cudaEventCreate(&ev), "cudaEventCreate")) return VSP_E_CUDA;
        evs->push_back(ev);
        return cuda_ok(cudaEventRecord(ev, st), "cudaEventRecord") ? VSP_OK : VSP_E_CUDA;
    };
    for (const ShapeClass& c : p->classes) {
        int rc = mark();
        if (rc != VSP_OK) return rc;
        if (p->dtype == VSP_F32 && p->gram_method == 1)
            rc = launch_gram_i8(p, c, ws, reinterpret_cast<unsigned char*>(ws) + plan_f64_bytes(p), inexact, st);
        else
            rc = (p->dtype == VSP_F32) ? launch_gram<float>(p, c, ws, st) : launch_gram<double>(p, c, ws, st);
        if (rc != VSP_OK) return rc;
        if ((rc = mark()) != VSP_OK) return rc;
        RefineGate gate{nullptr, nullptr, 0, nullptr};
        if (c.refine_slots > 0) {
            gate.counter = reinterpret_cast<int*>(refine_base) + c.refine_counter;
            gate.slot_items = reinterpret_cast<int*>(refine_base + c.refine_items_off);
            gate.slots = c.refine_slots;
            gate.inexact = (p->dtype == VSP_F32 && p->gram_method == 1) ? inexact : nullptr;
        }
        if (c.full == kGramTiled) {
            // two-stage reduction, bandwidth 8, matrix resident in shared memory (sbr8.cuh, chase8.cuh).  One launch
            // per order range: > 136 twelve warps / one CTA per SM, > 72 eight warps / two, else four warps / up to six.
            const int N = sbr8_order(c.n);
            int m_start = 0;
            for (;;) {
                const int order = m_start > 0 ? m_start : N;
                const int stv = sbr8_stride(order);
                int m_stop = order > 136 ? 128 : (order > 72 ? 64 : 0);
                static const bool one_launch8 = std::getenv("VSP_SBR8_ONE_LAUNCH") != nullptr;  // experiments
                if (one_launch8) m_stop = 0;
#define VSP_SBR8_LAUNCH(NW, MINB, NC, TB, TILES)                                                                        \
    {                                                                                                                   \
        const int all_tiles = ((order >> 3) * ((order >> 3) + 1)) >> 1;                                                 \
        const int tiles = std::min(all_tiles, (int)(TILES));                                                            \
        const size_t smem = sbr8_smem_bytes_tiles(tiles, stv, NW);                                                      \
        VSP_CUDA(cudaFuncSetAttribute(sbr8_kernel<NW, MINB, NC, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                      (int)std::max<size_t>(smem, 48 * 1024)));                                         \
        VSP_CUDA(cudaFuncSetAttribute(sbr8_kernel<NW, MINB, NC, TB>, cudaFuncAttributePreferredSharedMemoryCarveout,    \
                                      cudaSharedmemCarveoutMaxShared));                                                 \
        sbr8_kernel<NW, MINB, NC, TB><<<c.count, 32 * NW, smem, st>>>(p->d_items, c.begin, ws, stv, m_start, m_stop,    \
                                                                       tiles);                                          \
    }
                // n > 136: two CTAs per SM (six warps each) with the bottom tile rows in L2 while they last, or
                // (VSP_SBR8_LONE=1) one twelve-warp CTA per SM with everything in shared memory
                static const bool lone = std::getenv("VSP_SBR8_LONE") != nullptr;
                const int two_tiles = (int)((112 * 1024 - sizeof(double) * sbr8_fixed_doubles(stv, 6)) / 512);
                if (order > 136 && !lone) VSP_SBR8_LAUNCH(6, 2, 6, 4, two_tiles)
                else if (order > 136) VSP_SBR8_LAUNCH(12, 1, 6, 2, 1 << 20)
                else if (order > 72) VSP_SBR8_LAUNCH(8, 2, 4, 2, 1 << 20)
                else VSP_SBR8_LAUNCH(4, 4, 2, 2, 1 << 20)
#undef VSP_SBR8_LAUNCH
                g_launches++;
                t_timer.tick("sbr8", st);
                VSP_CUDA(cudaGetLastError());
                if (m_stop == 0) break;
                m_start = m_stop;
            }
            const size_t csm = chase8_smem_bytes(c.n);
            VSP_CUDA(cudaFuncSetAttribute(chase8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)std::max<size_t>(csm, 48 * 1024)));
            VSP_CUDA(cudaFuncSetAttribute(chase8_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            chase8_kernel<<<c.count, kChase8Threads, csm, st>>>(p->d_items, c.begin, c.count, ws, gate);
        } else if (c.full == kGramPacked) {
            // two-stage reduction: blocked Householder to bandwidth 4 (sbr_band.cuh), bulge chasing (band_tridiag.cuh).
            // The blocked stage is launched per order range (n -> 96 -> 48 -> end): a smaller active block means a
            // smaller CTA, so more matrices share an SM while the steps are latency-bound.
            static const bool one_launch = std::getenv("VSP_SBR_ONE_LAUNCH") != nullptr;  // experiments
            int m_start = 0;  // 0: fresh start from n
            static const std::vector<int> stops = [] {  // experiments: VSP_SBR_STOPS="144,96,48"
                std::vector<int> v;
                if (const char* e = std::getenv("VSP_SBR_STOPS")) {
                    for (const char* q = e; *q;) {
                        v.push_back(std::atoi(q));
                        while (*q && *q != ',') ++q;
                        if (*q == ',') ++q;
                    }
                } else {
                    v = {256, 96, 48};
                }
                v.push_back(0);
                return v;
            }();
            for (size_t si = 0; si < stops.size(); ++si) {
                const int order = m_start > 0 ? m_start : c.n;
                int m_stop = one_launch ? 0 : stops[si];
                if (order >= 256 + 32 && m_stop < 256) m_stop = one_launch ? 0 : 256;  // the small-order kernels take over below 256
                if (m_stop > 0 && order < m_stop + 32) continue;  // not worth a launch of its own
                const bool big = order > 256;  // three index blocks per warp, matrix mostly in the global workspace
                const int nw = big ? 8 : sbr_warps(order);
                // stride of the [4][stv] operand arrays: = 4 mod 8 doubles, so that the four k-rows of a DMMA operand
                // fragment (addresses t * stv + g) fall into four different 32-byte bank groups.  With stv a
                // multiple of 32 they all hit the same one (4-way conflict on every operand load): measured
                // 4.97 -> 4.63 ms on the first order range of the Scenario-A sweep.
                static const int stpad = std::getenv("VSP_SBR_STPAD") ? std::atoi(std::getenv("VSP_SBR_STPAD")) : 4;
                const int stv = round_up(order, 32) + stpad;
                switch (big ? 5 : (nw + 1) / 2) {
#define VSP_SBR_CASE(HALF, NQ, MINB, BPW)                                                                              \
    case HALF: {                                                                                                       \
        const size_t budget = std::min<size_t>(MINB == 1 ? 226 * 1024 : kSbrSmemBudget, (227 * 1024) / MINB - 1024);   \
        const int rows_smem = sbr_rows_in_smem(order, stv, nw, budget);                                                \
        const size_t smem = sbr_smem_bytes(std::min(rows_smem, order), stv, nw);                                       \
        VSP_CUDA(cudaFuncSetAttribute(sbr_band_kernel<NQ, MINB, BPW>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                      (int)std::min<size_t>(227 * 1024, std::max<size_t>(budget, 48 * 1024))));        \
        VSP_CUDA(cudaFuncSetAttribute(sbr_band_kernel<NQ, MINB, BPW>, cudaFuncAttributePreferredSharedMemoryCarveout,  \
                                      cudaSharedmemCarveoutMaxShared));                                                \
        sbr_band_kernel<NQ, MINB, BPW><<<c.count, 32 * nw, smem, st>>>(p->d_items, c.begin, ws, stv, rows_smem,        \
                                                                        m_start, m_stop);                              \
    } break;
                    VSP_SBR_CASE(1, 2, 6, 1)
                    VSP_SBR_CASE(2, 4, 3, 1)
                    VSP_SBR_CASE(3, 6, 2, 1)
                    VSP_SBR_CASE(4, 8, 1, 1)
                    VSP_SBR_CASE(5, 8, 1, 3)
#undef VSP_SBR_CASE
                    default:
                        return VSP_E_UNSUPPORTED;
                }
                g_launches++;
                t_timer.tick("sbr_band", st);
                VSP_CUDA(cudaGetLastError());
                if (m_stop == 0) break;
                m_start = order - 4 * ((order - m_stop + 3) / 4);  // the order the launch stopped at
            }
            const size_t csm = band_tridiag_smem_bytes(c.n);
            VSP_CUDA(cudaFuncSetAttribute(band_tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)std::max<size_t>(csm, 48 * 1024)));
            band_tridiag_kernel<<<(c.count + kChaseWarps - 1) / kChaseWarps, 32 * kChaseWarps, csm, st>>>(
                p->d_items, c.begin, c.count, ws, gate);
        } else {
            const int threads = std::min(1024, c.npad);
            tridiag_global_kernel<<<c.count, threads, tridiag_global_smem_bytes(c.npad), st>>>(p->d_items, c.begin,
                                                                                              ws, c.npad, gate);
        }
        g_launches++;
        t_timer.tick("chase/tridiag", st);
        VSP_CUDA(cudaGetLastError());
        if ((rc = mark()) != VSP_OK) return rc;
        // The re-solve of the items the tridiagonalisation flagged runs beside the bisection kernel.
        // Its few fat CTAs (1024 threads, whole SM) must be placed BEFORE the many thin bisection
        // CTAs flood the SMs, so the re-solve stays on the caller's stream (it starts the moment the
        // tridiagonalisation drains) and the bisection is the one that forks to the side stream.
        const bool fork = c.refine_slots > 0;
        cudaStream_t bst = st;
        if (fork) {
            RefinePool pool;
            pool.base = reinterpret_cast<double*>(refine_base + c.refine_off);
            pool.slots = c.refine_slots;
            pool.slot_doubles = c.refine_slot_doubles;
            const size_t fixed = refine_smem_fixed_bytes(c.npad);
            const size_t rsm = 227 * 1024;  // everything beyond the fixed scratch holds the trailing block
            const int xs_doubles = (int)((rsm - fixed) / sizeof(double));
            VSP_CUDA(cudaEventRecord(p->ev_fork, st));
            VSP_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
            if (c.refine_B > 0) {
                const int B = c.refine_B, Kpad = round_up(c.refine_kmax, 4), nloc = (c.n + B - 1) / B;
                const size_t smem_d = std::max(refine_cluster_fixed_doubles(c.npad, Kpad, nloc, c.refine_shared) + (size_t)c.refine_xs_cap,
                                               refine_cluster_tail_doubles(c.npad));
                const size_t csmem = smem_d * sizeof(double);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)(B * c.refine_slots));
                cfg.blockDim = dim3(kRcThreads);
                cfg.dynamicSmemBytes = csmem;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = (unsigned)B;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                const bool shared = c.refine_shared;  // runs beside the bisection kernel: capped registers, maximum carve-out
#define VSP_RC_LAUNCH(KERNEL)                                                                                             \
    {                                                                                                                   \
        VSP_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(csmem, 48 * 1024))); \
        if (shared)                                                                                                     \
            VSP_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        if (B > 8) VSP_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));           \
        if (c.refine_launch_slots == 0) {                                                                               \
            /* no more clusters than are resident at once: a cluster that cannot be placed (a 16-CTA cluster needs a */ \
            /* whole GPC) keeps the bisection kernel behind it out of the SMs until it has been dispatched           */ \
            int active = 0;                                                                                             \
            VSP_CUDA(cudaOccupancyMaxActiveClusters(&active, KERNEL, &cfg));                                            \
            c.refine_launch_slots = std::max(1, std::min(c.refine_slots, active));                                      \
        }                                                                                                               \
        pool.slots = c.refine_launch_slots;                                                                             \
        cfg.gridDim = dim3((unsigned)(B * c.refine_launch_slots));                                                      \
        VSP_CUDA(cudaLaunchKernelEx(&cfg, KERNEL, (const ItemDesc*)p->d_items, gate, pool, c.npad, Kpad, nloc, c.refine_xs_cap, \
                                    p->opts, d_sv, d_records, d_dist));                                                 \
    }
                const bool small = !shared && c.n <= 256;
                if (p->dtype == VSP_F32) {
                    if (shared) VSP_RC_LAUNCH(refine_cluster_shared_kernel<float>)
                    else if (small) VSP_RC_LAUNCH(refine_cluster_small_kernel<float>)
                    else VSP_RC_LAUNCH(refine_cluster_kernel<float>)
                } else {
                    if (shared) VSP_RC_LAUNCH(refine_cluster_shared_kernel<double>)
                    else if (small) VSP_RC_LAUNCH(refine_cluster_small_kernel<double>)
                    else VSP_RC_LAUNCH(refine_cluster_kernel<double>)
                }
#undef VSP_RC_LAUNCH
            } else if (p->dtype == VSP_F32) {
                VSP_CUDA(cudaFuncSetAttribute(refine_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                refine_kernel<float><<<c.refine_slots, 1024, rsm, st>>>(p->d_items, gate, pool, c.npad, xs_doubles, p->opts, d_sv, d_records, d_dist);
            } else {
                VSP_CUDA(cudaFuncSetAttribute(refine_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                refine_kernel<double><<<c.refine_slots, 1024, rsm, st>>>(p->d_items, gate, pool, c.npad, xs_doubles, p->opts, d_sv, d_records, d_dist);
            }
            g_launches++;
            t_timer.tick("refine", st);
            VSP_CUDA(cudaGetLastError());
            static const bool serial_refine = std::getenv("VSP_REFINE_SERIAL") != nullptr;  // experiments: no overlap
            bst = serial_refine ? st : p->side;
        }
        const int bthreads = bisect_threads(c.n);
        if (bthreads <= 128)
            bisect_metrics_kernel<128, 8><<<c.count, bthreads, bisect_smem_bytes(c.npad), bst>>>(p->d_items, c.begin, ws, c.npad,
                                                                                                 p->opts, d_sv, d_records, d_dist);
        else
            bisect_metrics_kernel<1024, 1><<<c.count, bthreads, bisect_smem_bytes(c.npad), bst>>>(p->d_items, c.begin, ws, c.npad,
                                                                                                  p->opts, d_sv, d_records, d_dist);
        g_launches++;
        t_timer.tick("bisect", bst);
        VSP_CUDA(cudaGetLastError());
        if (fork) {  // join
            VSP_CUDA(cudaEventRecord(p->ev_join, p->side));
            VSP_CUDA(cudaStreamWaitEvent(st, p->ev_join, 0));
        }
        if ((rc = mark()) != VSP_OK) return rc;
    }
    t_timer.tick("joined", st);
    t_timer.report(st);
    return VSP_OK;
}

// expand item `blockIdx.x`'s Gram matrix (packed padded rows or full) into a dense n x n block
__global__ void expand_gram_kernel(const ItemDesc* __restrict__ items, const double* __restrict__ ws,
                                   const int64_t* __restrict__ out_off, double* __restrict__ out) {
    const ItemDesc it = items[blockIdx.x];
    const int n = it.n;
    const double* G = ws + it.gram_off;
    double* o = out + out_off[it.item];
    for (int64_t e = threadIdx.x; e < (int64_t)n * n; e += blockDim.x) {
        const int i = (int)(e / n), j = (int)(e % n);
        const int r = i > j ? i : j, c = i > j ? j : i;
        if (it.full == kGramFull) {
            o[e] = G[(int64_t)r * n + c];
        } else if (it.full == kGramTiled) {
            const int off = sbr8_order(n) - n, fr = r + off, fc = c + off;
            o[e] = G[tile_off(fr >> 3, fc >> 3) + ((fr & 7) << 3) + (fc & 7)];
        } else {
            o[e] = G[poff(r) + c];
        }
    }
}

int vsp_plan_debug_gram(vsp_plan* p, const void* const* d_ptrs, double* d_out, void* d_workspace,
                        int64_t workspace_bytes, void* stream) {
    if (!p || !d_ptrs || !d_out || !d_workspace) return VSP_E_ARG;
    if (p->count == 0) return VSP_OK;
    uintptr_t base = reinterpret_cast<uintptr_t>(d_workspace);
    uintptr_t aligned = (base + 255) & ~uintptr_t(255);
    if ((int64_t)(aligned - base) + plan_f64_bytes(p) + p->i8_bytes > workspace_bytes) return VSP_E_WORKSPACE;
    double* ws = reinterpret_cast<double*>(aligned);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    for (int s = 0; s < p->count; ++s) p->items[s].ptr = d_ptrs[p->order[s]];
    VSP_CUDA(cudaMemcpyAsync(p->d_items, p->items.data(), sizeof(ItemDesc) * (size_t)p->count,
                             cudaMemcpyHostToDevice, st));
    for (const ShapeClass& c : p->classes) {
        int rc;
        if (p->dtype == VSP_F32 && p->gram_method == 1)
            rc = launch_gram_i8(p, c, ws, reinterpret_cast<unsigned char*>(ws) + plan_f64_bytes(p), nullptr, st);
        else
            rc = (p->dtype == VSP_F32) ? launch_gram<float>(p, c, ws, st) : launch_gram<double>(p, c, ws, st);
        if (rc != VSP_OK) return rc;
    }
    std::vector<int64_t> off(p->count + 1, 0);  // caller order
    {
        std::vector<int> nn(p->count);
        for (int s = 0; s < p->count; ++s) nn[p->items[s].item] = p->items[s].n;
        for (int i = 0; i < p->count; ++i) off[i + 1] = off[i] + (int64_t)nn[i] * nn[i];
    }
    int64_t* d_off = nullptr;
    VSP_CUDA(cudaMalloc(&d_off, sizeof(int64_t) * (size_t)p->count));
    VSP_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(int64_t) * (size_t)p->count, cudaMemcpyHostToDevice, st));
    expand_gram_kernel<<<p->count, 256, 0, st>>>(p->d_items, ws, d_off, d_out);
    VSP_CUDA(cudaGetLastError());
    VSP_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_off);
    return VSP_OK;
}

int vsp_plan_execute(vsp_plan* p, const void* const* d_ptrs, double* d_sv, vsp_record* d_records,
                     void* d_workspace, int64_t workspace_bytes, void* stream) {
    return execute_impl(p, d_ptrs, d_sv, d_records, d_workspace, workspace_bytes, stream, nullptr);
}

int vsp_plan_execute_dist(vsp_plan* p, const void* const* d_ptrs, double* d_sv, vsp_record* d_records, double* d_dist,
                          void* d_workspace, int64_t workspace_bytes, void* stream) {
    if (!p || (p->opts.dist_k <= 0 && p->opts.clauset <= 0) || !d_dist) return VSP_E_ARG;
    return execute_impl(p, d_ptrs, d_sv, d_records, d_workspace, workspace_bytes, stream, nullptr, d_dist);
}

int vsp_plan_execute_profiled(vsp_plan* p, const void* const* d_ptrs, double* d_sv, vsp_record* d_records,
                              void* d_workspace, int64_t workspace_bytes, void* stream, float* stage_ms) {
    if (!stage_ms) return VSP_E_ARG;
    std::vector<cudaEvent_t> evs;
    int rc = execute_impl(p, d_ptrs, d_sv, d_records, d_workspace, workspace_bytes, stream, &evs);
    stage_ms[0] = stage_ms[1] = stage_ms[2] = 0.f;
    if (rc == VSP_OK && !cuda_ok(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)), "cudaStreamSynchronize"))
        rc = VSP_E_CUDA;
    if (rc == VSP_OK) {
        for (size_t i = 0; i + 3 < evs.size(); i += 4)  // 4 marks per shape class
            for (int sidx = 0; sidx < 3; ++sidx) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, evs[i + sidx], evs[i + sidx + 1]) == cudaSuccess) stage_ms[sidx] += ms;
            }
    }
    for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
    return rc;
}

int vsp_dgemm_batched(int32_t batch, int32_t M, int32_t N, int32_t K, double alpha, const double* const* d_A, int64_t lda,
                      int32_t transA, const double* const* d_B, int64_t ldb, int32_t transB, double beta, double gamma,
                      double* const* d_C, int64_t ldc, void* stream) {
    if (batch < 0 || M < 1 || N < 1 || K < 1 || !d_A || !d_B || !d_C) return VSP_E_ARG;
    if (batch == 0) return VSP_OK;
    if (batch > 65535) return VSP_E_UNSUPPORTED;
    const dim3 grid((N + kGemmTile - 1) / kGemmTile, (M + kGemmTile - 1) / kGemmTile, batch);
    dgemm_dmma_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(M, N, K, alpha, d_A, lda, transA, d_B, ldb, transB,
                                                                                beta, gamma, d_C, ldc);
    g_launches++;
    return cuda_ok(cudaGetLastError(), "dgemm_dmma_kernel") ? VSP_OK : VSP_E_CUDA;
}

int vsp_analyze_batch(const void* const* d_ptrs, const int32_t* rows, const int32_t* cols, const int64_t* ld,
                      int32_t dtype, int32_t count, const vsp_opts* opts, double* d_sv, vsp_record* d_records,
                      void* d_workspace, int64_t workspace_bytes, void* stream) {
    vsp_plan* plan = nullptr;
    int rc = vsp_plan_create(count, rows, cols, ld, dtype, opts, &plan);
    if (rc != VSP_OK) return rc;
    rc = vsp_plan_execute(plan, d_ptrs, d_sv, d_records, d_workspace, workspace_bytes, stream);
    // the item table must outlive the kernels that read it
    if (plan->d_items) cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
    vsp_plan_destroy(plan);
    return rc;
}

int vsp_analyze_batch_host(const void* const* h_ptrs, const int32_t* rows, const int32_t* cols, const int64_t* ld,
                           int32_t dtype, int32_t count, const vsp_opts* opts, double* h_sv, vsp_record* h_records,
                           int32_t device) {
    if (dtype != VSP_F32 && dtype != VSP_F64) return VSP_E_UNSUPPORTED;
    int rc = validate(count, rows, cols, ld);
    if (rc != VSP_OK) return rc;
    if (count == 0) return VSP_OK;
    if (!h_ptrs || !h_records) return VSP_E_ARG;
    const bool want_sv = !(opts && opts->want_sv == 0);
    if (want_sv && !h_sv) return VSP_E_ARG;
    VSP_CUDA(cudaSetDevice(device));
    const size_t esz = dtype == VSP_F32 ? 4 : 8;

    // dense device layout, 16-byte aligned; contiguous host ranges become one copy
    std::vector<int64_t> doff(count + 1, 0);
    for (int i = 0; i < count; ++i)
        doff[i + 1] = round_up64(doff[i] + (int64_t)rows[i] * cols[i] * (int64_t)esz, 16);
    char* d_in = nullptr;
    void* d_ws = nullptr;
    double* d_sv = nullptr;
    vsp_record* d_rec = nullptr;
    vsp_plan* plan = nullptr;
    cudaStream_t st = nullptr;
    auto cleanup = [&]() {
        if (plan) vsp_plan_destroy(plan);
        if (d_in) cudaFree(d_in);
        if (d_ws) cudaFree(d_ws);
        if (d_sv) cudaFree(d_sv);
        if (d_rec) cudaFree(d_rec);
        if (st) cudaStreamDestroy(st);
    };
#define VSP_CUDA_C(call)                       \
    do {                                       \
        if (!cuda_ok((call), #call)) {         \
            cleanup();                         \
            return VSP_E_CUDA;                 \
        }                                      \
    } while (0)
    VSP_CUDA_C(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    rc = vsp_plan_create(count, rows, cols, nullptr, dtype, opts, &plan);
    if (rc != VSP_OK) {
        cleanup();
        return rc;
    }
    const int64_t ws_bytes = vsp_plan_workspace_bytes(plan);
    VSP_CUDA_C(cudaMalloc(&d_in, (size_t)doff[count] + 16));
    VSP_CUDA_C(cudaMalloc(&d_ws, (size_t)ws_bytes));
    VSP_CUDA_C(cudaMalloc(&d_rec, sizeof(vsp_record) * (size_t)count));
    if (want_sv) VSP_CUDA_C(cudaMalloc(&d_sv, sizeof(double) * (size_t)std::max<int64_t>(1, plan->sv_total)));

    std::vector<const void*> dptrs(count);
    int i = 0;
    while (i < count) {
        if (!h_ptrs[i]) {
            cleanup();
            return VSP_E_ARG;
        }
        dptrs[i] = d_in + doff[i];
        const int64_t hl = ld ? ld[i] : cols[i];
        if (hl != cols[i]) {  // strided host view
            VSP_CUDA_C(cudaMemcpy2DAsync(d_in + doff[i], (size_t)cols[i] * esz, h_ptrs[i], (size_t)hl * esz,
                                         (size_t)cols[i] * esz, (size_t)rows[i], cudaMemcpyHostToDevice, st));
            ++i;
            continue;
        }
        // merge a run of host-contiguous, device-contiguous matrices
        int j = i;
        int64_t bytes = (int64_t)rows[i] * cols[i] * (int64_t)esz;
        while (j + 1 < count && h_ptrs[j + 1] &&
               (!ld || ld[j + 1] == cols[j + 1]) &&
               static_cast<const char*>(h_ptrs[j + 1]) == static_cast<const char*>(h_ptrs[i]) + bytes &&
               doff[j + 1] == doff[i] + bytes) {
            ++j;
            dptrs[j] = d_in + doff[j];
            bytes += (int64_t)rows[j] * cols[j] * (int64_t)esz;
        }
        VSP_CUDA_C(cudaMemcpyAsync(d_in + doff[i], h_ptrs[i], (size_t)bytes, cudaMemcpyHostToDevice, st));
        i = j + 1;
    }
    rc = vsp_plan_execute(plan, dptrs.data(), d_sv, d_rec, d_ws, ws_bytes, st);
    if (rc != VSP_OK) {
        cudaStreamSynchronize(st);
        cleanup();
        return rc;
    }
    VSP_CUDA_C(cudaMemcpyAsync(h_records, d_rec, sizeof(vsp_record) * (size_t)count, cudaMemcpyDeviceToHost, st));
    if (want_sv && plan->sv_total > 0)
        VSP_CUDA_C(cudaMemcpyAsync(h_sv, d_sv, sizeof(double) * (size_t)plan->sv_total, cudaMemcpyDeviceToHost, st));
    VSP_CUDA_C(cudaStreamSynchronize(st));
    cleanup();
    return VSP_OK;
#undef VSP_CUDA_C
}

}  // extern "C"
