// torch_ext.cpp -- PyTorch extension layer of the drop-in boundary.
//
//   torch.ops.vision_spectra_b200.analyze_batch(Tensor[] matrices, int fit_start, int fit_end, int hill_k, bool want_sv,
//                                               int dist_k, bool clauset) -> (Tensor records, Tensor singular_values, Tensor aux)
//
// One ragged batch of 2-D CUDA tensors (one dtype, float32 or float64, unit column stride; q/k/v may be row-block
// views of a fused qkv buffer) through the C-ABI of include/vspectra.h: replaces the reference's per-matrix loop
// experiments/run_spectral_analysis.py:323-336 / training/base.py:399-405.  Asynchronous on the CURRENT stream of the
// tensors' device, under a device guard; inputs are borrowed; the outputs -- records uint8 [count, 64] (vsp_record) and
// singular values float64 [sum min(rows, cols)], descending per matrix, and with dist_k > 0 the truncated distribution
// arrays / Clauset block float64 [count, VSP_AUX_STRIDE] (vspectra.h: vsp_plan_execute_dist) -- and the workspace are ATen allocations, so
// the caching allocator's stream ordering keeps them alive exactly as long as the kernels need them.  No host
// synchronisation, no CPU fallback: a non-CUDA tensor is an error.
//
// Plans (validated shape tables resident on the device) are cached per (device, stream, shapes, options); a plan
// is executed by one stream at a time (vspectra.h), which the stream in the key guarantees.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <list>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/vspectra.h"

namespace {

struct PlanCache {
    struct Entry {
        std::string key;
        vsp_plan* plan;
        cudaStream_t stream;
    };
    std::list<Entry> lru;  // front = most recent
    std::unordered_map<std::string, std::list<Entry>::iterator> index;
    std::mutex mu;
    static constexpr size_t kMax = 32;

    vsp_plan* get(const std::string& key, cudaStream_t stream, const std::vector<int32_t>& rows, const std::vector<int32_t>& cols,
                  const std::vector<int64_t>& ld, int32_t dtype, const vsp_opts& opts) {
        std::lock_guard<std::mutex> lock(mu);
        auto it = index.find(key);
        if (it != index.end()) {
            lru.splice(lru.begin(), lru, it->second);
            return it->second->plan;
        }
        vsp_plan* plan = nullptr;
        const int rc = vsp_plan_create((int32_t)rows.size(), rows.data(), cols.data(), ld.data(), dtype, &opts, &plan);
        TORCH_CHECK(rc == VSP_OK, "vsp_plan_create failed: ", vsp_error_string(rc), " (", vsp_last_cuda_error(), ")");
        lru.push_front(Entry{key, plan, stream});
        index[key] = lru.begin();
        while (lru.size() > kMax) {
            Entry& old = lru.back();
            cudaStreamSynchronize(old.stream);  // its kernels may still read the item table
            vsp_plan_destroy(old.plan);
            index.erase(old.key);
            lru.pop_back();
        }
        return plan;
    }
};

PlanCache& cache() {
    static PlanCache* c = new PlanCache();  // leaked on purpose: plans die with the CUDA context
    return *c;
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> analyze_batch_cuda(at::TensorList matrices, int64_t fit_start, int64_t fit_end,
                                                                  int64_t hill_k, bool want_sv, int64_t dist_k, bool clauset) {
    const int64_t count = (int64_t)matrices.size();
    TORCH_CHECK(count > 0, "analyze_batch: empty batch");
    const at::Tensor& first = matrices[0];
    TORCH_CHECK(first.is_cuda(), "analyze_batch: CUDA tensors only (the spectral path has no CPU fallback)");
    const auto st = first.scalar_type();
    TORCH_CHECK(st == at::kFloat || st == at::kDouble, "analyze_batch: float32 or float64 tensors");
    const c10::cuda::CUDAGuard guard(first.device());
    const cudaStream_t stream = at::cuda::getCurrentCUDAStream(first.device().index()).stream();

    std::vector<int32_t> rows(count), cols(count);
    std::vector<int64_t> ld(count);
    std::vector<const void*> ptrs(count);
    int64_t sv_total = 0;
    for (int64_t i = 0; i < count; ++i) {
        const at::Tensor& t = matrices[i];
        TORCH_CHECK(t.device() == first.device() && t.scalar_type() == st && t.dim() == 2 && t.numel() > 0,
                    "analyze_batch: tensors must be non-empty, 2-D, of one dtype, on one device");
        TORCH_CHECK(t.size(1) == 1 || t.stride(1) == 1, "analyze_batch: unit column stride required");
        rows[i] = (int32_t)t.size(0);
        cols[i] = (int32_t)t.size(1);
        ld[i] = t.size(0) > 1 ? t.stride(0) : std::max<int64_t>(t.size(1), t.stride(0));
        TORCH_CHECK(ld[i] >= cols[i], "analyze_batch: overlapping rows (stride(0) < cols)");
        ptrs[i] = t.data_ptr();
        sv_total += std::min(rows[i], cols[i]);
    }
    vsp_opts opts{};
    opts.fit_start = (int32_t)fit_start;
    opts.fit_end = (int32_t)fit_end;
    opts.hill_k = (int32_t)hill_k;
    opts.want_sv = want_sv ? 1 : 0;
    opts.refine = -1;
    opts.dist_k = dist_k > 0 ? (int32_t)dist_k : 0;
    opts.clauset = clauset ? 1 : 0;
    const int32_t dtype = st == at::kFloat ? VSP_F32 : VSP_F64;

    std::string key;
    key.reserve(64 + 16 * (size_t)count);
    auto put = [&key](const void* p, size_t n) { key.append(reinterpret_cast<const char*>(p), n); };
    const int64_t head[9] = {(int64_t)first.device().index(), (int64_t)(uintptr_t)stream, dtype, fit_start, fit_end, hill_k, want_sv ? 1 : 0,
                             opts.dist_k, opts.clauset};
    put(head, sizeof head);
    put(rows.data(), sizeof(int32_t) * rows.size());
    put(cols.data(), sizeof(int32_t) * cols.size());
    put(ld.data(), sizeof(int64_t) * ld.size());
    vsp_plan* plan = cache().get(key, stream, rows, cols, ld, dtype, opts);

    const auto bytes = first.options().dtype(at::kByte);
    at::Tensor records = at::empty({count, (int64_t)sizeof(vsp_record)}, bytes);
    at::Tensor sv = at::empty({want_sv ? sv_total : 0}, first.options().dtype(at::kDouble));
    const int64_t ws_bytes = vsp_plan_workspace_bytes(plan);
    at::Tensor ws = at::empty({ws_bytes}, bytes);
    const bool aux = opts.dist_k > 0 || opts.clauset > 0;
    at::Tensor dist = at::empty({aux ? count : 0, (int64_t)VSP_AUX_STRIDE(opts.dist_k, opts.clauset)}, first.options().dtype(at::kDouble));
    double* svp = want_sv ? sv.data_ptr<double>() : nullptr;
    vsp_record* recp = reinterpret_cast<vsp_record*>(records.data_ptr());
    const int rc = aux
                       ? vsp_plan_execute_dist(plan, ptrs.data(), svp, recp, dist.data_ptr<double>(), ws.data_ptr(), ws_bytes, stream)
                       : vsp_plan_execute(plan, ptrs.data(), svp, recp, ws.data_ptr(), ws_bytes, stream);
    TORCH_CHECK(rc == VSP_OK, "vsp_plan_execute failed: ", vsp_error_string(rc), " (", vsp_last_cuda_error(), ")");
    return std::make_tuple(records, sv, dist);
}

}  // namespace

TORCH_LIBRARY(vision_spectra_b200, m) {
    m.def("analyze_batch(Tensor[] matrices, int fit_start=-1, int fit_end=-1, int hill_k=-1, bool want_sv=True, int dist_k=0, bool clauset=False) -> (Tensor, Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(vision_spectra_b200, CUDA, m) { m.impl("analyze_batch", &analyze_batch_cuda); }
