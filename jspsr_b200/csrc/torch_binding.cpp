// Thin torch extension over the C ABI (include/jspsr_spn.h, include/jspsr_tiles.h): the training-step chain of the
// hot path - PostProcessor.forward / backward (models/components/spn.py:99-118 + autograd) and the YAML configs'
// MultiLoss (losses/loss_schemes.py:55-72) - as C++ autograd functions.
//
// Why: at the reference's batch sizes (70 / 50 tiles, configs/*.yml:86) the three kernels of a step take ~90 us on a
// B200 while the Python + ctypes wrappers in functional.py / epilogue.py spend ~250 us of host time enqueueing them
// (tools/eager_overhead.py).  This file does the same bookkeeping (output allocation from torch's caching allocator,
// current stream, per-(device, stream) reduction workspace, dtype code, error -> exception) in C++.  No arithmetic
// lives here: every tensor operation below is an allocation or a view; the kernels are libjspsr_spn.so's.
// jspsr_b200/functional.py validates the arguments (same messages as before) and falls back to the ctypes binding of
// the SAME library when this extension has not been built; there is no CPU path in either.
#include <torch/extension.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/jspsr_peer.h"
#include "../../include/jspsr_tiles.h"

namespace {

std::atomic<int64_t> g_launches{0};
std::mutex g_ws_mutex;
// heap-allocated and never destroyed: the tensors must not be released by a static destructor after torch's CUDA
// allocator has shut down at interpreter exit
auto& g_ws = *new std::map<std::pair<int, void*>, at::Tensor>();

// zero-initialised reduction scratch, one per (device, stream); the kernels leave it zeroed
at::Tensor workspace_for(const at::Tensor& like, void* stream) {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    const auto key = std::make_pair((int)like.get_device(), stream);
    auto it = g_ws.find(key);
    if (it != g_ws.end()) return it->second;
    at::Tensor ws = at::zeros({(int64_t)jspsr_spn_workspace_bytes()}, like.options().dtype(at::kByte));
    g_ws.emplace(key, ws);
    return ws;
}

void check(int rc, const char* what) {
    TORCH_CHECK(rc == 0, what, " failed (", rc, "): ", jspsr_last_error());
}

int io_code(const at::Tensor& init, const at::Tensor& weight) {
    if (init.scalar_type() == at::kFloat && weight.scalar_type() == at::kBFloat16) return JSPSR_MIXED;
    if (init.scalar_type() == at::kFloat) return JSPSR_F32;
    TORCH_CHECK(init.scalar_type() == at::kBFloat16, "jspsr_b200: unsupported dtype (float32 and bfloat16 only)");
    return JSPSR_BF16;
}

// fp32, contiguous, on `like`'s device (parameters normally are: no copy then)
at::Tensor as_f32(const at::Tensor& t, const at::Tensor& like) {
    at::Tensor d = t.detach();
    if (d.scalar_type() != at::kFloat || d.device() != like.device()) d = d.to(like.device(), at::kFloat);
    return d.contiguous();
}

at::Tensor spn_forward_raw(const at::Tensor& init, const at::Tensor& weight, const at::Tensor& offset, const at::Tensor& w,
                           const at::Tensor& b, int64_t norm_mode, double scale) {
    const c10::cuda::CUDAGuard guard(init.device());
    const at::Tensor i = init.contiguous(), wt = weight.contiguous(), of = offset.contiguous();
    const at::Tensor w9 = as_f32(w, i), b1 = as_f32(b, i);
    at::Tensor out = at::empty_like(i);
    void* stream = at::cuda::getCurrentCUDAStream(i.get_device()).stream();
    check(jspsr_spn_forward(i.data_ptr(), wt.data_ptr(), of.data_ptr(), w9.data_ptr<float>(), b1.data_ptr<float>(),
                            out.data_ptr(), (int)i.size(0), (int)i.size(2), (int)i.size(3), (int)norm_mode, (float)scale,
                            io_code(i, wt), stream),
          "jspsr_spn_forward");
    ++g_launches;
    return out;
}

class PropagateFn : public torch::autograd::Function<PropagateFn> {
   public:
    static at::Tensor forward(torch::autograd::AutogradContext* ctx, const at::Tensor& init, const at::Tensor& weight,
                              const at::Tensor& offset, const at::Tensor& w, const at::Tensor& b, int64_t norm_mode,
                              double scale, int64_t reduce_ptr) {
        ctx->save_for_backward({init, weight, offset, w});
        ctx->saved_data["norm_mode"] = norm_mode;
        ctx->saved_data["scale"] = scale;
        ctx->saved_data["reduce_ptr"] = reduce_ptr;  // host address of a jspsr_peer_reduce kept alive by the caller, or 0
        return spn_forward_raw(init, weight, offset, w, b, norm_mode, scale);
    }

    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_outputs) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor init = saved[0].contiguous(), weight = saved[1].contiguous(), offset = saved[2].contiguous();
        const at::Tensor& w = saved[3];
        const int64_t norm_mode = ctx->saved_data["norm_mode"].toInt();
        const double scale = ctx->saved_data["scale"].toDouble();
        const auto* reduce = reinterpret_cast<const jspsr_peer_reduce*>(ctx->saved_data["reduce_ptr"].toInt());
        const bool need_init = ctx->needs_input_grad(0);
        const bool need_w = ctx->needs_input_grad(3) || ctx->needs_input_grad(4);

        const c10::cuda::CUDAGuard guard(init.device());
        at::Tensor gout = grad_outputs[0];
        if (gout.scalar_type() != init.scalar_type()) gout = gout.to(init.scalar_type());
        gout = gout.contiguous();
        const at::Tensor w9 = as_f32(w, init);
        const auto f32 = init.options().dtype(at::kFloat);
        at::Tensor grad_init, grad_wb, ws;
        if (need_init) grad_init = at::empty(init.sizes(), f32);
        at::Tensor grad_weight = at::empty_like(weight), grad_offset = at::empty_like(offset);
        if (need_w) grad_wb = at::empty({10}, f32);   // grad_w[9] then grad_b[1]
        void* stream = at::cuda::getCurrentCUDAStream(init.get_device()).stream();
        if (need_w) ws = workspace_for(init, stream);
        check(jspsr_spn_backward_reduce(gout.data_ptr(), init.data_ptr(), weight.data_ptr(), offset.data_ptr(), w9.data_ptr<float>(),
                                 need_init ? grad_init.data_ptr<float>() : nullptr, grad_weight.data_ptr(),
                                 grad_offset.data_ptr(), need_w ? grad_wb.data_ptr<float>() : nullptr,
                                 need_w ? grad_wb.data_ptr<float>() + 9 : nullptr, need_w ? ws.data_ptr() : nullptr,
                                 (int)init.size(0), (int)init.size(2), (int)init.size(3), (int)norm_mode, (float)scale,
                                 io_code(init, weight), 0u, need_w ? reduce : nullptr, stream),
              "jspsr_spn_backward");
        ++g_launches;

        at::Tensor gi, gw, gb;
        if (need_init) gi = grad_init.scalar_type() == init.scalar_type() ? grad_init : grad_init.to(init.scalar_type());
        if (need_w) {
            gw = grad_wb.narrow(0, 0, 9).view(w.sizes());
            gb = grad_wb.narrow(0, 9, 1);
            if (w.scalar_type() != at::kFloat) {
                gw = gw.to(w.scalar_type());
                gb = gb.to(w.scalar_type());
            }
        }
        return {gi,
                ctx->needs_input_grad(1) ? grad_weight : at::Tensor(),
                ctx->needs_input_grad(2) ? grad_offset : at::Tensor(),
                ctx->needs_input_grad(3) ? gw : at::Tensor(),
                ctx->needs_input_grad(4) ? gb : at::Tensor(),
                at::Tensor(),
                at::Tensor(),
                at::Tensor()};
    }
};

at::Tensor propagate(const at::Tensor& init, const at::Tensor& weight, const at::Tensor& offset, const at::Tensor& w,
                     const at::Tensor& b, int64_t norm_mode, double scale, int64_t reduce_ptr) {
    return PropagateFn::apply(init, weight, offset, w, b, norm_mode, scale, reduce_ptr);
}

// -> (losses [4] = L1, L2, Grad, Total; dTotal/dpred or an undefined tensor)
std::pair<at::Tensor, at::Tensor> loss_raw(const at::Tensor& pred, const at::Tensor& gt, double w_l1, double w_l2, double w_grad,
                                           bool want_grad) {
    const c10::cuda::CUDAGuard guard(pred.device());
    const at::Tensor p = pred.detach().contiguous(), g = gt.detach().contiguous();
    at::Tensor losses = at::empty({4}, p.options());
    at::Tensor grad;
    if (want_grad) grad = at::empty_like(p);
    void* stream = at::cuda::getCurrentCUDAStream(p.get_device()).stream();
    const at::Tensor ws = workspace_for(p, stream);
    check(jspsr_loss_l1_l2_grad(p.data_ptr<float>(), g.data_ptr<float>(), (float)w_l1, (float)w_l2, (float)w_grad,
                                losses.data_ptr<float>(), want_grad ? grad.data_ptr<float>() : nullptr, ws.data_ptr(),
                                (int)(p.size(0) * p.size(1)), (int)p.size(2), (int)p.size(3), stream),
          "jspsr_loss_l1_l2_grad");
    ++g_launches;
    return {losses, grad};
}

class LossFn : public torch::autograd::Function<LossFn> {
   public:
    // returns (Total [scalar, differentiable], losses [4], non-differentiable)
    static torch::autograd::variable_list forward(torch::autograd::AutogradContext* ctx, const at::Tensor& pred,
                                                  const at::Tensor& gt, double w_l1, double w_l2, double w_grad) {
        const bool need = pred.requires_grad();
        auto r = loss_raw(pred, gt, w_l1, w_l2, w_grad, need);
        if (need) ctx->save_for_backward({r.second});
        ctx->mark_non_differentiable({r.first});
        return {r.first.select(0, 3).clone(), r.first};
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_outputs) {
        // the kernel wrote dTotal/dpred for an upstream gradient of 1; scale by the actual one
        const auto saved = ctx->get_saved_variables();
        return {saved[0] * grad_outputs[0], at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor()};
    }
};

std::vector<at::Tensor> multi_loss(const at::Tensor& pred, const at::Tensor& gt, double w_l1, double w_l2, double w_grad) {
    return LossFn::apply(pred, gt, w_l1, w_l2, w_grad);
}

int64_t launch_count() { return g_launches.load(); }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "C++ autograd wrappers over libjspsr_spn.so's C ABI (propagation forward/backward, fused loss)";
    m.def("propagate", &propagate, "normalise -> deformable 3x3 gather -> (+ scale*init), differentiable; reduce_ptr: host "
          "address of a jspsr_peer_reduce (gradients of w / b all-reduced inside the backward kernel) or 0",
          py::arg("init"), py::arg("weight"), py::arg("offset"), py::arg("w"), py::arg("b"), py::arg("norm_mode"),
          py::arg("scale"), py::arg("reduce_ptr") = 0);
    m.def("spn_forward", &spn_forward_raw, "forward only, no autograd");
    m.def("multi_loss", &multi_loss, "(Total, losses[4]) of L1 + L2 + Sobel-L1; Total is differentiable w.r.t. pred");
    m.def("launch_count", &launch_count, "kernels enqueued through this extension");
    m.def("abi_version", []() { return jspsr_version(); });
}
