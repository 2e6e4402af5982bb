// ts2d_pybind.cpp -- PYBIND11_MODULE(_C_native): the thin C++ / pybind11 layer between torch and the C ABI of libts2d.so.
//
// Stands in for the reference's pybind module (R2D/ext.cpp:4-9) and its tensor front-end (R2D/src/extension_interface.cu:19-260):
// the same two entry points with the same positional arguments, tensor checks and error texts, output / state allocation through
// torch, the current CUDA stream -- and nothing else: all device work happens behind include/ts2d.h.  triangle_splatting_b200/_C.py
// routes the reference-shaped single-GPU call here (the tile-sharded and parameter-space variants stay in Python on the same ABI).
//
// What it adds over the reference's layer, and why it is not a Python concern:
//   * one-enqueue forward (ts2d_forward): the binning state is sized from the largest instance count seen for the problem shape;
//     the frame's true R is read behind an event while the rest of the frame is queued (the reference blocks, rasterizer.cu:190-193);
//   * one-enqueue backward: the row array of the atomics-free gradient write-back is sized the same way, the row count of the frame
//     is checked after the pass has been queued; a frame that needed more is composited again (idempotent) with the exact size;
//   * a pool of ts2d_counters handles (pinned host memory + events are expensive to create).
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <vector>

#include "ts2d.h"

namespace {

constexpr int kMaxChannels = TS2D_MAX_CHANNELS;

std::mutex g_mu;
std::vector<ts2d_counters *> g_pool;                                          // idle counters handles
using ShapeKey = std::tuple<int, int64_t, int, int, int>;                      // device, P, W, H, primitive
std::map<ShapeKey, int64_t> g_r_seen, g_rows_seen;                            // largest R / backward row count seen per shape
int64_t g_capacity_margin = int64_t(1) << 20, g_rows_margin = int64_t(1) << 16;
bool g_sync_forward = false, g_sync_backward = false;
int64_t g_counters_created = 0;

void check(int rc, const char *what)
{
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + ts2d_error_string(rc) + " (ts2d code " + std::to_string(rc) + ")");
}

// ts2d_counters handle of one forward pass, handed from forward to backward through Python (the NumRendered int carries it)
struct FrameCounters {
    ts2d_counters *h = nullptr;
    FrameCounters()
    {
        {
            std::lock_guard<std::mutex> lk(g_mu);
            if (!g_pool.empty()) {
                h = g_pool.back();
                g_pool.pop_back();
            }
        }
        if (!h) {
            check(ts2d_counters_create(&h), "ts2d_counters_create");
            std::lock_guard<std::mutex> lk(g_mu);
            g_counters_created++;
        }
    }
    ~FrameCounters()
    {
        if (!h) return;
        std::lock_guard<std::mutex> lk(g_mu);
        g_pool.push_back(h);
    }
    FrameCounters(const FrameCounters &) = delete;
    FrameCounters &operator=(const FrameCounters &) = delete;
    int64_t num_rendered()
    {
        int64_t v = 0;
        py::gil_scoped_release nogil;
        check(ts2d_counters_num_rendered(h, &v), "ts2d_counters_num_rendered");
        return v;
    }
    int64_t backward_rows()
    {
        int64_t v = 0;
        py::gil_scoped_release nogil;
        check(ts2d_counters_backward_rows(h, &v), "ts2d_counters_backward_rows");
        return v;
    }
};

const float *fptr(const torch::Tensor &t) { return t.defined() && t.numel() > 0 ? t.data_ptr<float>() : nullptr; }

void require_cuda_f32(const char *name, const torch::Tensor &t)
{
    if (!t.defined() || t.numel() == 0) return;
    if (!t.is_cuda()) throw std::runtime_error(std::string(name) + " must be a CUDA tensor: this rasterizer has no CPU path");
    if (t.scalar_type() != torch::kFloat32) throw std::runtime_error(std::string(name) + " must be float32");
}

struct Derived {
    int64_t P;
    bool use_shs;
    int C, M;
};

// P, use_shs, C, M exactly as extension_interface.cu:41-50
Derived derive(const torch::Tensor &vertex, const torch::Tensor &shs, const torch::Tensor &feature)
{
    Derived d;
    d.P = vertex.dim() > 0 ? vertex.size(0) : 0;
    d.use_shs = feature.dim() <= 1 || (feature.size(0) == 0 && shs.size(0) > 0);
    d.C = d.use_shs ? 3 : (int)feature.size(1);
    d.M = (shs.dim() > 0 && shs.size(0) != 0) ? (int)shs.size(1) : 0;
    return d;
}

struct Packed {
    ts2d_camera cam;
    ts2d_geometry geom;
    ts2d_flags flags;
};

Packed pack(int W, int H, float tan_fovx, float tan_fovy, const torch::Tensor &viewmatrix, const torch::Tensor &projmatrix, const torch::Tensor &campos,
            int sh_degree, float gamma, float scale_modifier, float background_depth, const torch::Tensor &background, const torch::Tensor &vertex,
            const torch::Tensor &shs, const torch::Tensor &feature, const torch::Tensor &opacity, const Derived &d, bool back_culling, bool rich_info,
            bool debug, int primitive, bool exact)
{
    Packed p;
    p.cam = ts2d_camera{W, H, tan_fovx, tan_fovy, fptr(viewmatrix), fptr(projmatrix), fptr(campos)};
    p.geom = ts2d_geometry{(int32_t)d.P, sh_degree, d.M, d.C, d.use_shs ? 1 : 0, gamma, scale_modifier, background_depth, fptr(background), fptr(vertex),
                           d.use_shs ? fptr(shs) : nullptr, d.use_shs ? nullptr : fptr(feature), fptr(opacity), nullptr};
    p.flags = ts2d_flags{back_culling ? 1 : 0, rich_info ? 1 : 0, debug ? 1 : 0, 0, 1, exact ? 1 : 0, primitive};
    return p;
}

// rasterizeTrianglesForward (extension_interface.cu:19-152)
// -> (num_rendered, out_feature, radii, depth, normal, contrib_sum, contrib_max, geometryBuffer, binningBuffer, imageBuffer, counters)
py::tuple rasterize_triangles(int image_width, int image_height, float tan_fovx, float tan_fovy, const torch::Tensor &viewmatrix,
                              const torch::Tensor &projmatrix, const torch::Tensor &campos, int sh_degree, float gamma, float scale_modifier,
                              float background_depth, const torch::Tensor &background, const torch::Tensor &vertex, const torch::Tensor &shs,
                              const torch::Tensor &feature, const torch::Tensor &opacity, bool back_culling, bool rich_info, bool debug, int primitive,
                              bool exact)
{
    const Derived d = derive(vertex, shs, feature);
    // extension_interface.cu:53-76 (AT_ERROR -> RuntimeError)
    if (vertex.dim() != 3 || vertex.size(1) != 3 || vertex.size(2) != 3) throw std::runtime_error("vertex must have dimensions (num_points, 3, 3)");
    if (!d.use_shs && feature.dim() != 2) throw std::runtime_error("feature must have dimensions (num_points, num_channels)");
    if (d.use_shs && shs.dim() != 3) throw std::runtime_error("shs must have dimensions (num_points, (1 + sh_degree) ** 2, 3)");
    if (d.C > kMaxChannels) throw std::runtime_error("feature's num_channels can't be larger than MAX_CHANNELS");
    if (d.C != background.size(0)) throw std::runtime_error("background must have the same number of channels as feature");
    if (gamma < 0.0f) throw std::runtime_error("gamma must be larger than 0");
    const int H = image_height, W = image_width;
    const torch::Tensor *all[] = {&viewmatrix, &projmatrix, &campos, &background, &vertex, &shs, &feature, &opacity};
    const char *names[] = {"viewmatrix", "projmatrix", "campos", "background", "vertex", "shs", "feature", "opacity"};
    for (const torch::Tensor *t : all)
        if (!t->is_contiguous()) throw std::runtime_error("input tensors must be contiguous");
    for (int i = 0; i < 8; i++) require_cuda_f32(names[i], *all[i]);
    if (d.use_shs && d.P > 0 && (sh_degree < 0 || sh_degree > 3 || (sh_degree + 1) * (sh_degree + 1) > d.M))
        throw std::runtime_error(ts2d_error_string(TS2D_E_SH_DEGREE));
    if (d.P > 0 && opacity.numel() != d.P) throw std::runtime_error("opacity must have dimensions (num_points, 1)");

    const torch::Device dev = vertex.is_cuda() ? vertex.device() : background.device();
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    const auto i32 = torch::TensorOptions().dtype(torch::kInt32).device(dev);
    const auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(dev);
    const int64_t P = d.P;
    // every pixel is written by the composite kernel: no zero-fill (the reference: torch::full / zeros, extension_interface.cu:99-128)
    torch::Tensor radii = P == 0 ? torch::zeros({P}, i32) : torch::empty({P}, i32);
    torch::Tensor out_feature = P == 0 ? torch::zeros({d.C, H, W}, f32) : torch::empty({d.C, H, W}, f32);
    torch::Tensor depth, normal, contrib_sum, contrib_max;
    if (rich_info) {
        depth = P == 0 ? torch::zeros({H, W}, f32) : torch::empty({H, W}, f32);
        normal = P == 0 ? torch::zeros({3, H, W}, f32) : torch::empty({3, H, W}, f32);
        contrib_sum = torch::empty({P}, f32);
        contrib_max = torch::empty({P}, f32);
    } else {
        depth = torch::empty({0}, f32);
        normal = torch::empty({0}, f32);
        contrib_sum = torch::empty({0}, f32);
        contrib_max = torch::empty({0}, f32);
    }
    if (P == 0) {  // extension_interface.cu:130 -- zero outputs, empty state
        return py::make_tuple((int64_t)0, out_feature, radii, depth, normal, contrib_sum, contrib_max, torch::empty({0}, u8), torch::empty({0}, u8),
                              torch::empty({0}, u8), py::none());
    }

    c10::cuda::CUDAGuard guard(dev);
    void *stream = (void *)c10::cuda::getCurrentCUDAStream(dev.index()).stream();
    Packed p = pack(W, H, tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier, background_depth, background, vertex,
                    shs, feature, opacity, d, back_culling, rich_info, debug, primitive, exact);
    const size_t gbytes = ts2d_geometry_state_bytes((int32_t)P), ibytes = ts2d_image_state_bytes(W, H);
    torch::Tensor geometryBuffer = torch::empty({(int64_t)gbytes}, u8), imageBuffer = torch::empty({(int64_t)ibytes}, u8), binningBuffer;
    auto counters = std::make_shared<FrameCounters>();
    const ShapeKey key{(int)dev.index(), P, W, H, primitive};
    int64_t cap = -1;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_r_seen.find(key);
        if (it != g_r_seen.end() && !g_sync_forward && !debug) cap = std::max(it->second + it->second / 2, it->second + g_capacity_margin);
    }
    ts2d_forward_out out{out_feature.data_ptr<float>(), radii.data_ptr<int32_t>(), rich_info ? depth.data_ptr<float>() : nullptr,
                         rich_info ? normal.data_ptr<float>() : nullptr, rich_info ? contrib_sum.data_ptr<float>() : nullptr,
                         rich_info ? contrib_max.data_ptr<float>() : nullptr};
    int64_t R = -1;
    if (cap >= 0) {
        // one enqueue, no synchronisation in front of the binning / composite kernels
        const size_t bbytes = ts2d_binning_state_bytes(cap, W, H);
        binningBuffer = torch::empty({(int64_t)bbytes}, u8);
        check(ts2d_forward(&p.cam, &p.geom, &p.flags, radii.data_ptr<int32_t>(), geometryBuffer.data_ptr(), gbytes, binningBuffer.data_ptr(), bbytes,
                           imageBuffer.data_ptr(), ibytes, &out, counters->h, stream),
              "ts2d_forward");
        R = counters->num_rendered();  // waits for the scan only: the GPU is busy with the rest of the frame
        if (R > ts2d_binning_capacity(bbytes)) cap = -1;  // the guess was too small: the frame is invalid, render again on the same geometry state
    }
    if (cap < 0) {
        if (R < 0) {
            py::gil_scoped_release nogil;
            check(ts2d_forward_geometry(&p.cam, &p.geom, &p.flags, radii.data_ptr<int32_t>(), geometryBuffer.data_ptr(), gbytes, &R, stream),
                  "ts2d_forward_geometry");
        }
        const size_t bbytes = ts2d_binning_state_bytes(R, W, H);
        binningBuffer = torch::empty({(int64_t)bbytes}, u8);
        check(ts2d_forward_render(&p.cam, &p.geom, &p.flags, R, geometryBuffer.data_ptr(), binningBuffer.data_ptr(), bbytes, imageBuffer.data_ptr(), ibytes,
                                  &out, counters->h, stream),
              "ts2d_forward_render");
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        int64_t &seen = g_r_seen[key];
        seen = std::max(seen, R);
    }
    return py::make_tuple(R, out_feature, radii, depth, normal, contrib_sum, contrib_max, geometryBuffer, binningBuffer, imageBuffer, counters);
}

// rasterizeTrianglesBackward (extension_interface.cu:154-260) -> (dL_dvertex, dL_dcenter2D, dL_dshs, dL_dfeature, dL_dopacity)
py::tuple rasterize_triangles_backward(float tan_fovx, float tan_fovy, const torch::Tensor &viewmatrix, const torch::Tensor &projmatrix,
                                       const torch::Tensor &campos, int sh_degree, float gamma, float scale_modifier, float background_depth,
                                       const torch::Tensor &background, const torch::Tensor &vertex, const torch::Tensor &shs, const torch::Tensor &feature,
                                       const torch::Tensor &opacity, int64_t num_rendered, const torch::Tensor &radii, const torch::Tensor &geometryBuffer,
                                       const torch::Tensor &binningBuffer, const torch::Tensor &imageBuffer, const torch::Tensor &dL_dout_feature,
                                       const c10::optional<torch::Tensor> &dL_dout_depth, const c10::optional<torch::Tensor> &dL_dout_normal, bool rich_info,
                                       bool debug, const std::shared_ptr<FrameCounters> &counters, int primitive, bool exact)
{
    (void)num_rendered;  // the instance count lives in the geometry state; the capacity is a function of the binning state's size
    const Derived d = derive(vertex, shs, feature);
    const int64_t P = d.P;
    const int H = (int)dL_dout_feature.size(1), W = (int)dL_dout_feature.size(2);
    const torch::Tensor *all[] = {&viewmatrix, &projmatrix, &campos, &background, &vertex, &shs, &feature, &opacity, &radii, &geometryBuffer,
                                  &binningBuffer, &imageBuffer, &dL_dout_feature};
    for (const torch::Tensor *t : all)
        if (!t->is_contiguous()) throw std::runtime_error("input tensors must be contiguous");
    if (rich_info) {
        if (!dL_dout_depth.has_value() || !dL_dout_normal.has_value()) throw std::runtime_error("rich_info needs dL_dout_depth and dL_dout_normal");
        if (!dL_dout_depth->is_contiguous() || !dL_dout_normal->is_contiguous()) throw std::runtime_error("input tensors must be contiguous");
    }
    const torch::Device dev = vertex.device();
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    if (P == 0)
        return py::make_tuple(torch::zeros({P, 3, 3}, f32), torch::zeros({P, 2}, f32), torch::zeros({P, d.M, 3}, f32), torch::zeros({P, d.C}, f32),
                              torch::zeros({P, 1}, f32));
    require_cuda_f32("dL_dout_feature", dL_dout_feature);
    // every element is written by the per-triangle kernel: no zero-fill (the reference: torch::zeros, extension_interface.cu:229-233)
    torch::Tensor dL_dvertex = torch::empty({P, 3, 3}, f32), dL_dcenter2D = torch::empty({P, 2}, f32), dL_dshs = torch::empty({P, d.M, 3}, f32),
                  dL_dfeature = torch::empty({P, d.C}, f32), dL_dopacity = torch::empty({P, 1}, f32);

    c10::cuda::CUDAGuard guard(dev);
    void *stream = (void *)c10::cuda::getCurrentCUDAStream(dev.index()).stream();
    Packed p = pack(W, H, tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier, background_depth, background, vertex,
                    shs, feature, opacity, d, false, rich_info, debug, primitive, exact);
    const ShapeKey key{(int)dev.index(), P, W, H, primitive};
    // rows of the atomics-free gradient write-back (0 with the mirror kernels, which keep the reference's atomics)
    int64_t rows = -1, rows_cap = 0;
    if (exact || !(gamma >= 0.6f && gamma <= 64.0f)) {
        rows = rows_cap = 0;
    } else if (counters) {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_rows_seen.find(key);
        if (it != g_rows_seen.end() && !g_sync_backward && !debug) rows_cap = it->second + it->second / 4 + g_rows_margin;
    }
    if (rows < 0 && rows_cap == 0) {
        if (counters) {
            rows = rows_cap = counters->backward_rows();
        } else {  // a caller that built the arguments itself: one blocking read of the geometry state
            ts2d_frame_counters fc;
            py::gil_scoped_release nogil;
            check(ts2d_read_counters(geometryBuffer.data_ptr(), (int32_t)P, &fc, stream), "ts2d_read_counters");
            rows = rows_cap = fc.backward_rows;
        }
    }
    const size_t bbytes = (size_t)binningBuffer.numel();
    ts2d_loss_in loss{dL_dout_feature.data_ptr<float>(), rich_info ? fptr(*dL_dout_depth) : nullptr, rich_info ? fptr(*dL_dout_normal) : nullptr};
    ts2d_backward_out out{dL_dvertex.data_ptr<float>(), dL_dcenter2D.data_ptr<float>(), d.M > 0 ? dL_dshs.data_ptr<float>() : nullptr,
                          dL_dfeature.data_ptr<float>(), dL_dopacity.data_ptr<float>(), nullptr};
    const auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(dev);
    for (;;) {
        const size_t sbytes = ts2d_backward_scratch_bytes((int32_t)P, bbytes, rows_cap);
        torch::Tensor scratch = torch::empty({(int64_t)sbytes}, u8);
        if (rows >= 0) {  // row count known up front: the whole pass in one call
            check(ts2d_backward(&p.cam, &p.geom, &p.flags, radii.data_ptr<int32_t>(), geometryBuffer.data_ptr(), binningBuffer.data_ptr(), bbytes,
                                imageBuffer.data_ptr(), &loss, &out, scratch.data_ptr(), sbytes, stream),
                  "ts2d_backward");
            break;
        }
        check(ts2d_backward_composite(&p.cam, &p.geom, &p.flags, geometryBuffer.data_ptr(), binningBuffer.data_ptr(), bbytes, imageBuffer.data_ptr(), &loss,
                                      scratch.data_ptr(), sbytes, nullptr, stream),
              "ts2d_backward_composite");
        rows = counters->backward_rows();  // landed with the end of the forward pass: the composite above is already queued behind it
        if (rows > rows_cap) {             // rows beyond the guess were dropped: composite again with the exact size
            rows_cap = rows;
            continue;
        }
        check(ts2d_backward_geometry(&p.cam, &p.geom, &p.flags, radii.data_ptr<int32_t>(), geometryBuffer.data_ptr(), &out, scratch.data_ptr(),
                                     (size_t)64 * (size_t)P, stream),
              "ts2d_backward_geometry");
        break;
    }
    if (rows_cap > 0) {
        std::lock_guard<std::mutex> lk(g_mu);
        int64_t &seen = g_rows_seen[key];
        seen = std::max(seen, rows);
    }
    return py::make_tuple(dL_dvertex, dL_dcenter2D, dL_dshs, dL_dfeature, dL_dopacity);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.doc() = "pybind11 layer over the C ABI of libts2d.so (include/ts2d.h); mirrors R2D/ext.cpp:4-9";
    py::class_<FrameCounters, std::shared_ptr<FrameCounters>>(m, "FrameCounters")
        .def("num_rendered", &FrameCounters::num_rendered)
        .def("backward_rows", &FrameCounters::backward_rows);
    m.def("rasterize_triangles", &rasterize_triangles, py::arg("image_width"), py::arg("image_height"), py::arg("tan_fovx"), py::arg("tan_fovy"),
          py::arg("viewmatrix"), py::arg("projmatrix"), py::arg("campos"), py::arg("sh_degree"), py::arg("gamma"), py::arg("scale_modifier"),
          py::arg("background_depth"), py::arg("background"), py::arg("vertex"), py::arg("shs"), py::arg("feature"), py::arg("opacity"),
          py::arg("back_culling"), py::arg("rich_info"), py::arg("debug"), py::arg("primitive") = 0, py::arg("exact") = false);
    m.def("rasterize_triangles_backward", &rasterize_triangles_backward, py::arg("tan_fovx"), py::arg("tan_fovy"), py::arg("viewmatrix"),
          py::arg("projmatrix"), py::arg("campos"), py::arg("sh_degree"), py::arg("gamma"), py::arg("scale_modifier"), py::arg("background_depth"),
          py::arg("background"), py::arg("vertex"), py::arg("shs"), py::arg("feature"), py::arg("opacity"), py::arg("num_rendered"), py::arg("radii"),
          py::arg("geometryBuffer"), py::arg("binningBuffer"), py::arg("imageBuffer"), py::arg("dL_dout_feature"), py::arg("dL_dout_depth"),
          py::arg("dL_dout_normal"), py::arg("rich_info"), py::arg("debug"), py::arg("counters") = std::shared_ptr<FrameCounters>(),
          py::arg("primitive") = 0, py::arg("exact") = false);
    m.def("abi_version", []() { return ts2d_abi_version(); });
    m.def("counters_created", []() { return g_counters_created; });
    // knobs of the one-enqueue paths (tests force the overflow / the synchronous variants through these)
    m.def("configure", [](c10::optional<bool> sync_forward, c10::optional<bool> sync_backward, c10::optional<int64_t> capacity_margin,
                          c10::optional<int64_t> rows_margin) {
        std::lock_guard<std::mutex> lk(g_mu);
        if (sync_forward) g_sync_forward = *sync_forward;
        if (sync_backward) g_sync_backward = *sync_backward;
        if (capacity_margin) g_capacity_margin = *capacity_margin;
        if (rows_margin) g_rows_margin = *rows_margin;
        return py::make_tuple(g_sync_forward, g_sync_backward, g_capacity_margin, g_rows_margin);
    }, py::arg("sync_forward") = py::none(), py::arg("sync_backward") = py::none(), py::arg("capacity_margin") = py::none(), py::arg("rows_margin") = py::none());
    m.def("forget_shapes", []() {
        std::lock_guard<std::mutex> lk(g_mu);
        g_r_seen.clear();
        g_rows_seen.clear();
    });
    m.def("poison_shapes", [](int64_t r, int64_t rows) {  // tests: pretend every shape seen so far had this many instances / rows
        std::lock_guard<std::mutex> lk(g_mu);
        for (auto &kv : g_r_seen) kv.second = r;
        for (auto &kv : g_rows_seen) kv.second = rows;
    });
    m.def("shapes_seen", []() {
        std::lock_guard<std::mutex> lk(g_mu);
        int64_t r = 0, rows = 0;
        for (auto &kv : g_r_seen) r = std::max(r, kv.second);
        for (auto &kv : g_rows_seen) rows = std::max(rows, kv.second);
        return py::make_tuple((int64_t)g_r_seen.size(), r, (int64_t)g_rows_seen.size(), rows);
    });
}
