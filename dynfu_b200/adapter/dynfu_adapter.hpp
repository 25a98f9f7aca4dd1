// dynfu_adapter.hpp -- the reference's hot-path classes re-implemented on top of the C-ABI of
// include/dynfu_b200.h, so that code written against swarth100/dynfu (src/dynfu/dyn_fusion.cpp, the gtest
// cases of test/opt_optimisation_test.cpp) compiles and runs against the B200 library.
//
//   Node, DualQuaternion<T>      include/dynfu/utils/node.hpp, dual_quaternion.hpp   (host-side value types)
//   dynfu::Frame                 include/dynfu/utils/frame.hpp
//   Warpfield                    include/dynfu/warp_field.hpp:32-78
//   CombinedSolver(+Parameters)  include/dynfu/utils/opt_solver.hpp:19-110 (+ Opt's CombinedSolverParameters)
//   kfusion::cuda::TsdfVolume    include/kfusion/cuda/tsdf_volume.hpp:7-73 (create/clear/integrate/data)
//   DynFusion (+DynFuParams)     include/dynfu/dyn_fusion.hpp, src/dynfu/dyn_fusion.cpp:48-210 (the frame operator, hot-path steps)
//
// PCL / OpenCV / Boost are not required: when their headers are absent the minimal stand-ins below
// (pcl::PointXYZ, pcl::Normal, pcl::PointCloud, cv::Vec3f, cv::Affine3f) are used; define
// DYNFU_ADAPTER_USE_PCL to build against the real ones.  All heavy work happens on the GPU through the
// C-ABI; this header only marshals between the reference's containers and flat device arrays.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dynfu_b200.h"

#ifndef KNN
#define KNN 8  // include/dynfu/warp_field.hpp:27
#endif

#ifndef DYNFU_ADAPTER_USE_PCL
namespace pcl {
struct PointXYZ {
    float x = 0, y = 0, z = 0;
    PointXYZ() = default;
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct Normal {
    float normal_x = 0, normal_y = 0, normal_z = 0;
    Normal() = default;
    Normal(float x_, float y_, float z_) : normal_x(x_), normal_y(y_), normal_z(z_) {}
};
template <class T>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    std::vector<T> points;
    void push_back(const T& p) { points.push_back(p); }
    size_t size() const { return points.size(); }
    void clear() { points.clear(); }
    T& operator[](size_t i) { return points[i]; }
    const T& operator[](size_t i) const { return points[i]; }
    typename std::vector<T>::iterator begin() { return points.begin(); }
    typename std::vector<T>::iterator end() { return points.end(); }
    typename std::vector<T>::const_iterator begin() const { return points.begin(); }
    typename std::vector<T>::const_iterator end() const { return points.end(); }
};
}  // namespace pcl
namespace cv {
struct Vec3f {
    float v[3] = {0, 0, 0};
    Vec3f() = default;
    Vec3f(float a, float b, float c) : v{a, b, c} {}
    float operator[](int i) const { return v[i]; }
};
struct Affine3f {  // only carried around: the reference's solver never uses it (opt_solver.cpp:175-177 FIXME)
    float m[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    Affine3f() = default;
    Affine3f(const Vec3f&, const Vec3f& t) {
        m[9] = t[0]; m[10] = t[1]; m[11] = t[2];
    }
};
}  // namespace cv

#endif

namespace dfu_adapter {
inline void check(int rc, const char* what) {
    if (rc != DFU_OK) throw std::runtime_error(std::string(what) + ": " + dfu_last_error());
}
inline void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
// small RAII device array
template <class T>
struct DevArray {
    T* p = nullptr;
    size_t n = 0;
    DevArray() = default;
    explicit DevArray(size_t n_) { resize(n_); }
    DevArray(const DevArray&) = delete;
    DevArray& operator=(const DevArray&) = delete;
    ~DevArray() { cudaFree(p); }
    void resize(size_t n_) {
        if (n_ > n) {
            cudaFree(p);
            p = nullptr;
            cuda_check(cudaMalloc(&p, n_ * sizeof(T)), "cudaMalloc");
        }
        n = n_;
    }
    void upload(const T* h, size_t cnt) {
        resize(cnt);
        if (cnt) cuda_check(cudaMemcpy(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice), "upload");
    }
    void download(T* h, size_t cnt) const {
        if (cnt) cuda_check(cudaMemcpy(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost), "download");
    }
};
}  // namespace dfu_adapter

// ---------------------------------------------------------------------------------------------------------
// DualQuaternion<T>: value type (real wxyz, dual wxyz).  Construction follows dual_quaternion.hpp:38-67.
template <class T>
class DualQuaternion {
public:
    T q[8];
    DualQuaternion() : q{1, 0, 0, 0, 0, 0, 0, 0} {}
    explicit DualQuaternion(const float* f) {
        for (int i = 0; i < 8; ++i) q[i] = f[i];
    }
    // Euler angles (yaw, pitch, roll) + translation (dual_quaternion.hpp:48-67)
    DualQuaternion(T yaw, T pitch, T roll, T x, T y, T z) {
        T cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5), cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5);
        T cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5);
        T r[4] = {cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                  sy * cr * cp - cy * sr * sp};
        T n = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];  // boost::math::norm = squared norm (:31)
        for (int i = 0; i < 4; ++i) q[i] = r[i] / n;
        // dual = ((0,t) * real) * 0.5f (:44)
        const T a = 0, b = x, c = y, d = z, ar = q[0], br = q[1], cr_ = q[2], dr = q[3];
        q[4] = (a * ar - b * br - c * cr_ - d * dr) * T(0.5);
        q[5] = (a * br + b * ar + c * dr - d * cr_) * T(0.5);
        q[6] = (a * cr_ - b * dr + c * ar + d * br) * T(0.5);
        q[7] = (a * dr + b * cr_ - c * br + d * ar) * T(0.5);
    }
    // transformVertex (dual_quaternion.hpp:204-215), host convenience for single points
    pcl::PointXYZ transformVertex(pcl::PointXYZ v) const {
        const T rw = q[0], rx = q[1], ry = q[2], rz = q[3], dw = q[4], dx = q[5], dy = q[6], dz = q[7];
        const T cx = ry * v.z - rz * v.y + rw * v.x, cy = rz * v.x - rx * v.z + rw * v.y, cz = rx * v.y - ry * v.x + rw * v.z;
        const T ax = 2 * (ry * cz - rz * cy), ay = 2 * (rz * cx - rx * cz), az = 2 * (rx * cy - ry * cx);
        const T bx = 2 * (rw * dx - dw * rx + (ry * dz - rz * dy)), by = 2 * (rw * dy - dw * ry + (rz * dx - rx * dz)),
                bz = 2 * (rw * dz - dw * rz + (rx * dy - ry * dx));
        return pcl::PointXYZ(v.x + ax + bx, v.y + ay + by, v.z + az + bz);
    }
};

// Node (include/dynfu/utils/node.hpp:33-59)
class Node {
public:
    Node(pcl::PointXYZ position, std::shared_ptr<DualQuaternion<float>> transformation, float radialBasisWeight)
        : dg_v(position), dg_se3(std::move(transformation)), dg_w(radialBasisWeight) {}
    pcl::PointXYZ getPosition() { return dg_v; }
    std::shared_ptr<DualQuaternion<float>>& getTransformation() { return dg_se3; }
    void setTransformation(std::shared_ptr<DualQuaternion<float>> t) { dg_se3 = std::move(t); }
    float getRadialBasisWeight() { return dg_w; }

private:
    pcl::PointXYZ dg_v;
    std::shared_ptr<DualQuaternion<float>> dg_se3;
    float dg_w;
};

namespace dynfu {
// Frame (include/dynfu/utils/frame.hpp:15-33)
class Frame {
public:
    Frame(int id_, pcl::PointCloud<pcl::PointXYZ> v, pcl::PointCloud<pcl::Normal> n) : id(id_), vertices(std::move(v)), normals(std::move(n)) {}
    pcl::PointCloud<pcl::PointXYZ>& getVertices() { return vertices; }
    pcl::PointCloud<pcl::Normal>& getNormals() { return normals; }
    int getId() const { return id; }

private:
    int id;
    pcl::PointCloud<pcl::PointXYZ> vertices;
    pcl::PointCloud<pcl::Normal> normals;
};
}  // namespace dynfu

// ---------------------------------------------------------------------------------------------------------
// Warpfield (include/dynfu/warp_field.hpp:32-78).  Copies share the device handle AND the Node objects, like the
// reference's by-value copies share their shared_ptr<Node>s (opt_solver.cpp:5).
class Warpfield {
public:
    Warpfield() = default;

    // src/dynfu/warp_field.cpp:10-28
    void init(float epsilon_, std::vector<std::shared_ptr<Node>> nodes_) {
        epsilon = epsilon_;
        nodes = std::move(nodes_);
        if (!handle) {
            dfu_warpfield* h = nullptr;
            int dev = 0;
            cudaGetDevice(&dev);
            dfu_adapter::check(dfu_warpfield_create(&h, dev), "dfu_warpfield_create");
            handle = std::shared_ptr<dfu_warpfield>(h, [](dfu_warpfield* p) { dfu_warpfield_destroy(p); });
        }
        const int N = (int) nodes.size();
        std::vector<float> pos(3 * N), dq(8 * N), w(N);
        for (int i = 0; i < N; ++i) {
            const pcl::PointXYZ p = nodes[i]->getPosition();
            pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
            for (int k = 0; k < 8; ++k) dq[8 * i + k] = nodes[i]->getTransformation()->q[k];
            w[i] = nodes[i]->getRadialBasisWeight();
        }
        dfu_adapter::check(dfu_warpfield_init_host(handle.get(), epsilon, pos.data(), dq.data(), w.data(), N, nullptr),
                           "dfu_warpfield_init_host");
    }
    std::vector<std::shared_ptr<Node>> getNodes() { return nodes; }  // warp_field.cpp:32

    // warp_field.cpp:111-122
    std::vector<size_t> findNeighborsIndex(int numNeighbor, pcl::PointXYZ vertex) {
        if (numNeighbor != KNN) throw std::invalid_argument("k is fixed at KNN = 8");
        pushTransforms();
        const float q[3] = {vertex.x, vertex.y, vertex.z};
        dfu_adapter::DevArray<float> dq_(3);
        dfu_adapter::DevArray<int32_t> di(8);
        dq_.upload(q, 3);
        dfu_adapter::check(dfu_warpfield_knn(handle.get(), dq_.p, 1, di.p, nullptr, nullptr), "dfu_warpfield_knn");
        int32_t idx[8];
        di.download(idx, 8);
        std::vector<size_t> out;
        for (int k = 0; k < 8; ++k)
            if (idx[k] >= 0) out.push_back((size_t) idx[k]);  // n < 8 nodes -> shorter vector, like nanoflann
        return out;
    }
    std::vector<std::shared_ptr<Node>> findNeighbors(int numNeighbor, pcl::PointXYZ vertex) {  // :99-109
        std::vector<std::shared_ptr<Node>> out;
        for (size_t i : findNeighborsIndex(numNeighbor, vertex)) out.push_back(nodes[i]);
        return out;
    }
    // warp_field.cpp:127-148
    std::shared_ptr<DualQuaternion<float>> calcDQB(pcl::PointXYZ point, int blend_mode = DFU_BLEND_REF_COMPOSE) {
        pushTransforms();
        const float p[3] = {point.x, point.y, point.z};
        dfu_adapter::DevArray<float> dp(3), dout(8);
        dp.upload(p, 3);
        dfu_adapter::check(dfu_warpfield_blend(handle.get(), dp.p, 1, dout.p, blend_mode, nullptr), "dfu_warpfield_blend");
        float q[8];
        dout.download(q, 8);
        return std::make_shared<DualQuaternion<float>>(q);
    }
    // warp_field.cpp:150-171
    std::shared_ptr<dynfu::Frame> warpToLive(std::shared_ptr<dynfu::Frame> canonicalFrame, int blend_mode = DFU_BLEND_REF_COMPOSE) {
        pushTransforms();
        auto& V = canonicalFrame->getVertices();
        auto& Nn = canonicalFrame->getNormals();
        const int P = (int) V.size();
        std::vector<float> v(3 * P), n(3 * P);
        for (int i = 0; i < P; ++i) {
            v[3 * i] = V[i].x; v[3 * i + 1] = V[i].y; v[3 * i + 2] = V[i].z;
            n[3 * i] = Nn[i].normal_x; n[3 * i + 1] = Nn[i].normal_y; n[3 * i + 2] = Nn[i].normal_z;
        }
        dfu_adapter::DevArray<float> dv, dn;
        dv.upload(v.data(), v.size());
        dn.upload(n.data(), n.size());
        dfu_adapter::check(dfu_warpfield_warp(handle.get(), dv.p, dn.p, P, dv.p, dn.p, blend_mode, DFU_NORMAL_REF, nullptr),
                           "dfu_warpfield_warp");
        dv.download(v.data(), v.size());
        dn.download(n.data(), n.size());
        pcl::PointCloud<pcl::PointXYZ> wv;
        pcl::PointCloud<pcl::Normal> wn;
        for (int i = 0; i < P; ++i) {
            wv.push_back(pcl::PointXYZ(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
            wn.push_back(pcl::Normal(n[3 * i], n[3 * i + 1], n[3 * i + 2]));
        }
        return std::make_shared<dynfu::Frame>(0, wv, wn);
    }

    void addNode(std::shared_ptr<Node> newNode) { nodes.emplace_back(std::move(newNode)); }  // warp_field.cpp:30

    // warp_field.cpp:34-62
    pcl::PointCloud<pcl::PointXYZ>::Ptr getUnsupportedVertices(std::shared_ptr<dynfu::Frame> frame) {
        pushTransforms();
        auto& V = frame->getVertices();
        const int P = (int) V.size();
        pcl::PointCloud<pcl::PointXYZ>::Ptr out(new pcl::PointCloud<pcl::PointXYZ>);
        if (P == 0) return out;
        std::vector<float> v(3 * (size_t) P);
        for (int i = 0; i < P; ++i) { v[3 * i] = V[i].x; v[3 * i + 1] = V[i].y; v[3 * i + 2] = V[i].z; }
        dfu_adapter::DevArray<float> dv;
        dfu_adapter::DevArray<uint8_t> df((size_t) P);
        dv.upload(v.data(), v.size());
        dfu_adapter::check(dfu_warpfield_unsupported(handle.get(), dv.p, P, df.p, nullptr), "dfu_warpfield_unsupported");
        std::vector<uint8_t> flags((size_t) P);
        df.download(flags.data(), flags.size());
        for (int i = 0; i < P; ++i)
            if (flags[i]) out->push_back(V[i]);
        return out;
    }

    // warp_field.cpp:64-95: new nodes appear at the end of getNodes(), existing shared Nodes are kept
    void update(std::shared_ptr<dynfu::Frame> frame, int blend_mode = DFU_BLEND_REF_COMPOSE) {
        pushTransforms();
        auto& V = frame->getVertices();
        const int P = (int) V.size();
        if (P == 0) return;
        std::vector<float> v(3 * (size_t) P);
        for (int i = 0; i < P; ++i) { v[3 * i] = V[i].x; v[3 * i + 1] = V[i].y; v[3 * i + 2] = V[i].z; }
        dfu_adapter::DevArray<float> dv;
        dv.upload(v.data(), v.size());
        int n_uns = 0, n_new = 0;
        dfu_adapter::check(dfu_warpfield_update(handle.get(), dv.p, P, blend_mode, &n_uns, &n_new, nullptr), "dfu_warpfield_update");
        if (n_new == 0) return;
        const size_t N2 = nodes.size() + (size_t) n_new;
        std::vector<float> pos(3 * N2), dq(8 * N2), w(N2);
        dfu_adapter::check(dfu_warpfield_get_nodes_host(handle.get(), pos.data(), dq.data(), w.data(), nullptr), "dfu_warpfield_get_nodes_host");
        for (size_t i = nodes.size(); i < N2; ++i)
            addNode(std::make_shared<Node>(pcl::PointXYZ(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]),
                                           std::make_shared<DualQuaternion<float>>(&dq[8 * i]), w[i]));
    }

    // ---- not in the reference: the bridge between the shared host Nodes and the device copy ----
    dfu_warpfield* raw() const { return handle.get(); }
    // host Node transforms -> device (cheap: 32 B per node); called before every device operation because the
    // reference lets callers mutate Nodes through the shared_ptrs at any time
    void pushTransforms() {
        if (!handle) throw std::runtime_error("Warpfield used before init");
        std::vector<float> dq(8 * nodes.size());
        for (size_t i = 0; i < nodes.size(); ++i)
            for (int k = 0; k < 8; ++k) dq[8 * i + k] = nodes[i]->getTransformation()->q[k];
        dfu_adapter::check(dfu_warpfield_set_transforms_host(handle.get(), dq.data(), nullptr), "dfu_warpfield_set_transforms_host");
    }
    // device transforms -> host Nodes (after the solver's write-back)
    void pullTransforms() {
        std::vector<float> dq(8 * nodes.size());
        dfu_adapter::check(dfu_warpfield_get_nodes_host(handle.get(), nullptr, dq.data(), nullptr, nullptr), "dfu_warpfield_get_nodes_host");
        for (size_t i = 0; i < nodes.size(); ++i) nodes[i]->setTransformation(std::make_shared<DualQuaternion<float>>(&dq[8 * i]));
    }

private:
    float epsilon = 0.f;
    std::vector<std::shared_ptr<Node>> nodes;
    std::shared_ptr<dfu_warpfield> handle;
};

// Opt's CombinedSolverParameters, the fields the reference sets (dyn_fusion.cpp:183-189, opt_optimisation_test.cpp:38-44)
struct CombinedSolverParameters {
    int numIter = 24, nonLinearIter = 16, linearIter = 256;
    bool useOpt = true, useOptLM = false, earlyOut = true, optDoublePrecision = false;
    float pcgTolerance = 1e-6f;  // not in Opt's struct
};

// CombinedSolver (include/dynfu/utils/opt_solver.hpp:19-110)
class CombinedSolver {
public:
    CombinedSolver(Warpfield warpfield, CombinedSolverParameters params, float tukeyOffset, float psi_data, float lambda, float psi_reg)
        : m_warpfield(std::move(warpfield)) {
        dfu_solver_params p{params.numIter, params.nonLinearIter, params.linearIter, tukeyOffset, psi_data, lambda,
                            psi_reg, params.pcgTolerance, params.earlyOut ? 1 : 0};
        dfu_solver* h = nullptr;
        dfu_adapter::check(dfu_solver_create(&h, m_warpfield.raw(), &p), "dfu_solver_create");
        handle = std::shared_ptr<dfu_solver>(h, [](dfu_solver* s) { dfu_solver_destroy(s); });
    }
    // opt_solver.cpp:15-54
    void initializeProblemInstance(const std::shared_ptr<dynfu::Frame> canonicalFrame, const std::shared_ptr<dynfu::Frame> liveFrame,
                                   cv::Affine3f /*affine*/) {
        m_warpfield.pushTransforms();
        auto& C = canonicalFrame->getVertices();
        auto& L = liveFrame->getVertices();
        if (C.size() != L.size()) throw std::invalid_argument("canonical and live frames must pair up vertex by vertex");
        const int P = (int) C.size();
        std::vector<float> c(3 * P), l(3 * P);
        for (int i = 0; i < P; ++i) {
            c[3 * i] = C[i].x; c[3 * i + 1] = C[i].y; c[3 * i + 2] = C[i].z;
            l[3 * i] = L[i].x; l[3 * i + 1] = L[i].y; l[3 * i + 2] = L[i].z;
        }
        d_canon.upload(c.data(), c.size());
        d_live.upload(l.data(), l.size());
        dfu_adapter::check(dfu_solver_init_problem(handle.get(), d_canon.p, nullptr, d_live.p, nullptr, P, nullptr, nullptr),
                           "dfu_solver_init_problem");
    }
    // CombinedSolverBase::solveAll [Opt]; the result lands in the shared Nodes (opt_solver.cpp:270-285)
    void solveAll() {
        dfu_adapter::check(dfu_solver_solve_all(handle.get(), nullptr), "dfu_solver_solve_all");
        dfu_adapter::check(dfu_solver_get_stats_host(handle.get(), stats, nullptr), "dfu_solver_get_stats_host");
        m_warpfield.pullTransforms();
    }
    double finalCost() const { return stats[1]; }
    double initialCost() const { return stats[0]; }

private:
    Warpfield m_warpfield;
    std::shared_ptr<dfu_solver> handle;
    dfu_adapter::DevArray<float> d_canon, d_live;
    double stats[4] = {0, 0, 0, 0};
};

// ---------------------------------------------------------------------------------------------------------
namespace kfusion {
namespace cuda {
// TsdfVolume (include/kfusion/cuda/tsdf_volume.hpp:7-73): create / clear / integrate / data, device resident
class TsdfVolume {
public:
    explicit TsdfVolume(int dx, int dy, int dz) : dims{dx, dy, dz} {
        data_.resize((size_t) dx * dy * dz);
        setTruncDist(trunc_dist_);
        clear();
    }
    void setSize(float sx, float sy, float sz) { size_[0] = sx; size_[1] = sy; size_[2] = sz; setTruncDist(trunc_dist_); }
    void voxelSize(float vs[3]) const { for (int i = 0; i < 3; ++i) vs[i] = size_[i] / dims[i]; }
    void setTruncDist(float d) { float vs[3]; voxelSize(vs); trunc_dist_ = dfu_tsdf_trunc_dist(d, vs); }  // tsdf_volume.cpp:57-61
    float getTruncDist() const { return trunc_dist_; }
    void setMaxWeight(int w) { max_weight_ = w; }
    void setPose(const float vol2world_t[3]) { for (int i = 0; i < 3; ++i) pose_t[i] = vol2world_t[i]; }
    void clear() { dfu_adapter::check(dfu_tsdf_clear(data_.p, dims, 0, dims[2], nullptr), "dfu_tsdf_clear"); }  // :74-80
    // :82-93 with camera pose = identity rotation (the hot path's case); wf == nullptr integrates rigidly
    void integrate(const uint16_t* dists_dev, size_t pitch, int rows, int cols, const float intr[4], Warpfield* wf = nullptr) {
        float vs[3];
        voxelSize(vs);
        const float v2c[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, pose_t[0], pose_t[1], pose_t[2]};
        if (wf) wf->pushTransforms();
        dfu_adapter::check(dfu_tsdf_integrate(data_.p, dims, vs, trunc_dist_, max_weight_, v2c, intr, dists_dev, pitch, rows, cols,
                                              wf ? wf->raw() : nullptr, DFU_BLEND_REF_COMPOSE, 0, dims[2], nullptr),
                           "dfu_tsdf_integrate");
    }
    // :95-129 with camera pose = identity rotation: depth (u16 mm) + normals (float4) of the fused model, device images
    void setRaycastStepFactor(float f) { raycast_step_factor_ = f; }
    void setGradientDeltaFactor(float f) { gradient_delta_factor_ = f; }
    void raycast(const float intr[4], int rows, int cols, uint16_t* depth_dev, size_t depth_pitch, float* normals4_dev, size_t normals_pitch) {
        float vs[3];
        voxelSize(vs);
        // cam2vol = pose_.inv() * camera_pose = translate(-pose_t); its rotation (and inverse) is the identity
        const float c2v[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, -pose_t[0], -pose_t[1], -pose_t[2]};
        const float rinv[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        dfu_adapter::check(dfu_tsdf_raycast(data_.p, dims, vs, trunc_dist_, c2v, rinv, intr, rows, cols, raycast_step_factor_,
                                            gradient_delta_factor_, nullptr, 0, depth_dev, depth_pitch, normals4_dev, normals_pitch, nullptr),
                           "dfu_tsdf_raycast");
    }
    uint32_t* data() { return data_.p; }
    int dims[3];

private:
    dfu_adapter::DevArray<uint32_t> data_;
    float size_[3] = {3.f, 3.f, 3.f};
    float pose_t[3] = {0.f, 0.f, 0.f};
    float trunc_dist_ = 0.03f;  // tsdf_volume.cpp:21-22
    int max_weight_ = 128;
    float raycast_step_factor_ = 0.75f, gradient_delta_factor_ = 0.75f;  // tsdf_volume.cpp:25-26
};
}  // namespace cuda
}  // namespace kfusion

// DynFusion::findCorrespondingFrame (src/dynfu/dyn_fusion.cpp:212-242): for every live vertex the nearest canonical
// vertex (and its normal), as a new Frame.  The reference builds a nanoflann KD-tree per call; this is one C-ABI call.
inline std::shared_ptr<dynfu::Frame> findCorrespondingFrame(pcl::PointCloud<pcl::PointXYZ> canonicalVertices,
                                                            pcl::PointCloud<pcl::Normal> canonicalNormals,
                                                            pcl::PointCloud<pcl::PointXYZ> liveVertices) {
    const int Pc = (int) canonicalVertices.size(), Pl = (int) liveVertices.size();
    if (Pc == 0) throw std::runtime_error("findCorrespondingFrame: empty canonical cloud");  // nanoflann throws too
    std::vector<float> c(3 * (size_t) Pc), cn(3 * (size_t) Pc), l(3 * (size_t) Pl);
    for (int i = 0; i < Pc; ++i) {
        c[3 * i] = canonicalVertices[i].x; c[3 * i + 1] = canonicalVertices[i].y; c[3 * i + 2] = canonicalVertices[i].z;
        cn[3 * i] = canonicalNormals[i].normal_x; cn[3 * i + 1] = canonicalNormals[i].normal_y; cn[3 * i + 2] = canonicalNormals[i].normal_z;
    }
    for (int i = 0; i < Pl; ++i) { l[3 * i] = liveVertices[i].x; l[3 * i + 1] = liveVertices[i].y; l[3 * i + 2] = liveVertices[i].z; }
    dfu_adapter::DevArray<float> dc, dcn, dl, dov((size_t) 3 * Pl + 1), don((size_t) 3 * Pl + 1);
    dc.upload(c.data(), c.size());
    dcn.upload(cn.data(), cn.size());
    dl.upload(l.data(), l.size());
    dfu_pointindex* pi = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    dfu_adapter::check(dfu_pointindex_create(&pi, dev), "dfu_pointindex_create");
    const int rc = dfu_find_corresponding(pi, dc.p, dcn.p, Pc, dl.p, Pl, dov.p, don.p, nullptr, nullptr);
    if (rc == DFU_OK) cudaStreamSynchronize(nullptr);
    dfu_pointindex_destroy(pi);
    dfu_adapter::check(rc, "dfu_find_corresponding");
    std::vector<float> ov(3 * (size_t) Pl), on(3 * (size_t) Pl);
    dov.download(ov.data(), ov.size());
    don.download(on.data(), on.size());
    pcl::PointCloud<pcl::PointXYZ> rv;
    pcl::PointCloud<pcl::Normal> rn;
    for (int i = 0; i < Pl; ++i) {
        rv.push_back(pcl::PointXYZ(ov[3 * i], ov[3 * i + 1], ov[3 * i + 2]));
        rn.push_back(pcl::Normal(on[3 * i], on[3 * i + 1], on[3 * i + 2]));
    }
    return std::make_shared<dynfu::Frame>(0, rv, rn);
}

// ---------------------------------------------------------------------------------------------------------
// DynFusion: the frame operator (src/dynfu/dyn_fusion.cpp:48-145) with its hot-path steps, and the members it calls
// (init :147-168, initCanonicalFrame :170-175, addLiveFrame :177-180, warpCanonicalToLiveOpt :182-210).
// DynFuParams carries the fields of DynFuParams::defaultParams (:6-31) plus the KinFuParams fields the path reads
// (src/kfusion/kinfu.cpp:10-44).  Left to the caller, as rows outside the path: the bilateral
// filter / depth truncation / ICP of :57-66,100-105 (ICP "not being done yet" in the reference either) and the
// marching-cubes extraction of the live vertices (:117-134) -- the live frame is an argument instead (it can come
// from dfu_compute_points_normals + dfu_compact_points).  Where the reference clears the volume and re-integrates
// rigidly just to run marching cubes on it (:113-116, "FIXME ... shouldn't have to do this"), this operator does the
// non-rigid surface fusion the reference's README lists as the next step: the live depth is integrated into the
// canonical volume through the warp field.
struct DynFuParams {
    int rows = 480, cols = 640;                                   // kinfu.cpp:16-17
    float intr[4] = {525.f, 525.f, 319.5f, 239.5f};               // :18 (fx, fy, cx, cy)
    int volume_dims[3] = {128, 128, 128};                         // dyn_fusion.cpp:10
    float volume_size[3] = {3.f, 3.f, 3.f};                       // kinfu.cpp:21
    float volume_pose_t[3] = {-1.5f, -1.5f, 0.5f};                // :22 translate(-size/2, -size/2, 0.5)
    float tsdf_trunc_dist = 0.04f;                                // :34
    int tsdf_max_weight = 64;                                     // :35
    float tukeyOffset = 4.652f, lambda = 200.f, psi_data = 0.01f, psi_reg = 1e-4f;  // dyn_fusion.cpp:13-20
    float epsilon = 0.1f;                                         // :28
    int node_step = 128;                                          // :150
    CombinedSolverParameters solver;                              // :183-189 (numIter 24, nonLinearIter 16, linearIter 256, earlyOut)
    static DynFuParams defaultParams() { return DynFuParams(); }
};

class DynFusion {
public:
    explicit DynFusion(const DynFuParams& p)
        : dynfuParams(p), volume_(p.volume_dims[0], p.volume_dims[1], p.volume_dims[2]),
          d_depth((size_t) p.rows * p.cols), d_dists((size_t) p.rows * p.cols) {
        volume_.setSize(p.volume_size[0], p.volume_size[1], p.volume_size[2]);
        volume_.setPose(p.volume_pose_t);
        volume_.setTruncDist(p.tsdf_trunc_dist);
        volume_.setMaxWeight(p.tsdf_max_weight);
        volume_.clear();
    }
    DynFuParams& params() { return dynfuParams; }

    // :48-145.  depth: HOST image, u16 millimetres, rows x cols.  Returns false for the first frame (the canonical one),
    // true afterwards, like the reference.
    bool operator()(const uint16_t* depth_mm, std::shared_ptr<dynfu::Frame> live = nullptr) {
        const DynFuParams& p = dynfuParams;
        d_depth.upload(depth_mm, (size_t) p.rows * p.cols);
        dfu_adapter::check(dfu_compute_dists(d_depth.p, (size_t) p.cols * 2, d_dists.p, (size_t) p.cols * 2, p.rows, p.cols, p.intr, nullptr),
                           "dfu_compute_dists");                                                         // :55
        if (frame_counter_ == 0) {
            volume_.integrate(d_dists.p, (size_t) p.cols * 2, p.rows, p.cols, p.intr, nullptr);          // :70
            return ++frame_counter_, false;
        }
        if (live && warpfield) {
            liveFrame = live;                                                                            // :137
            warpCanonicalToLiveOpt(cv::Affine3f());                                                      // :140
            warpfield->update(getCanonicalWarpedToLive());                                               // :142
        }
        volume_.integrate(d_dists.p, (size_t) p.cols * 2, p.rows, p.cols, p.intr, warpfield.get());      // (non-rigid fusion)
        return ++frame_counter_, true;
    }

    // :147-168: canonical frame, one deformation node every node_step-th vertex, dg_w = 3 epsilon
    void init(pcl::PointCloud<pcl::PointXYZ>& canonicalVertices, pcl::PointCloud<pcl::Normal>& canonicalNormals) {
        initCanonicalFrame(canonicalVertices, canonicalNormals);
        std::vector<std::shared_ptr<Node>> deformationNodes;
        for (size_t i = 0; i < canonicalVertices.size(); i += (size_t) dynfuParams.node_step)
            deformationNodes.push_back(std::make_shared<Node>(canonicalVertices[i],
                                                              std::make_shared<DualQuaternion<float>>(0.f, 0.f, 0.f, 0.f, 0.f, 0.f),
                                                              3 * dynfuParams.epsilon));
        warpfield = std::make_shared<Warpfield>();
        warpfield->init(dynfuParams.epsilon, deformationNodes);
    }
    void initCanonicalFrame(pcl::PointCloud<pcl::PointXYZ>& vertices, pcl::PointCloud<pcl::Normal>& normals) {  // :170-175
        canonicalFrame = std::make_shared<dynfu::Frame>(0, vertices, normals);
        canonicalFrameWarpedToLive = std::make_shared<dynfu::Frame>(0, vertices, normals);
    }
    void addLiveFrame(int frameID, pcl::PointCloud<pcl::PointXYZ>& vertices, pcl::PointCloud<pcl::Normal>& normals) {  // :177-180
        liveFrame = std::make_shared<dynfu::Frame>(frameID, vertices, normals);
    }
    // :182-210 (the solver parameters come from DynFuParams::solver instead of being hard-coded)
    void warpCanonicalToLiveOpt(cv::Affine3f affine) {
        CombinedSolver combinedSolver(*warpfield, dynfuParams.solver, dynfuParams.tukeyOffset, dynfuParams.psi_data, dynfuParams.lambda,
                                      dynfuParams.psi_reg);
        canonicalFrameWarpedToLive = warpfield->warpToLive(canonicalFrame);
        std::shared_ptr<dynfu::Frame> correspondingCanonicalFrame = findCorrespondingFrame(
            canonicalFrameWarpedToLive->getVertices(), canonicalFrameWarpedToLive->getNormals(), liveFrame->getVertices());
        combinedSolver.initializeProblemInstance(correspondingCanonicalFrame, liveFrame, affine);
        combinedSolver.solveAll();
    }
    std::shared_ptr<dynfu::Frame> getCanonicalWarpedToLive() { return canonicalFrameWarpedToLive; }
    std::shared_ptr<Warpfield> getWarpfield() { return warpfield; }
    kfusion::cuda::TsdfVolume& tsdf() { return volume_; }
    int frameCounter() const { return frame_counter_; }

private:
    DynFuParams dynfuParams;
    kfusion::cuda::TsdfVolume volume_;
    dfu_adapter::DevArray<uint16_t> d_depth, d_dists;
    std::shared_ptr<Warpfield> warpfield;
    std::shared_ptr<dynfu::Frame> canonicalFrame, canonicalFrameWarpedToLive, liveFrame;
    int frame_counter_ = 0;
};
