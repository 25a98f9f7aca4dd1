// The frame operator of the reference (DynFusion::operator(), src/dynfu/dyn_fusion.cpp:48-145) through the C++ adapter:
// frame 0 fuses the canonical surface rigidly, frame 1 estimates the warp field against a live frame and fuses the live
// depth through it.  The reference has no test of its frame operator; the post-conditions below are the solver tests'
// (warped canonical vertices land on the live ones, test/opt_optimisation_test.cpp:94) plus bit-equality of the first
// frame with a plain TsdfVolume::integrate.
#include "../../dynfu_b200/adapter/dynfu_adapter.hpp"
#include "mini_gtest.h"

class FrameTest {
public:
    void SetUp() {}
    void TearDown() {}
};

TEST_F(FrameTest, FrameOperatorFusesTracksAndFusesAgain) {
    DynFuParams p = DynFuParams::defaultParams();
    p.volume_dims[0] = p.volume_dims[1] = p.volume_dims[2] = 64;
    p.epsilon = 0.1f;   // dg_w = 0.3 m: every canonical vertex is supported, Warpfield::update adds no node
    p.node_step = 5;    // 125 nodes: the CPU oracle fits this frame to 0.32 mm with the same parameters
    p.solver.numIter = 8;
    p.solver.nonLinearIter = 4;
    p.solver.linearIter = 64;
    p.solver.earlyOut = false;
    const int rows = p.rows, cols = p.cols;
    // a slanted wall 1.7 .. 2.0 m in front of the camera
    std::vector<uint16_t> depth((size_t) rows * cols, 0);
    for (int y = 100; y < 380; ++y)
        for (int x = 160; x < 480; ++x) depth[(size_t) y * cols + x] = (uint16_t) (1700 + (x - 160) / 2 + (y - 100) / 3);
    // canonical vertices: every 12th pixel of the wall, back-projected into the volume frame (camera pose = identity,
    // volume_pose = translate(volume_pose_t): p_vol = p_cam - t)
    pcl::PointCloud<pcl::PointXYZ> cv_, lv_;
    pcl::PointCloud<pcl::Normal> cn_, ln_;
    const float shift[3] = {0.004f, -0.003f, 0.002f};
    for (int y = 106; y < 380; y += 12)
        for (int x = 166; x < 480; x += 12) {
            const float z = depth[(size_t) y * cols + x] * 1e-3f;
            const float px = (x - p.intr[2]) / p.intr[0] * z - p.volume_pose_t[0];
            const float py = (y - p.intr[3]) / p.intr[1] * z - p.volume_pose_t[1];
            const float pz = z - p.volume_pose_t[2];
            cv_.push_back(pcl::PointXYZ(px, py, pz));
            cn_.push_back(pcl::Normal(0.f, 0.f, -1.f));
            lv_.push_back(pcl::PointXYZ(px + shift[0], py + shift[1], pz + shift[2]));
            ln_.push_back(pcl::Normal(0.f, 0.f, -1.f));
        }
    ASSERT_NEAR(cv_.size() > 500 ? 1.0 : 0.0, 1.0, 0.0);

    DynFusion dynfu(p);
    const bool first = dynfu(depth.data());
    ASSERT_NEAR(first ? 1.0 : 0.0, 0.0, 0.0);  // "can't do more with the first frame" (:68)
    // frame 0 == TsdfVolume::integrate of the same distances
    const size_t nvox = (size_t) 64 * 64 * 64;
    std::vector<uint32_t> v0(nvox), ref(nvox), v1(nvox);
    dfu_adapter::cuda_check(cudaMemcpy(v0.data(), dynfu.tsdf().data(), nvox * 4, cudaMemcpyDeviceToHost), "download");
    {
        dfu_adapter::DevArray<uint16_t> d_depth, d_dists((size_t) rows * cols);
        d_depth.upload(depth.data(), depth.size());
        dfu_adapter::check(dfu_compute_dists(d_depth.p, cols * 2, d_dists.p, cols * 2, rows, cols, p.intr, nullptr), "dfu_compute_dists");
        kfusion::cuda::TsdfVolume vol(64, 64, 64);
        vol.setPose(p.volume_pose_t);
        vol.setTruncDist(p.tsdf_trunc_dist);
        vol.setMaxWeight(p.tsdf_max_weight);
        vol.integrate(d_dists.p, cols * 2, rows, cols, p.intr, nullptr);
        dfu_adapter::cuda_check(cudaMemcpy(ref.data(), vol.data(), nvox * 4, cudaMemcpyDeviceToHost), "download");
    }
    size_t touched = 0, diff = 0;
    for (size_t i = 0; i < nvox; ++i) {
        touched += v0[i] != 0;
        diff += v0[i] != ref[i];
    }
    ASSERT_NEAR((double) diff, 0.0, 0.0);
    ASSERT_NEAR(touched > 1000 ? 1.0 : 0.0, 1.0, 0.0);

    dynfu.init(cv_, cn_);
    ASSERT_NEAR(dynfu.getWarpfield()->getNodes().size() >= 8 ? 1.0 : 0.0, 1.0, 0.0);
    auto live = std::make_shared<dynfu::Frame>(1, lv_, ln_);
    const bool second = dynfu(depth.data(), live);
    ASSERT_NEAR(second ? 1.0 : 0.0, 1.0, 0.0);
    ASSERT_NEAR((double) dynfu.frameCounter(), 2.0, 0.0);
    // the warp field now carries the canonical vertices onto the live ones (opt_optimisation_test.cpp:94)
    auto wf = dynfu.getWarpfield();
    double worst = 0.0;
    for (size_t i = 0; i < cv_.size(); ++i) {
        auto r = wf->calcDQB(cv_[i])->transformVertex(cv_[i]);
        worst = std::fmax(worst, std::fabs(r.x - lv_[i].x));
        worst = std::fmax(worst, std::fabs(r.y - lv_[i].y));
        worst = std::fmax(worst, std::fabs(r.z - lv_[i].z));
    }
    ASSERT_NEAR(worst, 0.0, 1e-3);
    // and the second, warped integration raised the weights of voxels the first one had touched
    dfu_adapter::cuda_check(cudaMemcpy(v1.data(), dynfu.tsdf().data(), nvox * 4, cudaMemcpyDeviceToHost), "download");
    size_t twice = 0;
    for (size_t i = 0; i < nvox; ++i) twice += (v0[i] >> 16) == 1 && (v1[i] >> 16) == 2;
    ASSERT_NEAR(twice > 500 ? 1.0 : 0.0, 1.0, 0.0);
}

int main() { return RUN_ALL_TESTS(); }
