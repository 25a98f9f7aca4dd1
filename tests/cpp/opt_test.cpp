// The reference's solver tests (test/opt_optimisation_test.cpp:212-698), written against the SAME class names
// and call sequence, running on dynfu_b200 through the adapter header (dynfu_b200/adapter/dynfu_adapter.hpp).
// Fixture values are the reference's; the assertions are its post-conditions (warped vertex within 1e-3, :94).
#include "../../dynfu_b200/adapter/dynfu_adapter.hpp"
#include "mini_gtest.h"

class OptTest {
public:
    void SetUp() {
        params.numIter = 32;  // :38-44
        params.nonLinearIter = 16;
        params.linearIter = 256;
        params.useOpt = false;
        params.useOptLM = true;
        params.earlyOut = true;
        params.optDoublePrecision = true;
        auto dg_se3 = [] { return std::make_shared<DualQuaternion<float>>(0.f, 0.f, 0.f, 0.f, 0.f, 0.f); };
        float dg_w = 2;  // :53
        const float g1[8][3] = {{3, 1, -1}, {1, 1, 1}, {-1, 2, 3}, {-1, -1, 1}, {-2, -1, -1}, {2, -1, -3}, {-1, 1, -1}, {2, 1, 1}};
        const float g2[10][3] = {{10, 10, 10}, {9, 11.1f, 10}, {10, 9, 10}, {10, 12, 9}, {9, 11, 10}, {12, 10, 9}, {9, 9, 12},
                                 {10.5f, 9, 9}, {10.5f, 12, 12}, {11, 11, 10.9f}};
        for (auto& p : g1) nodesGroup1.push_back(std::make_shared<Node>(pcl::PointXYZ(p[0], p[1], p[2]), dg_se3(), dg_w));
        for (auto& p : g2) nodesGroup2.push_back(std::make_shared<Node>(pcl::PointXYZ(p[0], p[1], p[2]), dg_se3(), dg_w));
        allNodes = nodesGroup1;
        allNodes.insert(allNodes.end(), nodesGroup2.begin(), nodesGroup2.end());
    }
    void TearDown() {}

    CombinedSolverParameters params;
    float maxError = 1e-3;
    std::vector<std::shared_ptr<Node>> nodesGroup1, nodesGroup2, allNodes;
    Warpfield warpfield;
    cv::Affine3f affine = cv::Affine3f(cv::Vec3f(0, 0, 0), cv::Vec3f(0, 0, 0));
    float epsilon_dynfu = 0.0015;
    float tukeyOffset = 4.652;
    float psi_data = 1e-2;
    float lambda = 0.f;
    float psi_reg = 1e-4;
    pcl::PointCloud<pcl::PointXYZ> sourceVertices, targetVertices;
    pcl::PointCloud<pcl::Normal> sourceNormals, targetNormals;
    std::shared_ptr<dynfu::Frame> canonicalFrame, canonicalFrameWarpedToLive, liveFrame;

    void diag(pcl::PointCloud<pcl::PointXYZ>& v, pcl::PointCloud<pcl::Normal>& n, std::initializer_list<float> vals) {
        v.clear();
        n.clear();
        for (float a : vals) {
            v.push_back(pcl::PointXYZ(a, a, a));
            n.push_back(pcl::Normal(1, 1, 1));
        }
    }
    // the loop every reference test ends with: calcDQB(v).transformVertex(v) ~= target
    void expectWarped(int& failures_, std::shared_ptr<dynfu::Frame> from, std::shared_ptr<dynfu::Frame> to) {
        int j = 0;
        for (auto vertex : from->getVertices()) {
            auto totalTransformation = warpfield.calcDQB(vertex);
            auto result = totalTransformation->transformVertex(vertex);
            ASSERT_NEAR(result.x, to->getVertices()[j].x, maxError);
            ASSERT_NEAR(result.y, to->getVertices()[j].y, maxError);
            ASSERT_NEAR(result.z, to->getVertices()[j].z, maxError);
            j++;
        }
    }
    void solve(std::shared_ptr<dynfu::Frame> from, std::shared_ptr<dynfu::Frame> to) {
        CombinedSolver combinedSolver(warpfield, params, tukeyOffset, psi_data, lambda, psi_reg);
        combinedSolver.initializeProblemInstance(from, to, affine);
        combinedSolver.solveAll();
    }
};

TEST_F(OptTest, SingleVertexOneGroupOfDeformationNodesTest) {  // :212
    warpfield.init(epsilon_dynfu, nodesGroup1);
    sourceVertices.push_back(pcl::PointXYZ(0, 0.04, 0));
    sourceNormals.push_back(pcl::Normal(1, 1, 1));
    canonicalFrameWarpedToLive = std::make_shared<dynfu::Frame>(0, sourceVertices, sourceNormals);
    targetVertices.push_back(pcl::PointXYZ(0.01, 0.03, 0));
    targetNormals.push_back(pcl::Normal(1, 1, 1));
    liveFrame = std::make_shared<dynfu::Frame>(1, targetVertices, targetNormals);
    solve(canonicalFrameWarpedToLive, liveFrame);
    expectWarped(failures_, canonicalFrameWarpedToLive, liveFrame);
}

TEST_F(OptTest, TwoVerticesOneNotMovingOneGroupOfDeformationNodesTest) {  // :243
    warpfield.init(epsilon_dynfu, allNodes);
    sourceVertices.push_back(pcl::PointXYZ(0, 0.05, 1));
    sourceVertices.push_back(pcl::PointXYZ(2, 2, 2));
    targetVertices.push_back(pcl::PointXYZ(0.01, 0.04, 1.01));
    targetVertices.push_back(pcl::PointXYZ(2, 2, 2));
    for (int i = 0; i < 2; ++i) {
        sourceNormals.push_back(pcl::Normal(1, 1, 1));
        targetNormals.push_back(pcl::Normal(1, 1, 1));
    }
    canonicalFrameWarpedToLive = std::make_shared<dynfu::Frame>(0, sourceVertices, sourceNormals);
    liveFrame = std::make_shared<dynfu::Frame>(1, targetVertices, targetNormals);
    solve(canonicalFrameWarpedToLive, liveFrame);
    expectWarped(failures_, canonicalFrameWarpedToLive, liveFrame);
}

TEST_F(OptTest, MultipleVerticesOneGroupOfDeformationNodesTest) {  // :280
    warpfield.init(epsilon_dynfu, nodesGroup1);
    diag(sourceVertices, sourceNormals, {-3, -2, 0.01f, 2, 3});
    diag(targetVertices, targetNormals, {-2.99f, -1.99f, 0.02f, 2.01f, 3.01f});
    canonicalFrameWarpedToLive = std::make_shared<dynfu::Frame>(0, sourceVertices, sourceNormals);
    liveFrame = std::make_shared<dynfu::Frame>(1, targetVertices, targetNormals);
    solve(canonicalFrameWarpedToLive, liveFrame);
    expectWarped(failures_, canonicalFrameWarpedToLive, liveFrame);
}

TEST_F(OptTest, OneGroupOfVerticesTwoGroupsOfDeformationNodes) {  // :329
    warpfield.init(epsilon_dynfu, allNodes);
    diag(sourceVertices, sourceNormals, {-3, -2, 0.01f, 2, 3});
    diag(targetVertices, targetNormals, {-2.99f, -1.99f, 0.02f, 2.01f, 3.01f});
    canonicalFrameWarpedToLive = std::make_shared<dynfu::Frame>(0, sourceVertices, sourceNormals);
    liveFrame = std::make_shared<dynfu::Frame>(1, targetVertices, targetNormals);
    solve(canonicalFrameWarpedToLive, liveFrame);
    expectWarped(failures_, canonicalFrameWarpedToLive, liveFrame);
}

TEST_F(OptTest, TwoGroupsOfVerticesTwoGroupsOfDeformationNodes) {  // :378
    warpfield.init(epsilon_dynfu, allNodes);
    diag(sourceVertices, sourceNormals, {-3, -2, 0.01f, 2, 3, 12, 11, 10, 10.5f, 11.5f});
    diag(targetVertices, targetNormals, {-2.99f, -1.99f, 0.02f, 2.01f, 3.01f, 11.99f, 10.99f, 9.99f, 10.51f, 11.49f});
    canonicalFrameWarpedToLive = std::make_shared<dynfu::Frame>(0, sourceVertices, sourceNormals);
    liveFrame = std::make_shared<dynfu::Frame>(1, targetVertices, targetNormals);
    solve(canonicalFrameWarpedToLive, liveFrame);
    expectWarped(failures_, canonicalFrameWarpedToLive, liveFrame);
}

TEST_F(OptTest, MultipleVerticesOneGroupOfDeformationNodesWarpTwiceTest) {  // :454
    warpfield.init(epsilon_dynfu, nodesGroup1);
    diag(sourceVertices, sourceNormals, {-3, -2, 0.04f, 2, 3});
    diag(targetVertices, targetNormals, {-2.99f, -1.99f, 0.05f, 2.01f, 3.01f});
    canonicalFrame = std::make_shared<dynfu::Frame>(0, sourceVertices, sourceNormals);
    liveFrame = std::make_shared<dynfu::Frame>(1, targetVertices, targetNormals);
    solve(canonicalFrame, liveFrame);
    expectWarped(failures_, canonicalFrame, liveFrame);
    if (failures_) return;
    std::shared_ptr<dynfu::Frame> warped = warpfield.warpToLive(canonicalFrame);
    diag(targetVertices, targetNormals, {-2.98f, -1.98f, 0.06f, 2.02f, 3.02f});
    auto nextLiveFrame = std::make_shared<dynfu::Frame>(2, targetVertices, targetNormals);
    solve(warped, nextLiveFrame);
    expectWarped(failures_, canonicalFrame, nextLiveFrame);  // checked on the ORIGINAL canonical vertices (:518-527)
}

TEST_F(OptTest, MultipleVerticesOneGroupOfDeformationNodesWarpThriceTest) {  // :530
    warpfield.init(epsilon_dynfu, nodesGroup1);
    diag(sourceVertices, sourceNormals, {-3, -2, 0.04f, 2, 3});
    diag(targetVertices, targetNormals, {-2.99f, -1.99f, 0.05f, 2.01f, 3.01f});
    canonicalFrame = std::make_shared<dynfu::Frame>(0, sourceVertices, sourceNormals);
    liveFrame = std::make_shared<dynfu::Frame>(1, targetVertices, targetNormals);
    solve(canonicalFrame, liveFrame);
    std::shared_ptr<dynfu::Frame> warped1 = warpfield.warpToLive(canonicalFrame);
    diag(targetVertices, targetNormals, {-2.98f, -1.98f, 0.06f, 2.02f, 3.02f});
    auto nextLiveFrame = std::make_shared<dynfu::Frame>(2, targetVertices, targetNormals);
    solve(warped1, nextLiveFrame);
    expectWarped(failures_, canonicalFrame, nextLiveFrame);
    if (failures_) return;
    std::shared_ptr<dynfu::Frame> warped2 = warpfield.warpToLive(warped1);  // :586
    diag(targetVertices, targetNormals, {-2.96f, -1.96f, 0.09f, 2.04f, 3.05f});
    auto nextNextLiveFrame = std::make_shared<dynfu::Frame>(3, targetVertices, targetNormals);
    solve(warped2, nextNextLiveFrame);
    expectWarped(failures_, warped1, nextNextLiveFrame);  // iterates the ONCE-warped frame (:620-629)
}

TEST_F(OptTest, MultipleVerticesOneGroupOfDeformationNodesWarpAndReverseTest) {  // :632
    warpfield.init(epsilon_dynfu, nodesGroup1);
    diag(sourceVertices, sourceNormals, {-3, -2, 0.04f, 2, 3});
    diag(targetVertices, targetNormals, {-2.99f, -1.99f, 0.05f, 2.01f, 3.01f});
    canonicalFrameWarpedToLive = std::make_shared<dynfu::Frame>(0, sourceVertices, sourceNormals);
    liveFrame = std::make_shared<dynfu::Frame>(1, targetVertices, targetNormals);
    solve(canonicalFrameWarpedToLive, liveFrame);
    expectWarped(failures_, canonicalFrameWarpedToLive, liveFrame);
    if (failures_) return;
    std::swap(canonicalFrameWarpedToLive, liveFrame);
    solve(canonicalFrameWarpedToLive, liveFrame);
    expectWarped(failures_, liveFrame, liveFrame);  // :688-697: the net warp is ~ identity
}

// kfusion::cuda::TsdfVolume through the adapter: an identity-transform warp field integrates like no warp field
class TsdfTest {
public:
    void SetUp() {}
    void TearDown() {}
};
TEST_F(TsdfTest, IdentityWarpEqualsRigidIntegrate) {
    const int rows = 480, cols = 640, dim = 64;
    const float intr[4] = {525.f, 525.f, 319.5f, 239.5f};
    std::vector<uint16_t> depth((size_t) rows * cols, 0);
    for (int y = 100; y < 380; ++y)
        for (int x = 160; x < 480; ++x) depth[(size_t) y * cols + x] = 1800 + (uint16_t) ((x + y) % 50);
    dfu_adapter::DevArray<uint16_t> d_depth, d_dists((size_t) rows * cols);
    d_depth.upload(depth.data(), depth.size());
    dfu_adapter::check(dfu_compute_dists(d_depth.p, cols * 2, d_dists.p, cols * 2, rows, cols, intr, nullptr), "dfu_compute_dists");
    const float pose[3] = {-1.5f, -1.5f, 0.5f};
    kfusion::cuda::TsdfVolume a(dim, dim, dim), b(dim, dim, dim);
    a.setPose(pose);
    b.setPose(pose);
    a.setTruncDist(0.04f);
    b.setTruncDist(0.04f);
    std::vector<std::shared_ptr<Node>> nodes;
    for (int i = 0; i < 64; ++i)
        nodes.push_back(std::make_shared<Node>(pcl::PointXYZ(1.0f + 0.017f * i, 1.5f + 0.011f * (i % 7), 1.3f + 0.003f * i),
                                               std::make_shared<DualQuaternion<float>>(0.f, 0.f, 0.f, 0.f, 0.f, 0.f), 0.05f));
    Warpfield wf;
    wf.init(0.025f, nodes);
    a.integrate(d_dists.p, cols * 2, rows, cols, intr, nullptr);
    b.integrate(d_dists.p, cols * 2, rows, cols, intr, &wf);
    std::vector<uint32_t> ha((size_t) dim * dim * dim), hb(ha.size());
    dfu_adapter::cuda_check(cudaMemcpy(ha.data(), a.data(), ha.size() * 4, cudaMemcpyDeviceToHost), "download");
    dfu_adapter::cuda_check(cudaMemcpy(hb.data(), b.data(), hb.size() * 4, cudaMemcpyDeviceToHost), "download");
    size_t touched = 0, diff = 0;
    for (size_t i = 0; i < ha.size(); ++i) {
        touched += ha[i] != 0;
        diff += ha[i] != hb[i];
    }
    ASSERT_NEAR((double) diff, 0.0, 0.0);
    ASSERT_NEAR(touched > 1000 ? 1.0 : 0.0, 1.0, 0.0);
    // TsdfVolume::raycast of the fused slab gives back (about) the depth it was fused from
    dfu_adapter::DevArray<uint16_t> d_ray((size_t) rows * cols);
    dfu_adapter::DevArray<float> d_nrm((size_t) rows * cols * 4);
    a.raycast(intr, rows, cols, d_ray.p, cols * 2, d_nrm.p, cols * 16);
    std::vector<uint16_t> ray((size_t) rows * cols);
    d_ray.download(ray.data(), ray.size());
    long hits = 0, close = 0;
    for (int y = 120; y < 360; ++y)
        for (int x = 180; x < 460; ++x) {
            const size_t i = (size_t) y * cols + x;
            if (!ray[i]) continue;
            ++hits;
            close += std::abs((int) ray[i] - (int) depth[i]) <= 60;  // one 47 mm voxel at 64^3 + the ramp
        }
    ASSERT_NEAR(hits > 20000 ? 1.0 : 0.0, 1.0, 0.0);
    ASSERT_NEAR((double) close / (double) hits > 0.9 ? 1.0 : 0.0, 1.0, 0.0);
}

// ---- the rows around the solver, through the same class interface ------------------------------------------------
class UpdateTest {
public:
    void SetUp() {}
    void TearDown() {}
};

// Warpfield::getUnsupportedVertices / update (src/dynfu/warp_field.cpp:34-95) with hand-checkable numbers
TEST_F(UpdateTest, UnsupportedVerticesBecomeNodes) {
    std::vector<std::shared_ptr<Node>> nodes;
    for (int i = 0; i < 9; ++i)  // a 3x3 patch of nodes, 10 cm apart, radius 0.15
        nodes.push_back(std::make_shared<Node>(pcl::PointXYZ(1.0f + 0.1f * (i % 3), 1.0f + 0.1f * (i / 3), 1.0f),
                                               std::make_shared<DualQuaternion<float>>(0.f, 0.f, 0.f, 0.01f, 0.02f, 0.03f), 0.15f));
    Warpfield wf;
    wf.init(0.05f, nodes);
    pcl::PointCloud<pcl::PointXYZ> v;
    pcl::PointCloud<pcl::Normal> n;
    v.push_back(pcl::PointXYZ(1.05f, 1.05f, 1.0f));    // inside the patch: supported
    v.push_back(pcl::PointXYZ(1.62f, 1.01f, 1.0f));    // 0.42 m from the nearest node: unsupported
    v.push_back(pcl::PointXYZ(1.63f, 1.02f, 1.0f));    // same 5 cm cell as the previous one -> one centroid
    v.push_back(pcl::PointXYZ(1.01f, 1.01f, 1.9f));    // unsupported, far away in z
    for (int i = 0; i < 4; ++i) n.push_back(pcl::Normal(0, 0, 1));
    auto frame = std::make_shared<dynfu::Frame>(1, v, n);
    auto uns = wf.getUnsupportedVertices(frame);
    ASSERT_NEAR((double) uns->size(), 3.0, 0.0);
    ASSERT_NEAR((*uns)[0].x, 1.62f, 0.0);
    wf.update(frame);
    auto all = wf.getNodes();
    ASSERT_NEAR((double) all.size(), 11.0, 0.0);
    // cells in ascending linear index: the z = 1.0 cell comes before the z = 1.9 cell
    ASSERT_NEAR(all[9]->getPosition().x, 1.625f, 1e-6);
    ASSERT_NEAR(all[9]->getPosition().y, 1.015f, 1e-6);
    ASSERT_NEAR(all[10]->getPosition().z, 1.9f, 1e-6);
    ASSERT_NEAR(all[9]->getRadialBasisWeight(), 0.1f, 0.0);  // 2 * epsilon
    // all old nodes carry the same translation -> calcDQB at the new node returns it (weights are ~0 far away, so
    // the blend degenerates exactly like the reference's; just check the node is usable)
    auto q = wf.findNeighborsIndex(KNN, pcl::PointXYZ(1.62f, 1.01f, 1.0f));
    ASSERT_NEAR((double) q[0], 9.0, 0.0);  // the new node is now the nearest neighbour of the vertex it came from
}

// DynFusion::findCorrespondingFrame (src/dynfu/dyn_fusion.cpp:212-242)
TEST_F(UpdateTest, FindCorrespondingFrame) {
    pcl::PointCloud<pcl::PointXYZ> canon, live;
    pcl::PointCloud<pcl::Normal> cn;
    for (int i = 0; i < 100; ++i) {
        canon.push_back(pcl::PointXYZ(0.01f * i, 0.5f, 2.f));
        cn.push_back(pcl::Normal((float) i, 0, 0));
    }
    live.push_back(pcl::PointXYZ(0.302f, 0.6f, 2.1f));   // nearest: i = 30
    live.push_back(pcl::PointXYZ(-5.f, 0.5f, 2.f));      // nearest: i = 0
    live.push_back(pcl::PointXYZ(0.996f, 0.5f, 2.f));    // nearest: i = 99 (clamped at the end of the row)
    auto f = findCorrespondingFrame(canon, cn, live);
    ASSERT_NEAR((double) f->getVertices().size(), 3.0, 0.0);
    ASSERT_NEAR(f->getNormals()[0].normal_x, 30.f, 0.0);
    ASSERT_NEAR(f->getNormals()[1].normal_x, 0.f, 0.0);
    ASSERT_NEAR(f->getNormals()[2].normal_x, 99.f, 0.0);
    ASSERT_NEAR(f->getVertices()[0].x, 0.3f, 1e-7);
}

int main() { return RUN_ALL_TESTS(); }
