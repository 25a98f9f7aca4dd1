// Minimal stand-in for the handful of gtest macros the reference's tests use (gtest is not in this image).
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

struct MiniTestRegistry {
    struct Case { std::string name; std::function<void(int&)> fn; };
    static std::vector<Case>& cases() { static std::vector<Case> c; return c; }
};
#define TEST_F(fixture, name)                                                                          \
    struct fixture##_##name : fixture { void TestBody(int& failures_); };                              \
    static int reg_##fixture##_##name = (MiniTestRegistry::cases().push_back(                          \
        {#fixture "." #name, [](int& f) { fixture##_##name t; t.SetUp(); t.TestBody(f); t.TearDown(); }}), 0); \
    void fixture##_##name::TestBody(int& failures_)
#define ASSERT_NEAR(a, b, tol)                                                                         \
    do {                                                                                               \
        const double a_ = (a), b_ = (b);                                                               \
        if (!(std::fabs(a_ - b_) <= (tol))) {                                                          \
            std::printf("  %s:%d: ASSERT_NEAR(%s, %s) failed: %.9g vs %.9g\n", __FILE__, __LINE__, #a, #b, a_, b_); \
            ++failures_;                                                                               \
            return;                                                                                    \
        }                                                                                              \
    } while (0)
inline int RUN_ALL_TESTS() {
    int failed = 0;
    for (auto& c : MiniTestRegistry::cases()) {
        int f = 0;
        try {
            c.fn(f);
        } catch (const std::exception& e) {
            std::printf("  exception: %s\n", e.what());
            ++f;
        }
        std::printf("[%s] %s\n", f ? " FAILED " : "   OK   ", c.name.c_str());
        failed += f ? 1 : 0;
    }
    std::printf("%zu tests, %d failed\n", MiniTestRegistry::cases().size(), failed);
    return failed ? 1 : 0;
}
