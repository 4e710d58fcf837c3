// C++ caller of include/mc_simd.hpp: the reference's module surface, name for name.  Without a GPU every function
// must return NaN with a reason; with one (argv[1] == "gpu") the reference's own test valid_price1 must pass.
#include <cmath>
#include <cstdio>
#include <cstring>
#include "mc_simd.hpp"

int main(int argc, char** argv) {
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    const float p = mc_simd::call_price(100.f, 110.f, 0.25f, 0.05f, 0.5f, 0.02f, 100.f, 1000.f);       // mc_simd.rs:781-787
    const float d = mc_simd::call_delta(100.f, 0.001f, 110.f, 0.25f, 0.05f, 0.5f, 0.02f, 100.f, 1000.f); // :837-843
    const float t = mc_simd::put_theta(110.f, 120.f, 0.1f, 0.1f, 0.5f, 0.001f, 0.02f, 100.f, 10000.f);   // :893-899
    if (!gpu) {
        if (!std::isnan(p) || !std::isnan(d) || !std::isnan(t)) return 1;
        if (std::strstr(mc_simd::last_error(), "no CPU path") == nullptr) return 2;
        std::printf("cpp client ok (no device)\n");
        return 0;
    }
    if (!(std::fabs(p - 3.859768f) <= 1.25f)) { std::printf("price %f\n", p); return 3; }
    if (!(std::fabs(d - 0.353660f) <= 0.05f)) { std::printf("delta %f\n", d); return 4; }
    if (!(std::fabs(t - 4.531535f) <= 0.1f)) { std::printf("theta %f\n", t); return 5; }
    std::printf("cpp client ok (gpu): %f %f %f\n", p, d, t);
    return 0;
}
