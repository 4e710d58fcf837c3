// mc_simd.hpp -- C++ mirror of the reference's public module `mc_simd` (/root/reference/src/mc_simd.rs:477-779) on top of
// the C ABI of libmcb200.so (mcb200.h).  Same twelve names, same f32 parameters in the same order, same meaning; like
// the Rust functions they return a bare float (NaN on failure -- the reason is in mc_simd::last_error()).
//
//     #include "mc_simd.hpp"
//     float p = mc_simd::call_price(100.f, 110.f, 0.25f, 0.05f, 0.5f, 0.02f, 100.f, 1000.f);
//
// Everything runs on the B200; there is no CPU fallback.  For throughput use the batched entry points of mcb200.h
// (one launch for a whole grid of options, call + put + standard errors, all Greeks from one set of paths).
#pragma once
#include "mcb200.h"

namespace mc_simd {

inline float call_price(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                        float dividend_yield, float steps, float num_trials) {                         // mc_simd.rs:477
    return mc_simd_call_price(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float put_price(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                       float dividend_yield, float steps, float num_trials) {                          // :500
    return mc_simd_put_price(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float call_price_av(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                           float dividend_yield, float steps, float num_trials) {                      // :523
    return mc_simd_call_price_av(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float put_price_av(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                          float dividend_yield, float steps, float num_trials) {                       // :546
    return mc_simd_put_price_av(spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float call_delta(float spot, float delta_spot, float strike, float volatility, float risk_free_rate,
                        float years_to_expiry, float dividend_yield, float steps, float num_trials) {  // :569
    return mc_simd_call_delta(spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float put_delta(float spot, float delta_spot, float strike, float volatility, float risk_free_rate,
                       float years_to_expiry, float dividend_yield, float steps, float num_trials) {   // :595
    return mc_simd_put_delta(spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float gamma(float spot, float delta_spot, float strike, float volatility, float risk_free_rate,
                   float years_to_expiry, float dividend_yield, float steps, float num_trials) {       // :621
    return mc_simd_gamma(spot, delta_spot, strike, volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float vega(float spot, float strike, float volatility, float delta_volatility, float risk_free_rate,
                  float years_to_expiry, float dividend_yield, float steps, float num_trials) {        // :647
    return mc_simd_vega(spot, strike, volatility, delta_volatility, risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float call_rho(float spot, float strike, float volatility, float risk_free_rate, float delta_risk_free_rate,
                      float years_to_expiry, float dividend_yield, float steps, float num_trials) {    // :673
    return mc_simd_call_rho(spot, strike, volatility, risk_free_rate, delta_risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float put_rho(float spot, float strike, float volatility, float risk_free_rate, float delta_risk_free_rate,
                     float years_to_expiry, float dividend_yield, float steps, float num_trials) {     // :701
    return mc_simd_put_rho(spot, strike, volatility, risk_free_rate, delta_risk_free_rate, years_to_expiry, dividend_yield, steps, num_trials);
}
inline float call_theta(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                        float delta_years_to_expiry, float dividend_yield, float steps, float num_trials) {   // :728
    return mc_simd_call_theta(spot, strike, volatility, risk_free_rate, years_to_expiry, delta_years_to_expiry, dividend_yield, steps, num_trials);
}
inline float put_theta(float spot, float strike, float volatility, float risk_free_rate, float years_to_expiry,
                       float delta_years_to_expiry, float dividend_yield, float steps, float num_trials) {    // :755
    return mc_simd_put_theta(spot, strike, volatility, risk_free_rate, years_to_expiry, delta_years_to_expiry, dividend_yield, steps, num_trials);
}
inline const char* last_error() { return mcb200_last_error(); }

}  // namespace mc_simd
