// tools/ref_vectors.rs -- UNTESTED TEXT (this image has no Rust toolchain).  For a maintainer of the reference crate.
//
// The B200 library's parity claims are pinned to a CPU restatement of src/mc_simd.rs (oracle/ref_mc.c).  Two facts about
// the third-party crate simd_rand (Cargo.lock:716-718) could not be read offline and are ASSUMED there:
//   (1) how the 256 seed bytes of Xoshiro256PlusPlusX8Seed map onto the 8 lanes x 4 state words,
//   (2) how next_f64x8() turns a 64-bit output into an f64 in [0,1).
// wide 0.7.13's f32x8 ln / sin_cos / exp polynomials are the third unknown (tolerance, not convention).
// This test dumps exact vectors from the REAL crate so those assumptions can be falsified in one cargo run.
//
// How to use (in a checkout of ashayp22/monte-carlo-options-simd):
//   1. paste the function below into the `#[cfg(test)]`-style test section at the end of src/mc_simd.rs
//      (it needs the private items speed_update and get_rand_uniform_f32x8, hence in-crate);
//   2. cargo +nightly test dump_ref_vectors -- --nocapture | sed -n '/^REFVEC_BEGIN$/,/^REFVEC_END$/p' \
//          | sed '1d;$d' > ref_vectors.json
//   3. copy ref_vectors.json to tests/golden/ref_vectors.json of the B200 repository and run
//          python -m pytest tests/test_oracle.py -k reference_vectors
//      The test tries both seed layouts and both uniform conventions, reports which combination reproduces the
//      vectors bit for bit (uniforms) / to 1e-6 relative (chunk sums, where wide's polynomials differ from libm), and
//      fails if it is not the assumed one -- in which case flip g_seed_layout / g_uniform_mode in oracle/ref_mc.c and
//      the matching lines of the STREAM_REF kernel (csrc/mcb200_kernels.cuh, "simd_rand seed layout", and
//      u53_to_f32 in csrc/mcb200_pair.cuh).
//
// #[test]
// fn dump_ref_vectors() {
//     use std::fmt::Write;
//     // a fixed, non-trivial seed: byte i = (37 * i + 11) mod 256
//     let mut seed: Xoshiro256PlusPlusX8Seed = Default::default();
//     for (i, b) in (&mut *seed).iter_mut().enumerate() { *b = ((37 * i + 11) % 256) as u8; }
//     let mut s = String::new();
//     write!(s, "{{\n \"seed_rule\": \"byte i = (37*i + 11) mod 256\",\n").unwrap();
//
//     // (a) raw f64 outputs of next_f64x8: 4 calls x 8 lanes, as IEEE-754 bit patterns
//     let mut rng = Xoshiro256PlusPlusX8::from_seed(seed.clone());
//     write!(s, " \"next_f64x8_bits\": [").unwrap();
//     for call in 0..4 {
//         let v: [f64; 8] = rng.next_f64x8().to_array();
//         write!(s, "{}[{}]", if call > 0 { ", " } else { "" },
//                v.iter().map(|x| x.to_bits().to_string()).collect::<Vec<_>>().join(", ")).unwrap();
//     }
//     write!(s, "],\n").unwrap();
//
//     // (b) the f32 uniforms the hot loop sees (rand32x8.rs:5-21): 4 calls x 8 lanes, bit patterns
//     let mut rng = Xoshiro256PlusPlusX8::from_seed(seed.clone());
//     write!(s, " \"uniform_f32_bits\": [").unwrap();
//     for call in 0..4 {
//         let v: [f32; 8] = get_rand_uniform_f32x8(&mut rng).to_array();
//         write!(s, "{}[{}]", if call > 0 { ", " } else { "" },
//                v.iter().map(|x| x.to_bits().to_string()).collect::<Vec<_>>().join(", ")).unwrap();
//     }
//     write!(s, "],\n").unwrap();
//
//     // (c) one chunk's stock_price_mult after half_steps = 50 calls of speed_update (mc_simd.rs:9-21,56-60)
//     let mut rng = Xoshiro256PlusPlusX8::from_seed(seed.clone());
//     let two_pi: f32x8 = f32x8::splat(2.0 * std::f32::consts::PI);
//     let mut m: f32x8 = f32x8::splat(0.0);
//     for _ in 0..50 { m = speed_update(two_pi, m, &mut rng); }
//     let mv: [f32; 8] = m.to_array();
//     write!(s, " \"half_steps\": 50,\n \"stock_price_mult_f32_bits\": [{}],\n",
//            mv.iter().map(|x| x.to_bits().to_string()).collect::<Vec<_>>().join(", ")).unwrap();
//
//     // (d) the payoff vector of that chunk for the README bench option (mc_simd.rs:36-45,62-69):
//     //     spot 100, strike 110, vol 0.25, r 0.05, T 0.5, q 0.02, steps 100
//     let (spot, strike, vol, r, t, q, steps) = (100.0f32, 110.0f32, 0.25f32, 0.05f32, 0.5f32, 0.02f32, 100.0f32);
//     let dt = t / steps;
//     let nudt = (r - q - 0.5 * (vol * vol)) * dt;
//     let sidt = vol * dt.sqrt();
//     let pay = f32x8::fast_max(
//         f32x8::mul_sub(f32x8::splat(spot),
//                        f32x8::mul_add(m, f32x8::splat(std::f32::consts::SQRT_2 * sidt), f32x8::splat(steps * nudt)).exp(),
//                        f32x8::splat(strike)),
//         f32x8::splat(0.0));
//     let pv: [f32; 8] = pay.to_array();
//     write!(s, " \"call_payoff_f32_bits\": [{}]\n}}\n",
//            pv.iter().map(|x| x.to_bits().to_string()).collect::<Vec<_>>().join(", ")).unwrap();
//     println!("REFVEC_BEGIN\n{}REFVEC_END", s);
// }
