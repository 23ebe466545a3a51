//! Times faer's own CPU path (rayon, all cores) on the configurations of BASELINE.json and prints one JSON line per
//! configuration, with the flop counts of SURVEY.md section 8d. Only the high-level solver API is used (`Mat::llt`,
//! `partial_piv_lu`, `qr`, `svd`, `&A * &B`), which is stable across the 0.2x releases.
//! usage: faer_rayon_bench [gemm|llt|lu|qr|svd|all] [n]
use faer::prelude::*;
use faer::{Mat, Par, Side};
use rand::SeedableRng;
use rand_distr::{Distribution, StandardNormal};
use std::time::Instant;

fn gaussian(m: usize, n: usize, seed: u64) -> Mat<f64> {
    let mut rng = rand::rngs::StdRng::seed_from_u64(seed);
    Mat::from_fn(m, n, |_, _| StandardNormal.sample(&mut rng))
}

fn gaussian_f32(m: usize, n: usize, seed: u64) -> Mat<f32> {
    let mut rng = rand::rngs::StdRng::seed_from_u64(seed);
    Mat::from_fn(m, n, |_, _| { let x: f64 = StandardNormal.sample(&mut rng); x as f32 })
}

fn best_of<F: FnMut()>(reps: usize, mut f: F) -> f64 {
    let mut best = f64::INFINITY;
    for _ in 0..reps {
        let t = Instant::now();
        f();
        best = best.min(t.elapsed().as_secs_f64());
    }
    best
}

fn report(what: &str, n: usize, flops: f64, secs: f64, cores: usize) {
    println!(
        "{{\"impl\": \"faer-rayon\", \"what\": \"{}\", \"n\": {}, \"seconds\": {:.6}, \"tflops\": {:.4}, \"cores\": {}}}",
        what, n, secs, flops / secs / 1e12, cores
    );
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let what = args.get(1).map(|s| s.as_str()).unwrap_or("all").to_string();
    let n_arg: Option<usize> = args.get(2).and_then(|s| s.parse().ok());
    faer::set_global_parallelism(Par::rayon(0));
    let cores = std::thread::available_parallelism().map(|c| c.get()).unwrap_or(1);

    if what == "gemm" || what == "all" {
        let n = n_arg.unwrap_or(16384);
        let a = gaussian(n, n, 0);
        let b = gaussian(n, n, 1);
        let secs = best_of(2, || { let c = &a * &b; std::hint::black_box(&c); });
        report("f64 matmul", n, 2.0 * (n as f64).powi(3), secs, cores);
    }
    if what == "llt" || what == "all" {
        let n = n_arg.unwrap_or(16384);
        let g = gaussian(n, n, 2);
        let mut a = &g * g.transpose();
        for i in 0..n { a[(i, i)] += n as f64; }
        let secs = best_of(2, || { let f = a.llt(Side::Lower); std::hint::black_box(&f).is_ok(); });
        report("f64 LLT", n, (n as f64).powi(3) / 3.0, secs, cores);
    }
    if what == "lu" || what == "all" {
        let n = n_arg.unwrap_or(32768);
        let a = gaussian(n, n, 3);
        let secs = best_of(1, || { let f = a.partial_piv_lu(); std::hint::black_box(&f); });
        report("f64 partial-pivoting LU", n, 2.0 * (n as f64).powi(3) / 3.0, secs, cores);
    }
    if what == "qr" || what == "all" {
        let (m, n) = (65536usize, n_arg.unwrap_or(4096));
        let a = gaussian_f32(m, n, 4);
        let secs = best_of(2, || { let f = a.qr(); std::hint::black_box(&f); });
        report("f32 QR 65536 x n", n, 2.0 * (m as f64) * (n as f64).powi(2) - 2.0 * (n as f64).powi(3) / 3.0, secs, cores);
    }
    if what == "svd" || what == "all" {
        let n = n_arg.unwrap_or(8192);
        let a = gaussian(n, n, 5);
        let secs = best_of(1, || { let f = a.svd(); std::hint::black_box(&f).is_ok(); });
        report("f64 SVD (full U, V)", n, 8.0 * (n as f64).powi(3) / 3.0, secs, cores);
    }
}
