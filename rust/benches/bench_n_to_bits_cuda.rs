// The two lines a maintainer adds to the reference's criterion groups
// (benches/bench_n_to_bits.rs:15-19 and :44-47), as a standalone bench so it can live outside the crate.
use criterion::*;
use cute_nucleotides_cuda::*;

fn bench_n_to_bits(c: &mut Criterion) {
    let n = black_box(b"ATCG".repeat(10000));
    let mut group = c.benchmark_group("n_to_bits");
    group.throughput(Throughput::Bytes(40000));
    group.bench_function("n_to_bits_cuda", |b| b.iter(|| n_to_bits_cuda(&n)));
    group.finish();
}

fn bench_bits_to_n(c: &mut Criterion) {
    let bits = black_box(n_to_bits_cuda(&(b"ATCG".repeat(10000))));
    let len = black_box(4 * 10000);
    let mut group = c.benchmark_group("bits_to_n");
    group.throughput(Throughput::Bytes(40000));
    group.bench_function("bits_to_n_cuda", |b| b.iter(|| bits_to_n_cuda(&bits, len)));
    group.finish();
}

criterion_group!(benches, bench_n_to_bits, bench_bits_to_n);
criterion_main!(benches);
