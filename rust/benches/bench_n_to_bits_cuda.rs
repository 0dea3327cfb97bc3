// Criterion harness for the `_cuda` variants (source only here: no rustc/cargo in this image).
//
// Unlike the reference's fixed 40 000-nucleotide bench, this sweeps input sizes, because a GPU call is
// launch/PCIe-latency bound at the small end and link-bound at the large end.  The two group names match the
// reference's (`n_to_bits`, `bits_to_n`) so the results land next to its variants in criterion's reports; inside the
// reference crate itself the drop-in is the one-liner per group shown in INTEGRATION.md section 4.
use criterion::{black_box, criterion_group, criterion_main, BenchmarkId, Criterion, Throughput};
use cute_nucleotides_cuda::{bits_to_n_cuda, n_to_bits_cuda, n_to_bits_cuda_batch};

const SIZES: [usize; 4] = [40_000, 1 << 20, 1 << 26, 1 << 30];

fn sequence(len: usize) -> Vec<u8> {
    b"ATCG".iter().cycle().take(len).copied().collect()
}

fn codec_benches(c: &mut Criterion) {
    for &len in SIZES.iter() {
        let nucleotides = black_box(sequence(len));
        let packed = black_box(n_to_bits_cuda(&nucleotides));

        let mut encode = c.benchmark_group("n_to_bits");
        encode.throughput(Throughput::Bytes(len as u64));
        encode.bench_with_input(BenchmarkId::new("n_to_bits_cuda", len), &nucleotides, |b, n| {
            b.iter(|| n_to_bits_cuda(n))
        });
        encode.finish();

        let mut decode = c.benchmark_group("bits_to_n");
        decode.throughput(Throughput::Bytes(len as u64));
        decode.bench_with_input(BenchmarkId::new("bits_to_n_cuda", len), &packed, |b, w| {
            b.iter(|| bits_to_n_cuda(w, len))
        });
        decode.finish();
    }
}

// The reference's own bench shape (one 40 000-nucleotide string per call, benches/bench_n_to_bits.rs:10) many at a time:
// one kernel launch for the whole batch instead of one per string.
fn batch_bench(c: &mut Criterion) {
    let one = sequence(40_000);
    let many: Vec<&[u8]> = (0..4096).map(|_| one.as_slice()).collect();
    let mut group = c.benchmark_group("n_to_bits");
    group.throughput(Throughput::Bytes((40_000 * many.len()) as u64));
    group.bench_function("n_to_bits_cuda_batch", |b| b.iter(|| n_to_bits_cuda_batch(black_box(&many))));
    group.finish();
}

criterion_group!(benches, codec_benches, batch_bench);
criterion_main!(benches);
