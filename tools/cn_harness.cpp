// cn_harness.cpp -- C++ stand-in for the reference's `cargo test` / `cargo bench` on the `_cuda` variant
// (Rust cannot be compiled in this image).  It only uses the C++ host mirror and, through it, the C ABI.
//
//   cn_harness test    known-answer tests written like the reference's (src/n_to_bits.rs:408-470) plus
//                      the cases the reference leaves untested; exit code 0 = all pass
//   cn_harness bench   the reference's criterion groups `n_to_bits` and `bits_to_n`
//                      (benches/bench_n_to_bits.rs:9-23, 37-50): 40 000 nt of "ATCG", Throughput::Bytes(40000),
//                      output allocation inside the timed region; also 1 MiB, 64 MiB and 1 GiB inputs
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../cute_nucleotides_b200/cpp/cute_nucleotides.hpp"

using namespace cute_nucleotides::n_to_bits;
using namespace cute_nucleotides::n_to_bits2;

static int g_failed = 0;
template <class A, class B> static bool same(const A &a, const B &b)
{
    if constexpr (std::is_arithmetic<A>::value) return a == b;
    else return a.size() == b.size() && std::equal(a.begin(), a.end(), b.begin());
}
#define ASSERT_EQ(a, b)                                                                        \
    do {                                                                                       \
        if (!same((a), (b))) { std::printf("FAILED %s:%d  %s == %s\n", __FILE__, __LINE__, #a, #b); g_failed++; } \
    } while (0)

static std::vector<uint8_t> bytes(const std::string &s) { return std::vector<uint8_t>(s.begin(), s.end()); }
static std::string repeat(const std::string &unit, size_t times)
{
    std::string s;
    s.reserve(unit.size() * times);
    for (size_t i = 0; i < times; i++) s += unit;
    return s;
}

static void test_n_to_bits_cuda()
{   // src/n_to_bits.rs:412-417 (and every other encoder test)
    ASSERT_EQ(n_to_bits_cuda("ATCGATCGATCGATCGATCGATCGATCGATCG"),
              (std::vector<uint64_t>{0b1101100011011000110110001101100011011000110110001101100011011000ull}));
    ASSERT_EQ(n_to_bits_cuda("ATCG"), (std::vector<uint64_t>{0b11011000ull}));
}

static void test_bits_to_n_cuda()
{   // src/n_to_bits.rs:419-423
    ASSERT_EQ(bits_to_n_cuda(std::vector<uint64_t>{0b1101100011011000110110001101100011011000110110001101100011011000ull}, 32),
              bytes("ATCGATCGATCGATCGATCGATCGATCGATCG"));
}

static void test_case_insensitive_and_u()
{   // BYTE_LUT maps a/A, c/C, g/G, t/T, u/U (src/n_to_bits.rs:8-21); decode is always upper case, U -> T
    ASSERT_EQ(n_to_bits_cuda("atcguAUCG"), n_to_bits_cuda("ATCGTATCG"));
    auto w = n_to_bits_cuda("acgtuACGTU");
    ASSERT_EQ(bits_to_n_cuda(w, 10), bytes("ACGTTACGTT"));
}

static void test_body_and_tail_lengths()
{
    const std::string unit = "GATTACAgattacaUuCcGgAaTt";
    for (size_t len : {0, 1, 31, 32, 33, 63, 64, 65, 127, 129, 1000, 40000, 100003}) {
        std::string s = repeat(unit, len / unit.size() + 1).substr(0, len);
        auto w = n_to_bits_cuda(s);
        ASSERT_EQ(w.size(), (len + 31) / 32);
        if (len % 32) ASSERT_EQ(w.back() >> (2 * (len % 32)), 0ull);          // unused high bits are zero
        std::string canon = s;
        for (auto &c : canon) { c = (char)std::toupper((unsigned char)c); if (c == 'U') c = 'T'; }
        ASSERT_EQ(bits_to_n_cuda(w, len), bytes(canon));
        if (len > 40) ASSERT_EQ(bits_to_n_cuda(w, len - 37), bytes(canon.substr(0, len - 37)));   // truncating len
    }
}

static void test_length_panic()
{   // src/n_to_bits.rs:52-54: panic!("The length is greater than the number of nucleotides!")
    bool thrown = false;
    try { bits_to_n_cuda(std::vector<uint64_t>{0, 0}, 65); }
    catch (const std::length_error &e) { thrown = std::string(e.what()) == "The length is greater than the number of nucleotides!"; }
    ASSERT_EQ(thrown, true);
    ASSERT_EQ(bits_to_n_cuda(std::vector<uint64_t>{0, 0}, 64), bytes(std::string(64, 'A')));
}

static void test_unaligned_slice()
{
    std::string s = repeat("ATCGGCTA", 5000);
    for (size_t off : {1, 3, 7, 13}) {
        std::string_view v(s.data() + off, s.size() - off);
        auto w = n_to_bits_cuda(v);
        ASSERT_EQ(bits_to_n_cuda(w, v.size()), bytes(std::string(v)));
    }
}

static void test_n_to_bits2_cuda()
{   // src/n_to_bits2.rs:275-279 (and :289-293)
    ASSERT_EQ(n_to_bits2_cuda("ATCGNATCGNATCGNATCGNATCGNATCGNATCGN"),
              (std::vector<uint64_t>{0b11011010100100010111010001111101000110110101001000101110100011ull, 0b1011101000111110100ull}));
    ASSERT_EQ(n_to_bits2_cuda("ATCGN"), (std::vector<uint64_t>{0b101110100011ull}));
}

static void test_bits_to_n2_cuda()
{   // src/n_to_bits2.rs:282-286 (and :296-298)
    ASSERT_EQ(bits_to_n2_cuda(std::vector<uint64_t>{0b11011010100100010111010001111101000110110101001000101110100011ull, 0b1011101000111110100ull}, 35),
              bytes("ATCGNATCGNATCGNATCGNATCGNATCGNATCGN"));
    bool thrown = false;
    try { bits_to_n2_cuda(std::vector<uint64_t>{0, 0}, 55); }
    catch (const std::length_error &e) { thrown = std::string(e.what()) == "The length is greater than the number of nucleotides!"; }
    ASSERT_EQ(thrown, true);
}

static void test_base5_round_trip_lengths()
{
    const std::string unit = "GATTACANgattacanUuCcGgAaTtNn";
    for (size_t len : {0, 1, 2, 26, 27, 28, 55, 3456, 3457, 40000, 100003}) {
        std::string s = repeat(unit, len / unit.size() + 1).substr(0, len);
        auto w = n_to_bits2_cuda(s);
        ASSERT_EQ(w.size(), (len + 26) / 27);
        std::string canon = s;
        for (auto &c : canon) { c = (char)std::toupper((unsigned char)c); if (c == 'U') c = 'T'; }
        ASSERT_EQ(bits_to_n2_cuda(w, len), bytes(canon));
    }
}

static void test_lut_exact_mode()
{   // n_to_bits_lut maps every byte outside {ACGTUacgtu} to 0 (BYTE_LUT, src/n_to_bits.rs:8-21); the plain variant follows
    // the SIMD bodies' (b >> 1) & 3.  'N' = 0x4E -> 3 there, 0 here.
    const std::string s = "ACGTNNNNacgtnnnn-*.ACGT";
    uint64_t invalid = 0;
    auto w = n_to_bits_lut_cuda(reinterpret_cast<const uint8_t *>(s.data()), s.size(), &invalid);
    ASSERT_EQ(invalid, (uint64_t)11);
    std::string clean = s;
    for (auto &c : clean) if (std::string("ACGTUacgtu").find(c) == std::string::npos) c = 'A';
    ASSERT_EQ(w, n_to_bits_cuda(clean));
    ASSERT_EQ(bits_to_n_cuda(w, s.size()), bytes("ACGTAAAAACGTAAAAAAAACGT"));
}

static void test_batch()
{   // element i of the batch == the single call on sequence i (the reference's bench loop, benches/bench_n_to_bits.rs:15-19)
    std::vector<std::string> store;
    const std::string unit = "GATTACAgattacaUuCcGgAaTt";
    for (size_t len : {4, 32, 0, 33, 150, 40000, 31, 100003, 1}) store.push_back(repeat(unit, len / unit.size() + 1).substr(0, len));
    std::vector<std::string_view> views(store.begin(), store.end());
    auto res = n_to_bits_cuda_batch(views);
    ASSERT_EQ(res.size(), store.size());
    for (size_t i = 0; i < store.size(); i++) ASSERT_EQ(res[i], n_to_bits_cuda(store[i]));
}

static void test_fan_out_is_transparent()
{   // cn_set_devices: the same call, the same words, however many GPUs carry it (device 0 twice when only one is visible)
    int count = 0;
    cn_device_count(&count);
    std::vector<int> devs;
    for (int d = 0; d < std::max(count, 2); d++) devs.push_back(d % std::max(count, 1));
    const std::string s = repeat("ATCGGCTAacgu", (40u << 20) / 12);
    const auto one = n_to_bits_cuda(s);
    cute_nucleotides::n_to_bits::set_devices(devs);
    const auto many = n_to_bits_cuda(s);
    ASSERT_EQ(bits_to_n_cuda(many, s.size()).size(), s.size());
    cute_nucleotides::n_to_bits::set_devices({});
    ASSERT_EQ(one, many);
}

template <typename F> static double median_seconds(F f, int min_iters, double min_total)
{
    std::vector<double> t;
    double total = 0;
    while ((int)t.size() < min_iters || total < min_total) {
        auto a = std::chrono::steady_clock::now();
        f();
        auto b = std::chrono::steady_clock::now();
        t.push_back(std::chrono::duration<double>(b - a).count());
        total += t.back();
        if (t.size() > 200000) break;
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

static void bench_group(size_t repeat_count)
{
    const std::string n = repeat("ATCG", repeat_count);          // get_nucleotides(), benches/bench_n_to_bits.rs:68-70
    const size_t len = n.size();
    const auto bits = n_to_bits_cuda(n);                         // get_bits(), :76-78
    n_to_bits_cuda(n);
    double te = median_seconds([&] { auto r = n_to_bits_cuda(n); asm volatile("" ::"r"(r.data()) : "memory"); }, 20, 0.5);
    double td = median_seconds([&] { auto r = bits_to_n_cuda(bits, len); asm volatile("" ::"r"(r.data()) : "memory"); }, 20, 0.5);
    const double gib = 1024.0 * 1024.0 * 1024.0;
    std::printf("{\"group\": \"n_to_bits\", \"function\": \"n_to_bits_cuda\", \"nucleotides\": %zu, \"time_us\": %.2f, \"thrpt_gib_s\": %.4f}\n",
                len, te * 1e6, len / te / gib);
    std::printf("{\"group\": \"bits_to_n\", \"function\": \"bits_to_n_cuda\", \"nucleotides\": %zu, \"time_us\": %.2f, \"thrpt_gib_s\": %.4f}\n",
                len, td * 1e6, len / td / gib);
    std::fflush(stdout);
}

static void bench_batch(size_t repeat_count, size_t sequences)
{
    const std::string n = repeat("ATCG", repeat_count);
    std::vector<std::string_view> many(sequences, std::string_view(n));
    n_to_bits_cuda_batch(many);
    double t = median_seconds([&] { auto r = n_to_bits_cuda_batch(many); asm volatile("" ::"r"(r.data()) : "memory"); }, 5, 0.5);
    const double gib = 1024.0 * 1024.0 * 1024.0;
    std::printf("{\"group\": \"n_to_bits\", \"function\": \"n_to_bits_cuda_batch\", \"sequences\": %zu, \"nucleotides_each\": %zu, "
                "\"time_us_per_sequence\": %.3f, \"thrpt_gib_s\": %.4f}\n", sequences, n.size(), t * 1e6 / sequences, n.size() * sequences / t / gib);
    std::fflush(stdout);
}

int main(int argc, char **argv)
{
    const std::string mode = argc > 1 ? argv[1] : "test";
    try {
        if (mode == "test") {
            test_n_to_bits_cuda();
            test_bits_to_n_cuda();
            test_case_insensitive_and_u();
            test_body_and_tail_lengths();
            test_length_panic();
            test_unaligned_slice();
            test_n_to_bits2_cuda();
            test_bits_to_n2_cuda();
            test_base5_round_trip_lengths();
            test_lut_exact_mode();
            test_batch();
            test_fan_out_is_transparent();
            std::printf(g_failed ? "test result: FAILED. %d failed\n" : "test result: ok. 12 passed; %d failed\n", g_failed);
            return g_failed ? 1 : 0;
        }
        if (mode == "bench") {
            bench_batch(10000, 4096);           // 4096 strings of the reference's bench size in ONE call
            bench_group(10000);                 // the reference's bench size: 40 000 nt
            bench_group((1u << 20) / 4);        // 1 MiB
            bench_group((64u << 20) / 4);       // 64 MiB
            bench_group((1u << 30) / 4);        // 1 GiB
            return 0;
        }
    } catch (const std::exception &e) {
        std::printf("error: %s\n", e.what());
        return 2;
    }
    std::printf("usage: cn_harness test|bench\n");
    return 64;
}
