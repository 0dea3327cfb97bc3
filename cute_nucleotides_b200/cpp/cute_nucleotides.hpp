// cute_nucleotides.hpp -- C++ host-side mirror of the reference module `cute_nucleotides::n_to_bits`
// (src/n_to_bits.rs) on top of the C ABI (include/cute_nucleotides_cuda.h).
//
// The reference is compiled code (Rust) whose operator API is a naming family:
//     pub fn n_to_bits_<variant>(n: &[u8]) -> Vec<u64>                  src/n_to_bits.rs:34,80,121,172,213
//     pub fn bits_to_n_<variant>(bits: &[u64], len: usize) -> Vec<u8>   src/n_to_bits.rs:51,265,309,346
// This header adds the `_cuda` variant with the same argument meaning, the same ownership (the callee
// returns an owned vector that the caller's allocator frees) and the same error behaviour: where the
// reference panics with "The length is greater than the number of nucleotides!" (:52-54) this throws
// std::length_error carrying the same text.  Any CUDA failure throws std::runtime_error -- there is
// no CPU fallback.  Rust cannot be compiled in this image; rust/src/lib.rs is the same shim in Rust.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/cute_nucleotides_cuda.h"

namespace cute_nucleotides {
namespace n_to_bits {

// Rust's Vec::with_capacity + set_len does not zero the buffer the library is about to overwrite;
// std::vector(n) would.  This allocator default-initialises instead, so the mirrors cost the same.
template <class T> struct default_init_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
using Words = std::vector<uint64_t, default_init_allocator<uint64_t>>;   // the Vec<u64> of the reference
using Bytes = std::vector<uint8_t, default_init_allocator<uint8_t>>;     // the Vec<u8> of the reference

inline void check_status(int status)
{
    if (status == CN_OK) return;
    if (status == CN_ERR_LENGTH) throw std::length_error(cn_length_panic_message());
    throw std::runtime_error(std::string("cute_nucleotides_cuda: ") + cn_last_error());
}

/// Encode `{A, T/U, C, G}` (any case) into pairs of bits (`{00, 10, 01, 11}`) packed into 64-bit integers
/// on the GPU.  Mirrors n_to_bits_lut (src/n_to_bits.rs:34).
inline Words n_to_bits_cuda(const uint8_t *n, size_t len)
{
    Words res(cn_words_for_len(len));                        // caller-side allocation, filled by the library
    check_status(cn_n_to_bits_host(n, len, res.data()));
    return res;
}
inline Words n_to_bits_cuda(std::string_view n)
{
    return n_to_bits_cuda(reinterpret_cast<const uint8_t *>(n.data()), n.size());
}
inline Words n_to_bits_cuda(const std::vector<uint8_t> &n) { return n_to_bits_cuda(n.data(), n.size()); }

/// n_to_bits_lut-exact: bytes outside {A,C,G,T,U,a,c,g,t,u} encode as 0 like BYTE_LUT (src/n_to_bits.rs:8-21), so the result
/// equals n_to_bits_lut(n) on EVERY input; *invalid (optional) receives the number of such bytes.
inline Words n_to_bits_lut_cuda(const uint8_t *n, size_t len, uint64_t *invalid = nullptr)
{
    Words res(cn_words_for_len(len));
    uint64_t count = 0;
    check_status(cn_n_to_bits_ex_host(n, len, res.data(), CN_ENC_LUT_EXACT, &count));
    if (invalid) *invalid = count;
    return res;
}

/// Many independent sequences in ONE call (one kernel launch for thousands of reads): element i == n_to_bits_cuda(seqs[i]).
inline std::vector<Words> n_to_bits_cuda_batch(const std::vector<std::string_view> &seqs)
{
    std::vector<Words> res(seqs.size());
    std::vector<const uint8_t *> in(seqs.size());
    std::vector<size_t> lens(seqs.size());
    std::vector<uint64_t *> out(seqs.size());
    for (size_t i = 0; i < seqs.size(); i++) {
        in[i] = reinterpret_cast<const uint8_t *>(seqs[i].data());
        lens[i] = seqs[i].size();
        res[i] = Words(cn_words_for_len(lens[i]));
        out[i] = res[i].data();
    }
    check_status(cn_n_to_bits_host_batch(in.data(), lens.data(), seqs.size(), out.data()));
    return res;
}

/// Fan every later *_cuda call out over these GPUs (one PCIe link each); an empty list restores single-GPU behaviour.
inline void set_devices(const std::vector<int> &devices) { check_status(cn_set_devices(devices.data(), (int)devices.size())); }

/// Decode pairs of bits from packed 64-bit integers to a byte string of `{A, T, C, G}` on the GPU.
/// Mirrors bits_to_n_lut (src/n_to_bits.rs:51): throws std::length_error if len > 32 * bits.size().
inline Bytes bits_to_n_cuda(const uint64_t *bits, size_t nwords, size_t len)
{
    if (len > (nwords << 5)) throw std::length_error(cn_length_panic_message());
    Bytes res(len);
    check_status(cn_bits_to_n_host(bits, nwords, len, res.data()));
    return res;
}
template <class A> inline Bytes bits_to_n_cuda(const std::vector<uint64_t, A> &bits, size_t len)
{
    return bits_to_n_cuda(bits.data(), bits.size(), len);
}

}  // namespace n_to_bits

// Mirror of `cute_nucleotides::n_to_bits2` (src/n_to_bits2.rs): {A, C, T/U, G, N}, 27 nucleotides per u64.
namespace n_to_bits2 {
using n_to_bits::Bytes;
using n_to_bits::Words;
using n_to_bits::check_status;

/// Mirrors n_to_bits2_lut (src/n_to_bits2.rs:37).
inline Words n_to_bits2_cuda(const uint8_t *n, size_t len)
{
    Words res(cn_words2_for_len(len));
    check_status(cn_n_to_bits2_host(n, len, res.data()));
    return res;
}
inline Words n_to_bits2_cuda(std::string_view n) { return n_to_bits2_cuda(reinterpret_cast<const uint8_t *>(n.data()), n.size()); }

/// Mirrors bits_to_n2_lut (src/n_to_bits2.rs:78): throws std::length_error if len > 27 * nwords (:79-81).
inline Bytes bits_to_n2_cuda(const uint64_t *bits, size_t nwords, size_t len)
{
    if (len > nwords * 27) throw std::length_error(cn_length_panic_message());
    Bytes res(len);
    check_status(cn_bits_to_n2_host(bits, nwords, len, res.data()));
    return res;
}
template <class A> inline Bytes bits_to_n2_cuda(const std::vector<uint64_t, A> &bits, size_t len)
{
    return bits_to_n2_cuda(bits.data(), bits.size(), len);
}
}  // namespace n_to_bits2
}  // namespace cute_nucleotides
