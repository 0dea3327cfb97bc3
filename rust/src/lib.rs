//! `_cuda` variants of cute-nucleotides' 2-bit codec, backed by hand-written sm_100a kernels.
//!
//! Same signatures as the reference family (`src/n_to_bits.rs:34,51` in cute-nucleotides):
//! `n_to_bits_cuda(&[u8]) -> Vec<u64>` and `bits_to_n_cuda(&[u64], usize) -> Vec<u8>`, so the
//! reference's bench groups and tests can call them as one more variant.
//!
//! Ownership: the `Vec` is allocated here, by Rust's allocator, and filled in place by the C side;
//! no C-allocated memory ever becomes a `Vec`.  Errors: the length check panics with the reference's
//! message; a CUDA failure panics with `cn_last_error()`.  There is no CPU fallback.
//!
//! This file is SOURCE ONLY in this repository: the build image has no rustc/cargo.  The identical
//! logic is exercised through the C++ mirror (`cute_nucleotides_b200/cpp/cute_nucleotides.hpp`,
//! `tools/cn_harness.cpp`) and the Python mirror (`cute_nucleotides_b200/n_to_bits.py`).

use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

pub mod ffi {
    use super::*;
    pub const CN_OK: c_int = 0;
    pub const CN_ERR_LENGTH: c_int = 1;

    extern "C" {
        pub fn cn_abi_version() -> c_int;
        pub fn cn_init(device: c_int) -> c_int;
        pub fn cn_last_error() -> *const c_char;
        pub fn cn_length_panic_message() -> *const c_char;
        pub fn cn_words_for_len(len: usize) -> usize;
        pub fn cn_n_to_bits_host(n: *const u8, len: usize, out: *mut u64) -> c_int;
        pub fn cn_bits_to_n_host(bits: *const u64, nwords: usize, len: usize, out: *mut u8) -> c_int;
        pub fn cn_n_to_bits2_host(n: *const u8, len: usize, out: *mut u64) -> c_int;
        pub fn cn_bits_to_n2_host(bits: *const u64, nwords: usize, len: usize, out: *mut u8) -> c_int;
        pub fn cn_words2_for_len(len: usize) -> usize;
        // ABI v2
        pub fn cn_n_to_bits_ex_host(n: *const u8, len: usize, out: *mut u64, mode: c_int, invalid_count: *mut u64) -> c_int;
        pub fn cn_n_to_bits_host_batch(seqs: *const *const u8, lens: *const usize, count: usize, outs: *const *mut u64) -> c_int;
        pub fn cn_bits_to_n_host_batch(bits: *const *const u64, lens: *const usize, count: usize, outs: *const *mut u8) -> c_int;
        pub fn cn_set_devices(devices: *const c_int, count: c_int) -> c_int;
        pub fn cn_n_to_bits_host_async(n: *const u8, len: usize, out: *mut u64, req: *mut *mut CnRequest) -> c_int;
        pub fn cn_bits_to_n_host_async(bits: *const u64, nwords: usize, len: usize, out: *mut u8, req: *mut *mut CnRequest) -> c_int;
        pub fn cn_wait(req: *mut CnRequest) -> c_int;
        pub fn cn_hamming_host(a: *const u64, b: *const u64, nwords: usize, len: usize, result: *mut u64) -> c_int;
        pub fn cn_reverse_complement_host(bits: *const u64, nwords: usize, len: usize, out: *mut u64) -> c_int;
    }
    pub const CN_ENC_LUT_EXACT: c_int = 2;
    #[repr(C)]
    pub struct CnRequest { _private: [u8; 0] }
}

fn fail(status: c_int) -> ! {
    let msg = unsafe {
        let p = if status == ffi::CN_ERR_LENGTH { ffi::cn_length_panic_message() } else { ffi::cn_last_error() };
        CStr::from_ptr(p).to_string_lossy().into_owned()
    };
    panic!("{}", msg);
}

/// Encode `{A, T/U, C, G}` from the byte string into pairs of bits (`{00, 10, 01, 11}`) packed into
/// 64-bit integers, on a B200.
pub fn n_to_bits_cuda(n: &[u8]) -> Vec<u64> {
    let words = (n.len() >> 5) + if n.len() & 31 == 0 { 0 } else { 1 };
    let mut res: Vec<u64> = Vec::with_capacity(words);
    unsafe {
        let status = ffi::cn_n_to_bits_host(n.as_ptr(), n.len(), res.as_mut_ptr());
        if status != ffi::CN_OK {
            fail(status);
        }
        res.set_len(words);
    }
    res
}

/// Like `n_to_bits_cuda`, but bytes outside `{A,C,G,T,U,a,c,g,t,u}` encode as 0 exactly as `n_to_bits_lut`'s `BYTE_LUT` does
/// (`src/n_to_bits.rs:8-21`), so the result equals `n_to_bits_lut(n)` on EVERY input.  Also returns how many such bytes there were.
pub fn n_to_bits_lut_cuda(n: &[u8]) -> (Vec<u64>, u64) {
    let words = (n.len() >> 5) + if n.len() & 31 == 0 { 0 } else { 1 };
    let mut res: Vec<u64> = Vec::with_capacity(words);
    let mut invalid = 0u64;
    unsafe {
        let status = ffi::cn_n_to_bits_ex_host(n.as_ptr(), n.len(), res.as_mut_ptr(), ffi::CN_ENC_LUT_EXACT, &mut invalid);
        if status != ffi::CN_OK {
            fail(status);
        }
        res.set_len(words);
    }
    (res, invalid)
}

/// Many independent sequences in ONE call (one kernel launch for thousands of reads): element `i` of the result is
/// exactly `n_to_bits_cuda(seqs[i])`.
pub fn n_to_bits_cuda_batch(seqs: &[&[u8]]) -> Vec<Vec<u64>> {
    let lens: Vec<usize> = seqs.iter().map(|s| s.len()).collect();
    let ins: Vec<*const u8> = seqs.iter().map(|s| s.as_ptr()).collect();
    let mut res: Vec<Vec<u64>> = lens.iter().map(|&l| Vec::with_capacity((l >> 5) + if l & 31 == 0 { 0 } else { 1 })).collect();
    let outs: Vec<*mut u64> = res.iter_mut().map(|v| v.as_mut_ptr()).collect();
    unsafe {
        let status = ffi::cn_n_to_bits_host_batch(ins.as_ptr(), lens.as_ptr(), seqs.len(), outs.as_ptr());
        if status != ffi::CN_OK {
            fail(status);
        }
        for (v, &l) in res.iter_mut().zip(lens.iter()) {
            v.set_len((l >> 5) + if l & 31 == 0 { 0 } else { 1 });
        }
    }
    res
}

/// An encode running on a library thread; `wait()` returns the words.  Keeping one of these in flight while the previous
/// batch is decoded loads both PCIe directions at once.
pub struct PendingEncode<'a> {
    req: *mut ffi::CnRequest,
    res: Vec<u64>,
    words: usize,
    _input: std::marker::PhantomData<&'a [u8]>,
}

pub fn n_to_bits_cuda_async(n: &[u8]) -> PendingEncode<'_> {
    let words = (n.len() >> 5) + if n.len() & 31 == 0 { 0 } else { 1 };
    let mut res: Vec<u64> = Vec::with_capacity(words);
    let mut req: *mut ffi::CnRequest = std::ptr::null_mut();
    let status = unsafe { ffi::cn_n_to_bits_host_async(n.as_ptr(), n.len(), res.as_mut_ptr(), &mut req) };
    if status != ffi::CN_OK {
        fail(status);
    }
    PendingEncode { req, res, words, _input: std::marker::PhantomData }
}

impl<'a> PendingEncode<'a> {
    pub fn wait(mut self) -> Vec<u64> {
        let status = unsafe { ffi::cn_wait(self.req) };
        self.req = std::ptr::null_mut();
        if status != ffi::CN_OK {
            fail(status);
        }
        unsafe { self.res.set_len(self.words) };
        std::mem::take(&mut self.res)
    }
}

impl<'a> Drop for PendingEncode<'a> {
    fn drop(&mut self) {
        if !self.req.is_null() {
            unsafe { ffi::cn_wait(self.req) };      // the library still writes into `res`: never free it under a running call
        }
    }
}

/// Fan every later `*_cuda` call out over these GPUs (one PCIe link each); an empty slice restores single-GPU behaviour.
pub fn set_devices(devices: &[i32]) {
    let status = unsafe { ffi::cn_set_devices(devices.as_ptr(), devices.len() as c_int) };
    if status != ffi::CN_OK {
        fail(status);
    }
}

/// Decode pairs of bits from packed 64-bit integers to get a byte string of `{A, T, C, G}`, on a B200.
pub fn bits_to_n_cuda(bits: &[u64], len: usize) -> Vec<u8> {
    if len > (bits.len() << 5) {
        panic!("The length is greater than the number of nucleotides!");
    }
    let mut res: Vec<u8> = Vec::with_capacity(len);
    unsafe {
        let status = ffi::cn_bits_to_n_host(bits.as_ptr(), bits.len(), len, res.as_mut_ptr());
        if status != ffi::CN_OK {
            fail(status);
        }
        res.set_len(len);
    }
    res
}

/// Encode `{A, T/U, C, G, N}` triplets into 7 bits each, 9 triplets per 64-bit integer, on a B200
/// (mirrors `n_to_bits2_lut`, `src/n_to_bits2.rs:37`).
pub fn n_to_bits2_cuda(n: &[u8]) -> Vec<u64> {
    let words = (n.len() / 27) + if n.len() % 27 == 0 { 0 } else { 1 };
    let mut res: Vec<u64> = Vec::with_capacity(words);
    unsafe {
        let status = ffi::cn_n_to_bits2_host(n.as_ptr(), n.len(), res.as_mut_ptr());
        if status != ffi::CN_OK {
            fail(status);
        }
        res.set_len(words);
    }
    res
}

/// Decode 9 triplets per 64-bit integer back to `{A, T, C, G, N}` (mirrors `bits_to_n2_lut`,
/// `src/n_to_bits2.rs:78`).
pub fn bits_to_n2_cuda(bits: &[u64], len: usize) -> Vec<u8> {
    if len > bits.len() * 27 {
        panic!("The length is greater than the number of nucleotides!");
    }
    let mut res: Vec<u8> = Vec::with_capacity(len);
    unsafe {
        let status = ffi::cn_bits_to_n2_host(bits.as_ptr(), bits.len(), len, res.as_mut_ptr());
        if status != ffi::CN_OK {
            fail(status);
        }
        res.set_len(len);
    }
    res
}

#[cfg(test)]
mod tests {
    use super::*;

    // same vectors as the reference's tests (src/n_to_bits.rs:412-423)
    #[test]
    fn test_n_to_bits_cuda() {
        assert_eq!(n_to_bits_cuda(b"ATCGATCGATCGATCGATCGATCGATCGATCG"),
                vec![0b1101100011011000110110001101100011011000110110001101100011011000]);
        assert_eq!(n_to_bits_cuda(b"ATCG"), vec![0b11011000]);
    }

    #[test]
    fn test_bits_to_n_cuda() {
        assert_eq!(bits_to_n_cuda(&vec![0b1101100011011000110110001101100011011000110110001101100011011000], 32),
                "ATCGATCGATCGATCGATCGATCGATCGATCG".as_bytes());
    }

    #[test]
    #[should_panic(expected = "The length is greater than the number of nucleotides!")]
    fn test_bits_to_n_cuda_length() {
        bits_to_n_cuda(&vec![0u64; 2], 65);
    }

    // src/n_to_bits2.rs:275-298
    #[test]
    fn test_n_to_bits2_cuda() {
        assert_eq!(n_to_bits2_cuda(b"ATCGNATCGNATCGNATCGNATCGNATCGNATCGN"),
                vec![0b11011010100100010111010001111101000110110101001000101110100011, 0b1011101000111110100]);
        assert_eq!(n_to_bits2_cuda(b"ATCGN"), vec![0b101110100011]);
    }

    #[test]
    fn test_bits_to_n2_cuda() {
        assert_eq!(bits_to_n2_cuda(&vec![0b11011010100100010111010001111101000110110101001000101110100011, 0b1011101000111110100], 35),
                "ATCGNATCGNATCGNATCGNATCGNATCGNATCGN".as_bytes());
    }
}
