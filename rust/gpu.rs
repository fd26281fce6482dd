//! Drop-in replacement for ark-blst's `src/gpu.rs` (reference src/gpu.rs:1-286).
//!
//! The reference file wraps ec-gpu-gen's generated kernel: per call it enumerates devices, builds a
//! `Program`, uploads bases and exponents, launches one kernel, downloads 18 944 partials and folds
//! them on the host (src/gpu.rs:126-241).  This file shrinks to an `extern "C"` block over
//! `libb200msm.so` (include/b200msm.h); planning, kernels and the whole reduction live there.
//!
//! NOT COMPILED IN THIS REPOSITORY: the build image has no cargo/rustc.  The struct sizes the
//! pointer casts rely on are asserted at compile time below and mirrored by C `_Static_assert`s in
//! tests/test_layout_mirror.c.
#![cfg(feature = "b200")]

use ark_ec::AffineRepr;

use crate::{g1::G1Affine, g1::G1Projective, g2::G2Affine, g2::G2Projective, scalar::Scalar};

#[allow(non_camel_case_types)]
type c_int = core::ffi::c_int;

#[link(name = "b200msm")]
extern "C" {
    fn b200msm_g1(bases: *const u64, scalars: *const u64, n: usize, scalars_are_montgomery: c_int, out: *mut u64) -> c_int;
    fn b200msm_g2(bases: *const u64, scalars: *const u64, n: usize, scalars_are_montgomery: c_int, out: *mut u64) -> c_int;
    // resident bases (a proving key uploaded once) and their fixed-base window table
    fn b200msm_bases_upload(group: c_int, bases: *const u64, n: usize, handle: *mut *mut core::ffi::c_void) -> c_int;
    fn b200msm_bases_precompute(handle: *mut core::ffi::c_void, window_bits: c_int) -> c_int;
    fn b200msm_run(handle: *const core::ffi::c_void, scalars: *const u64, n: usize, scalars_are_montgomery: c_int, out: *mut u64) -> c_int;
    fn b200msm_bases_free(handle: *mut core::ffi::c_void) -> c_int;
    // page-lock a long-lived host buffer (a Vec is pageable memory: the driver stages it at a fraction of the PCIe rate)
    fn b200msm_host_register(ptr: *const core::ffi::c_void, bytes: usize) -> c_int;
    fn b200msm_host_unregister(ptr: *const core::ffi::c_void) -> c_int;
}

/// Page-locks `buf` for as long as the guard lives, so `msm` copies from it at full PCIe rate.
pub struct Pinned<'a, T>(&'a [T]);
impl<'a, T> Pinned<'a, T> {
    pub fn new(buf: &'a [T]) -> Result<Self, usize> {
        let rc = unsafe { b200msm_host_register(buf.as_ptr() as *const _, core::mem::size_of_val(buf)) };
        if rc != 0 { Err(0) } else { Ok(Self(buf)) }
    }
    pub fn as_slice(&self) -> &'a [T] {
        self.0
    }
}
impl<T> Drop for Pinned<'_, T> {
    fn drop(&mut self) {
        unsafe { b200msm_host_unregister(self.0.as_ptr() as *const _) };
    }
}

// The casts below are sound only because every wrapper is #[repr(transparent)] over the blst type
// (src/g1.rs:54-56,435-437; src/g2.rs:66-68,415-417; src/scalar.rs:23-25).
const _: () = assert!(core::mem::size_of::<G1Affine>() == 96);
const _: () = assert!(core::mem::size_of::<G1Projective>() == 144);
const _: () = assert!(core::mem::size_of::<G2Affine>() == 192);
const _: () = assert!(core::mem::size_of::<G2Projective>() == 288);
const _: () = assert!(core::mem::size_of::<Scalar>() == 32);
const _: () = assert!(core::mem::size_of::<ark_ff::BigInt<4>>() == 32);

/// Scalars as the caller holds them.
pub(crate) enum Scalars<'a> {
    /// `&[Scalar]`: Montgomery Fr, passed through untouched (no `into_bigint` pass,
    /// cf. reference src/g1.rs:624-627 + src/scalar.rs:450-463 which allocate per element).
    Montgomery(&'a [Scalar]),
    /// `&[BigInt<4>]`: canonical little-endian limbs (`msm_bigint`).
    BigInt(&'a [ark_ff::BigInt<4>]),
}

impl Scalars<'_> {
    fn len(&self) -> usize {
        match self {
            Scalars::Montgomery(s) => s.len(),
            Scalars::BigInt(s) => s.len(),
        }
    }
    fn ptr_and_flag(&self) -> (*const u64, c_int) {
        match self {
            Scalars::Montgomery(s) => (s.as_ptr() as *const u64, 1),
            Scalars::BigInt(s) => (s.as_ptr() as *const u64, 0),
        }
    }
}

/// `Err(min(len))` on a length mismatch (arkworks' convention for `msm`), `Err(0)` on any device
/// error (the reference GPU arm's convention, src/g1.rs:628-630). Never panics, never unwinds
/// across the FFI boundary (the reference `assert_eq!`s and `expect`s, src/gpu.rs:131,235-237).
pub(crate) fn msm_g1(bases: &[G1Affine], scalars: Scalars<'_>) -> Result<G1Projective, usize> {
    if bases.len() != scalars.len() {
        return Err(bases.len().min(scalars.len()));
    }
    let (sp, mont) = scalars.ptr_and_flag();
    let mut out = core::mem::MaybeUninit::<G1Projective>::uninit();
    let rc = unsafe { b200msm_g1(bases.as_ptr() as *const u64, sp, bases.len(), mont, out.as_mut_ptr() as *mut u64) };
    if rc != 0 {
        return Err(0);
    }
    Ok(unsafe { out.assume_init() })
}

pub(crate) fn msm_g2(bases: &[G2Affine], scalars: Scalars<'_>) -> Result<G2Projective, usize> {
    if bases.len() != scalars.len() {
        return Err(bases.len().min(scalars.len()));
    }
    let (sp, mont) = scalars.ptr_and_flag();
    let mut out = core::mem::MaybeUninit::<G2Projective>::uninit();
    let rc = unsafe { b200msm_g2(bases.as_ptr() as *const u64, sp, bases.len(), mont, out.as_mut_ptr() as *mut u64) };
    if rc != 0 {
        return Err(0);
    }
    Ok(unsafe { out.assume_init() })
}

// keep the generic bound the old entry point had so call sites outside g1.rs/g2.rs still name it
#[allow(dead_code)]
pub(crate) fn _assert_affine<G: AffineRepr>() {}

/// A proving key's G1 bases kept on the device(s): `upload` once, `msm` per proof (32 B/point of
/// H2D instead of 128). `precompute` turns them into a fixed-base window table (no Horner chain,
/// one-window bucket reduction). The reference has no counterpart: it re-uploads the bases and
/// rebuilds its program on every call (src/gpu.rs:149-150,233-237).
pub struct ResidentG1Bases {
    handle: *mut core::ffi::c_void,
    len: usize,
}
unsafe impl Send for ResidentG1Bases {}
unsafe impl Sync for ResidentG1Bases {}

impl ResidentG1Bases {
    pub fn upload(bases: &[G1Affine]) -> Result<Self, usize> {
        let mut handle = core::ptr::null_mut();
        let rc = unsafe { b200msm_bases_upload(0, bases.as_ptr() as *const u64, bases.len(), &mut handle) };
        if rc != 0 { Err(0) } else { Ok(Self { handle, len: bases.len() }) }
    }
    pub fn precompute(&mut self) -> Result<(), usize> {
        if unsafe { b200msm_bases_precompute(self.handle, 0) } != 0 { Err(0) } else { Ok(()) }
    }
    pub fn msm(&self, scalars: &[Scalar]) -> Result<G1Projective, usize> {
        if scalars.len() > self.len {
            return Err(self.len);
        }
        let mut out = core::mem::MaybeUninit::<G1Projective>::uninit();
        let rc = unsafe { b200msm_run(self.handle, scalars.as_ptr() as *const u64, scalars.len(), 1, out.as_mut_ptr() as *mut u64) };
        if rc != 0 { Err(0) } else { Ok(unsafe { out.assume_init() }) }
    }
}
impl Drop for ResidentG1Bases {
    fn drop(&mut self) {
        unsafe { b200msm_bases_free(self.handle) };
    }
}
