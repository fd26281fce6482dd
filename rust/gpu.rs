//! Drop-in replacement for ark-blst's `src/gpu.rs` (reference src/gpu.rs:1-286).
//!
//! The reference file wraps ec-gpu-gen's generated kernel: per call it enumerates devices, builds a
//! `Program`, uploads bases and exponents, launches one kernel, downloads 18 944 partials and folds
//! them on the host (src/gpu.rs:126-241).  This file shrinks to an `extern "C"` block over
//! `libb200msm.so` (include/b200msm.h); planning, kernels and the whole reduction live there.
//!
//! Install: copy this file over `src/gpu.rs`, `rust/build.rs` over `build.rs`, then
//! `patch -p1 < rust/ark-blst-b200.patch` (Cargo.toml, src/lib.rs, src/g1.rs, src/g2.rs; applies
//! cleanly to the reference tree — tests/test_rust_patch.py dry-runs it).
//!
//! NOT COMPILED IN THIS REPOSITORY: the build image has no cargo/rustc.  The struct sizes the
//! pointer casts rely on are asserted at compile time below and mirrored by C `_Static_assert`s in
//! tests/test_layout_mirror.c.
#![cfg(feature = "b200")]

use core::ffi::{c_char, c_int, c_void};
use core::marker::PhantomData;

use crate::{g1::G1Affine, g1::G1Projective, g2::G2Affine, g2::G2Projective, scalar::Scalar};

#[link(name = "b200msm")]
extern "C" {
    // engine binding: which GPUs a sharded MSM runs on (reference: Device::all()[0], src/gpu.rs:233-234)
    fn b200msm_init(first_device: c_int, n_devices: c_int) -> c_int;
    fn b200msm_shutdown();
    fn b200msm_device_count() -> c_int;
    fn b200msm_last_error() -> *const c_char;
    fn b200msm_g1(bases: *const u64, scalars: *const u64, n: usize, scalars_are_montgomery: c_int, out: *mut u64) -> c_int;
    fn b200msm_g2(bases: *const u64, scalars: *const u64, n: usize, scalars_are_montgomery: c_int, out: *mut u64) -> c_int;
    // resident bases (a proving key uploaded once) and their fixed-base window table
    fn b200msm_bases_upload(group: c_int, bases: *const u64, n: usize, handle: *mut *mut c_void) -> c_int;
    fn b200msm_bases_precompute(handle: *mut c_void, window_bits: c_int) -> c_int;
    fn b200msm_run(handle: *const c_void, scalars: *const u64, n: usize, scalars_are_montgomery: c_int, out: *mut u64) -> c_int;
    fn b200msm_bases_free(handle: *mut c_void) -> c_int;
    // page-lock a long-lived host buffer (a Vec is pageable memory: the driver stages it at a fraction of the PCIe rate)
    fn b200msm_host_register(ptr: *const c_void, bytes: usize) -> c_int;
    fn b200msm_host_unregister(ptr: *const c_void) -> c_int;
}

/// Bind the engine to `n_devices` GPUs starting at `first_device` (`n_devices = 0`: all visible).
/// Every later `msm` shards its points evenly over them and adds the per-GPU partials on the
/// first one.  Optional: without it the first `msm` binds the current device only.  Binding a
/// different range while bound is an error (`shutdown_devices` first).
pub fn init_devices(first_device: usize, n_devices: usize) -> Result<usize, usize> {
    let rc = unsafe { b200msm_init(first_device as c_int, n_devices as c_int) };
    if rc != 0 { Err(0) } else { Ok(device_count()) }
}
/// Release every device buffer, stream and worker thread of the engine.  Resident bases uploaded
/// before must be dropped or re-uploaded: they do not survive a re-binding.
pub fn shutdown_devices() {
    unsafe { b200msm_shutdown() }
}
pub fn device_count() -> usize {
    unsafe { b200msm_device_count() as usize }
}
/// The library's message for the calling thread's most recent failure.
pub fn last_error() -> String {
    let p = unsafe { b200msm_last_error() };
    if p.is_null() {
        return String::new();
    }
    unsafe { std::ffi::CStr::from_ptr(p) }.to_string_lossy().into_owned()
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

/// The two groups the engine knows, as the C-ABI numbers them (B200MSM_G1 / B200MSM_G2).
pub trait MsmGroup {
    type Affine;
    type Projective;
    const ID: c_int;
}
pub struct G1Tag;
pub struct G2Tag;
impl MsmGroup for G1Tag {
    type Affine = G1Affine;
    type Projective = G1Projective;
    const ID: c_int = 0;
}
impl MsmGroup for G2Tag {
    type Affine = G2Affine;
    type Projective = G2Projective;
    const ID: c_int = 1;
}

/// A proving key's bases kept on the device(s): `upload` once (sharded evenly over the bound
/// GPUs), `msm` per proof (32 B/point of H2D instead of 128 / 224). `precompute` turns them into a
/// fixed-base window table (no Horner chain, one-window bucket reduction). The reference has no
/// counterpart: it re-uploads the bases and rebuilds its program on every call
/// (src/gpu.rs:149-150,233-237).
pub struct ResidentBases<G: MsmGroup> {
    handle: *mut c_void,
    len: usize,
    _g: PhantomData<G>,
}
pub type ResidentG1Bases = ResidentBases<G1Tag>;
pub type ResidentG2Bases = ResidentBases<G2Tag>;
unsafe impl<G: MsmGroup> Send for ResidentBases<G> {}
unsafe impl<G: MsmGroup> Sync for ResidentBases<G> {}

impl<G: MsmGroup> ResidentBases<G> {
    pub fn upload(bases: &[G::Affine]) -> Result<Self, usize> {
        let mut handle = core::ptr::null_mut();
        let rc = unsafe { b200msm_bases_upload(G::ID, bases.as_ptr() as *const u64, bases.len(), &mut handle) };
        if rc != 0 { Err(0) } else { Ok(Self { handle, len: bases.len(), _g: PhantomData }) }
    }
    pub fn len(&self) -> usize {
        self.len
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
    pub fn precompute(&mut self) -> Result<(), usize> {
        if unsafe { b200msm_bases_precompute(self.handle, 0) } != 0 { Err(0) } else { Ok(()) }
    }
    /// Σ sᵢ·Pᵢ over the first `scalars.len()` resident bases. `Err(len)` when more scalars than
    /// bases are passed, `Err(0)` on a device error.
    pub fn msm(&self, scalars: &[Scalar]) -> Result<G::Projective, usize> {
        if scalars.len() > self.len {
            return Err(self.len);
        }
        let mut out = core::mem::MaybeUninit::<G::Projective>::uninit();
        let rc = unsafe { b200msm_run(self.handle, scalars.as_ptr() as *const u64, scalars.len(), 1, out.as_mut_ptr() as *mut u64) };
        if rc != 0 { Err(0) } else { Ok(unsafe { out.assume_init() }) }
    }
}
impl<G: MsmGroup> Drop for ResidentBases<G> {
    fn drop(&mut self) {
        unsafe { b200msm_bases_free(self.handle) };
    }
}
