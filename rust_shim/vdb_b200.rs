//! Source-only Rust binding of include/vdb_b200.h for lab-1806-vec-db (cannot be compiled in the build
//! image: there is no Rust toolchain). A maintainer adds this file as `src/gpu/mod.rs`, links
//! `libvdb_b200.so` from build.rs (`println!("cargo:rustc-link-lib=dylib=vdb_b200")`) and swaps
//! `FlatIndex<T>` for `GpuFlatIndex<T>` behind the same traits (src/index_algorithm/mod.rs:35-154).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

use crate::{
    distance::{pq_table::PQTable, DistanceAlgorithm},
    index_algorithm::{prelude::*, CandidatePair},
    scalar::Scalar,
    vec_set::VecSet,
};

#[repr(C)] pub struct vdb_dataset { _p: [u8; 0] }
#[repr(C)] pub struct vdb_pq { _p: [u8; 0] }
#[repr(C)] pub struct vdb_ivf { _p: [u8; 0] }
#[repr(C)] pub struct vdb_hnsw { _p: [u8; 0] }

extern "C" {
    pub fn vdb_last_error() -> *const c_char;
    pub fn vdb_dataset_create(rows: *const c_void, n: u64, dim: u32, dtype: c_int, metric: c_int,
                              id_base: u64, out: *mut *mut vdb_dataset) -> c_int;
    pub fn vdb_dataset_append(ds: *mut vdb_dataset, rows: *const c_void, n: u64) -> c_int;
    pub fn vdb_dataset_swap_remove(ds: *mut vdb_dataset, idx: u64) -> c_int;
    pub fn vdb_dataset_destroy(ds: *mut vdb_dataset) -> c_int;
    pub fn vdb_flat_knn(ds: *const vdb_dataset, queries: *const c_void, nq: u32, k: u32,
                        ids: *mut u64, dist: *mut f32, counts: *mut u32) -> c_int;
    pub fn vdb_pq_create(ds: *const vdb_dataset, codebooks: *const c_void, m: u32, n_bits: u32,
                         codes: *mut u8, out: *mut *mut vdb_pq) -> c_int;
    pub fn vdb_pq_create_from_codes(ds: *const vdb_dataset, codebooks: *const c_void, m: u32, n_bits: u32,
                                    codes: *const u8, out: *mut *mut vdb_pq) -> c_int;
    pub fn vdb_pq_destroy(pq: *mut vdb_pq) -> c_int;
    pub fn vdb_pq_knn(ds: *const vdb_dataset, pq: *const vdb_pq, queries: *const c_void, nq: u32, k: u32,
                      ef: u32, ids: *mut u64, dist: *mut f32, counts: *mut u32) -> c_int;
    pub fn vdb_kmeans_assign(rows: *const c_void, n: u64, dim: u32, dtype: c_int, metric: c_int,
                             centroids: *const c_void, k: u32, sel_lo: u32, sel_hi: u32, out: *mut u32) -> c_int;
    pub fn vdb_kmeans_train(rows: *const c_void, n: u64, dim: u32, dtype: c_int, metric: c_int,
                            centroids: *mut c_void, k: u32, sel_lo: u32, sel_hi: u32, max_iter: u32,
                            tol: f32, iters: *mut u32) -> c_int;
    pub fn vdb_kmeans_pp_weights(rows: *const c_void, n: u64, dim: u32, dtype: c_int, metric: c_int,
                                 centroid: *const c_void, sel_lo: u32, sel_hi: u32, weights: *mut f32) -> c_int;
    pub fn vdb_ivf_create(ds: *const vdb_dataset, centroids: *const c_void, nlist: u32,
                          assign_out: *mut u32, out: *mut *mut vdb_ivf) -> c_int;
    pub fn vdb_ivf_destroy(ivf: *mut vdb_ivf) -> c_int;
    pub fn vdb_ivf_knn(ds: *const vdb_dataset, ivf: *const vdb_ivf, queries: *const c_void, nq: u32, k: u32,
                       n_probes: u32, ids: *mut u64, dist: *mut f32, counts: *mut u32) -> c_int;
    pub fn vdb_hnsw_build(ds: *const vdb_dataset, m: u32, ef_construction: u32, levels: *const u32, max_batch: u32,
                          out: *mut *mut vdb_hnsw) -> c_int;
    pub fn vdb_hnsw_destroy(h: *mut vdb_hnsw) -> c_int;
    pub fn vdb_hnsw_append(h: *mut vdb_hnsw, ds: *const vdb_dataset, new_levels: *const u32, max_batch: u32) -> c_int;
    pub fn vdb_hnsw_knn(ds: *const vdb_dataset, h: *const vdb_hnsw, queries: *const c_void, nq: u32, k: u32, ef: u32,
                        ids: *mut u64, dist: *mut f32, counts: *mut u32) -> c_int;
    pub fn vdb_hnsw_knn_pq(ds: *const vdb_dataset, h: *const vdb_hnsw, pq: *const vdb_pq, queries: *const c_void, nq: u32,
                           k: u32, ef: u32, ids: *mut u64, dist: *mut f32, counts: *mut u32) -> c_int;
    pub fn vdb_hnsw_links0(h: *const vdb_hnsw, links0: *mut u32, len0: *mut u32) -> c_int;
    pub fn vdb_hnsw_upper(h: *const vdb_hnsw, levels: *mut u32, ulinks: *mut u32, ulen: *mut u32) -> c_int;
    pub fn vdb_hnsw_create_from_graph(ds: *const vdb_dataset, m: u32, ef_construction: u32, levels: *const u32,
                                      links0: *const u32, len0: *const u32, ulinks: *const u32, ulen: *const u32,
                                      enter_point: i64, enter_level: i32, out: *mut *mut vdb_hnsw) -> c_int;
    pub fn vdb_gather_dist(ds: *const vdb_dataset, queries: *const c_void, nq: u32, cand_ids: *const u32,
                           cand_off: *const u64, out: *mut f32) -> c_int;
    pub fn vdb_row_cache(ds: *const vdb_dataset, out: *mut f32) -> c_int;
    pub fn vdb_calc_dist(a: *const c_void, b: *const c_void, count: u64, dim: u32, dtype: c_int,
                         metric: c_int, out: *mut f32) -> c_int;
}

fn check(rc: c_int) {
    // the reference's trait methods are infallible and panic on misuse (assert!), so does the shim
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(vdb_last_error()) }.to_string_lossy().into_owned();
        panic!("vdb_b200 error {rc}: {msg}");
    }
}
fn dtype_of<T: Scalar>() -> c_int { if std::mem::size_of::<T>() == 4 { 0 } else { 1 } }
fn metric_of(d: DistanceAlgorithm) -> c_int { match d { DistanceAlgorithm::L2Sqr => 0, DistanceAlgorithm::Cosine => 1 } }

/// Drop-in for `FlatIndex<T>`: host `VecSet` stays authoritative, the GPU mirror is a cache.
pub struct GpuFlatIndex<T: Scalar> {
    pub dist: DistanceAlgorithm,
    pub vec_set: VecSet<T>,
    ds: *mut vdb_dataset,
}
unsafe impl<T: Scalar> Send for GpuFlatIndex<T> {}
unsafe impl<T: Scalar> Sync for GpuFlatIndex<T> {} // vdb_*_knn are re-entrant on one handle

impl<T: Scalar> GpuFlatIndex<T> {
    fn collect(ids: Vec<u64>, dist: Vec<f32>, count: u32) -> Vec<CandidatePair> {
        (0..count as usize).map(|j| CandidatePair::new(ids[j] as usize, dist[j])).collect()
    }
    /// Additive batch entry (the trait call is nq == 1).
    pub fn knn_batch(&self, queries: &[T], nq: usize, k: usize) -> Vec<Vec<CandidatePair>> {
        let (mut ids, mut dist, mut cnt) = (vec![0u64; nq * k], vec![0f32; nq * k], vec![0u32; nq]);
        check(unsafe { vdb_flat_knn(self.ds, queries.as_ptr() as _, nq as u32, k as u32,
                                    ids.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr()) });
        (0..nq).map(|q| Self::collect(ids[q * k..(q + 1) * k].to_vec(), dist[q * k..(q + 1) * k].to_vec(), cnt[q])).collect()
    }
}
impl<T: Scalar> std::ops::Index<usize> for GpuFlatIndex<T> {
    type Output = [T];
    fn index(&self, i: usize) -> &[T] { &self.vec_set[i] }
}
impl<T: Scalar> IndexIter<T> for GpuFlatIndex<T> {
    fn dim(&self) -> usize { self.vec_set.dim() }
    fn len(&self) -> usize { self.vec_set.len() }
}
impl<T: Scalar> IndexFromVecSet<T> for GpuFlatIndex<T> {
    type Config = ();
    fn from_vec_set(vec_set: VecSet<T>, dist: DistanceAlgorithm, _: (), _: &mut impl rand::Rng) -> Self {
        let mut ds = std::ptr::null_mut();
        check(unsafe { vdb_dataset_create(vec_set.as_slice().as_ptr() as _, vec_set.len() as u64,
                                          vec_set.dim() as u32, dtype_of::<T>(), metric_of(dist), 0, &mut ds) });
        Self { dist, vec_set, ds }
    }
}
impl<T: Scalar> IndexKNN<T> for GpuFlatIndex<T> {
    /// replaces FlatIndex::knn (src/index_algorithm/flat_index.rs:48-57)
    fn knn(&self, query: &[T], k: usize) -> Vec<CandidatePair> {
        self.knn_batch(query, 1, k).pop().unwrap()
    }
}
impl<T: Scalar> IndexPQ<T> for GpuFlatIndex<T> {
    /// replaces FlatIndex::knn_pq (flat_index.rs:84-104). The device PQ mirror is created once per
    /// PQTable (vdb_pq_create_from_codes with the table's codebooks + encoded_vec_set) and cached on it.
    fn knn_pq(&self, query: &[T], k: usize, ef: usize, pq_table: &PQTable<T>) -> Vec<CandidatePair> {
        let pq = pq_table.gpu_mirror(self.ds); // lazily built, dropped with the table
        let (mut ids, mut dist, mut cnt) = (vec![0u64; k], vec![0f32; k], vec![0u32; 1]);
        check(unsafe { vdb_pq_knn(self.ds, pq, query.as_ptr() as _, 1, k as u32, ef as u32,
                                  ids.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr()) });
        Self::collect(ids, dist, cnt[0])
    }
}
impl<T: Scalar> Drop for GpuFlatIndex<T> {
    fn drop(&mut self) { unsafe { vdb_dataset_destroy(self.ds); } }
}
